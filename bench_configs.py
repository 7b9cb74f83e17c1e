"""bench_configs.py — BASELINE.json's OTHER configurations, measured by bench.py on one B200 next to the headline
(configs[1]) and reported inside its JSON line under "configs":

  cfg1          configs[0]  uniform 100k / 100k, knn=1 (the reference's own CPU-runnable case)
  cfg3_knn16    configs[2]  cfg2 cloud, knn=16
  cfg3_radius   configs[2]  cfg2 cloud, search_radius r^2 = 0.01, unsorted (visit order)
  box           (a11)       search_box, 1M boxes of 0.4 m edge centred on cfg2 queries
  cfg4_exact    configs[3]  sift-shape 1M x 128, 10k queries, knn=10, runtime-dim path
  cfg4_approx   configs[3]  the same with e = 2.25 (search_approximate_knn)
  cfg4_forest   (f4)        kd_forest on the same data: 8 and 16 trees, max_leaf_size 32, 64 leaves per tree — the
                            settings of the reference's examples/kd_forest/kd_forest.cpp:113-123 for SIFT — recall
                            against the exact neighbour, q/s, and bit parity with the oracle's forest on a reduced set

Every entry: `value` (device-resident, CUDA events or the device-pointer call), `e2e` (public host API, host
buffers), `parity` against the unmodified reference (oracle/_ref; the oracle port where that is absent) — index
for index and bit for bit, ragged results IN VISIT ORDER —, `cpu_baseline`, and a `roofline` whose algorithmic
bytes follow SURVEY.md §8d with counters from the oracle's instrumented reference traversal.

The oracle is used here only as checker / CPU baseline / byte counter, never inside a timed region.
"""
import ctypes as C
import time

import numpy as np


def _cpu_tree(O, pts, leaf=10):
    return (O.RefTree(pts, leaf), "reference") if O.ref_available() else (O.OracleTree(pts, leaf), "port")


def _timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def _roof(bytes_per_unit, units, kernel_ms, peak, peak_kind, kernel, counters, note=None):
    if not kernel_ms or kernel_ms <= 0:
        return None
    achieved = bytes_per_unit * units / (kernel_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "kernel": kernel, "kernel_ms": kernel_ms, "algorithmic_bytes_per_unit": bytes_per_unit,
         "per_unit_branches_leaves_points_results": counters, "peak_source": peak_kind}
    if note:
        r["note"] = note
    return r


def _resident_knn(torch, _lib, tree, q, k, e, reps, warmup=3):
    """Device pointers in and out, CUDA events around `reps` calls + the traversal kernel's own event spans."""
    L = _lib.lib()
    dev = torch.device("cuda", tree._device)
    qd = torch.from_numpy(q).to(dev)
    words = 2 if q.dtype == np.float32 else 4
    out = torch.empty((len(q), k, words), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    _lib.check(L.pico_b200_set_stream(C.c_void_p(stream.cuda_stream)))
    fl = _lib.FLAG_DEVICE_POINTERS | _lib.FLAG_ASYNC

    def step():
        _lib.check(L.pico_b200_knn(tree._h, C.c_void_p(qd.data_ptr()), len(q), q.shape[1], k, float(e),
                                   C.c_void_p(out.data_ptr()), fl, None))
    try:
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            _lib.check(L.pico_b200_profile_begin())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                step()
            e1.record(stream)
            torch.cuda.synchronize()
            ms, n = C.c_double(), C.c_uint64()
            _lib.check(L.pico_b200_profile_end(C.byref(ms), C.byref(n)))
    finally:
        _lib.check(L.pico_b200_set_stream(None))
    return e0.elapsed_time(e1) / reps, ms.value / reps, out.cpu().numpy()


def _knn_parity(got_words, want, k):
    got = got_words.reshape(len(want), k, 2)
    return {"queries": int(len(want)), "distances_bit_equal": bool(np.array_equal(got[..., 1].view(np.float32),
                                                                                  want["distance"])),
            "index_mismatches": int(np.count_nonzero(got[..., 0] != want["index"]))}


def knn_config(ctx, name, tree, tree_pts, q, k, ref, kind, e=0.0, reps=10, e2e_reps=3, cpu_sample=None,
               counter_sample=100_000, unit_scale=1e6, unit="Mq/s", kernel_name=None, warmup=3):
    torch, _lib, O = ctx["torch"], ctx["lib"], ctx["oracle"]
    nq = len(q)
    ms, kernel_ms, res = _resident_knn(torch, _lib, tree, q, k, e, reps, warmup)
    qp = torch.from_numpy(q).pin_memory().numpy()
    outp = torch.empty((nq, k, 2), dtype=torch.int32).pin_memory().numpy().view(tree.dtype_neighbor).reshape(nq, k)
    args = [e] if e else []
    e2e_s = _timed(lambda: tree.search_knn(qp, k, *args, outp), e2e_reps)
    threads = O.max_threads()
    ns = nq if cpu_sample is None else min(nq, cpu_sample)
    qs = np.ascontiguousarray(q[:ns])
    t0 = time.perf_counter()
    want = ref.search_knn(qs, k, e=e, threads=threads)
    cpu_s = time.perf_counter() - t0
    parity = _knn_parity(res.reshape(nq, k, 2)[:ns], want, k)
    parity["e2e_equals_resident"] = bool(np.array_equal(outp.view(np.int32).reshape(nq, k, 2), res.reshape(nq, k, 2)))
    line = {"config": name, "value": nq / ms * 1e3 / unit_scale, "unit": unit, "ms_per_step": ms, "steps": reps,
            "n_tree": int(len(tree_pts)), "sdim": int(tree_pts.shape[1]), "n_query": int(nq), "k": int(k), "e": float(e),
            "e2e": {"value": nq / e2e_s / unit_scale, "unit": unit, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(q.nbytes), "d2h_bytes_per_step": int(nq * k * 8)},
            "parity": parity,
            "cpu_baseline": {"value": ns / cpu_s / unit_scale, "unit": unit, "cores": threads, "kind": kind,
                             "sample": "first %d of %d queries, all host threads, once" % (ns, nq)}}
    if counter_sample:
        oc = ctx["oracle_tree"](tree_pts)
        cs = min(nq, counter_sample)
        sel = np.ascontiguousarray(q[:: max(nq // cs, 1)][:cs])
        _, cnt = oc.search_knn(sel, k, e=e, counters=True, threads=1)
        c = (cnt / len(sel)).tolist()
        sdim = tree_pts.shape[1]
        node_b = 16
        b = 4 * sdim + 8 * k + c[0] * node_b + c[2] * (4 * sdim + 4)
        line["roofline"] = _roof(b, nq, kernel_ms, ctx["peak"], ctx["peak_kind"], kernel_name, c + [float(k)])
    return line


def radius_config(ctx, tree, tree_pts, q, ref, kind, r2=0.01, sample=1_000_000, counter_sample=50_000):
    torch, _lib, O = ctx["torch"], ctx["lib"], ctx["oracle"]
    L = _lib.lib()
    nq = len(q)
    dev = torch.device("cuda", tree._device)
    qd = torch.from_numpy(q).to(dev)
    offs = torch.empty(nq + 1, dtype=torch.int64, device=dev)

    def resident():
        hits = C.c_void_p()
        _lib.check(L.pico_b200_radius(tree._h, C.c_void_p(qd.data_ptr()), nq, 3, float(r2), 0.0,
                                      C.c_void_p(offs.data_ptr()), C.byref(hits), _lib.FLAG_DEVICE_POINTERS, None))
        L.pico_b200_free_device(hits)
    resident()
    torch.cuda.synchronize()
    _lib.check(L.pico_b200_profile_begin())
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        resident()
    torch.cuda.synchronize()
    res_s = (time.perf_counter() - t0) / reps
    ms, n = C.c_double(), C.c_uint64()
    _lib.check(L.pico_b200_profile_end(C.byref(ms), C.byref(n)))
    kernel_ms = ms.value / reps
    qp = torch.from_numpy(q).pin_memory().numpy()
    box = {}

    def run():
        box["nns"] = tree.search_radius(qp, r2)
    e2e_s = _timed(run, 2)
    nns = box["nns"]
    total = int(nns._offsets[-1])
    threads = O.max_threads()
    ns = min(nq, sample)
    qs = np.ascontiguousarray(q[:ns])
    t0 = time.perf_counter()
    w_offs, w_flat = ref.search_radius(qs, r2, threads=threads) if kind == "reference" else ref.search_radius(qs, r2)
    cpu_s = time.perf_counter() - t0
    m = int(w_offs[-1])
    parity = {"queries": int(ns), "offsets_equal": bool(np.array_equal(nns._offsets[:ns + 1], w_offs)),
              "indices_equal_in_visit_order": bool(np.array_equal(nns._flat["index"][:m], w_flat["index"])),
              "distances_bit_equal_in_visit_order": bool(np.array_equal(nns._flat["distance"][:m], w_flat["distance"])),
              "device_offsets_equal_host_path": bool(np.array_equal(offs.cpu().numpy().astype(np.uint64), nns._offsets))}
    oc = ctx["oracle_tree"](tree_pts)
    cs = min(nq, counter_sample)
    sel = np.ascontiguousarray(q[:: max(nq // cs, 1)][:cs])
    c = (oc.radius_counters(sel, r2).astype(np.float64) / len(sel)).tolist()
    b = 12 + 8 * c[3] + 16 * c[0] + 16 * c[2]
    return {"config": "configs[2]: cfg2 cloud, search_radius r^2=%g, unsorted" % r2, "value": nq / res_s / 1e6,
            "unit": "Mq/s", "ms_per_step": res_s * 1e3, "steps": reps, "n_tree": int(len(tree_pts)), "n_query": int(nq),
            "mean_hits_per_query": total / nq, "total_hits": total,
            "value_note": "device pointers in, offsets and hits left in HBM (count pass, scan, fill pass)",
            "e2e": {"value": nq / e2e_s / 1e6, "unit": "Mq/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(q.nbytes), "d2h_bytes_per_step": int(total * 8 + (nq + 1) * 8)},
            "parity": parity,
            "cpu_baseline": {"value": ns / cpu_s / 1e6, "unit": "Mq/s", "cores": threads if kind == "reference" else 1,
                             "kind": kind, "sample": "first %d of %d queries, once" % (ns, nq)},
            "roofline": _roof(b, nq, kernel_ms, ctx["peak"], ctx["peak_kind"],
                              "radius_thread_kernel<float,3,COUNT> + <FILL>", c,
                              "kernel_ms sums the count pass and the fill pass; the byte model counts ONE traversal")}


def box_config(ctx, tree, tree_pts, q, ref, kind, nb=1_000_000, half=0.2, sample=200_000, counter_sample=20_000):
    torch, _lib, O = ctx["torch"], ctx["lib"], ctx["oracle"]
    L = _lib.lib()
    dev = torch.device("cuda", tree._device)
    nb = min(nb, len(q))
    mins = np.ascontiguousarray(q[:nb] - np.float32(half))
    maxs = np.ascontiguousarray(q[:nb] + np.float32(half))
    boxes = np.empty((2 * nb, 3), np.float32)
    boxes[0::2], boxes[1::2] = mins, maxs
    dmin, dmax = torch.from_numpy(mins).to(dev), torch.from_numpy(maxs).to(dev)
    offs = torch.empty(nb + 1, dtype=torch.int64, device=dev)

    def resident():
        hits = C.c_void_p()
        _lib.check(L.pico_b200_box(tree._h, C.c_void_p(dmin.data_ptr()), C.c_void_p(dmax.data_ptr()), nb, 3,
                                   C.c_void_p(offs.data_ptr()), C.byref(hits), _lib.FLAG_DEVICE_POINTERS, None))
        L.pico_b200_free_device(hits)
    resident()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        resident()
    torch.cuda.synchronize()
    res_s = (time.perf_counter() - t0) / reps
    box = {}

    def run():
        box["r"] = tree.search_box(boxes)
    e2e_s = _timed(run, 2)
    res = box["r"]
    total = int(res._offsets[-1])
    threads = O.max_threads()
    ns = min(nb, sample)
    t0 = time.perf_counter()
    if kind == "reference":
        w_offs, w_flat = ref.search_box(mins[:ns], maxs[:ns], threads=threads)
    else:
        w_offs, w_flat = ref.search_box(mins[:ns], maxs[:ns])
    cpu_s = time.perf_counter() - t0
    parity = {"boxes": int(ns), "offsets_equal": bool(np.array_equal(res._offsets[:ns + 1], w_offs)),
              "indices_equal_in_dfs_order": bool(np.array_equal(res._flat[:int(w_offs[-1])], w_flat)),
              "device_offsets_equal_host_path": bool(np.array_equal(offs.cpu().numpy().astype(np.uint64), res._offsets))}
    oc = ctx["oracle_tree"](tree_pts)
    cs = min(nb, counter_sample)
    c = (oc.box_counters(mins[:cs], maxs[:cs]).astype(np.float64) / cs).tolist()
    b = 24 + 16 * c[0] + 16 * c[2] + 8 * c[3]
    st = tree.last_stats
    return {"config": "search_box (a11): %d boxes of %.1f m edge centred on cfg2 queries" % (nb, 2 * half),
            "value": nb / res_s / 1e6, "unit": "Mboxes/s", "ms_per_step": res_s * 1e3, "steps": reps,
            "n_tree": int(len(tree_pts)), "n_boxes": int(nb), "mean_hits_per_box": total / nb, "total_hits": total,
            "value_note": "device pointers in, offsets and indices left in HBM (count pass, scan, fill pass)",
            "e2e": {"value": nb / e2e_s / 1e6, "unit": "Mboxes/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(boxes.nbytes), "d2h_bytes_per_step": int(total * 4 + (nb + 1) * 8)},
            "parity": parity,
            "cpu_baseline": {"value": ns / cpu_s / 1e6, "unit": "Mboxes/s",
                             "cores": threads if kind == "reference" else 1, "kind": kind,
                             "sample": "first %d of %d boxes, once" % (ns, nb)},
            "roofline": _roof(b, nb, st.kernel_ms if st.kernel_ms > 0 else res_s * 1e3, ctx["peak"], ctx["peak_kind"],
                              "box_warp_kernel<float,PACKED> count + fill", c,
                              "24 B box + 16 B per branch node + 16 B per tested point + 8 B per reported index "
                              "(index read + write); kernel_ms = count + scan + fill of the host-buffer call")}


def forest_config(ctx, pts, q, exact_index, n_trees, max_leaf_size=32, max_leaves=64, reduced=100_000):
    """kd_forest search_nn over all queries; recall = share of queries whose exact neighbour is found."""
    torch, O, pt = ctx["torch"], ctx["oracle"], ctx["pt"]
    rng = np.random.default_rng(17)
    rot = rng.normal(size=(n_trees, pts.shape[1]))
    rot = (rot / np.linalg.norm(rot, axis=1, keepdims=True)).astype(pts.dtype)
    t0 = time.perf_counter()
    f = pt.KdForest(pts, max_leaf_size, n_trees, rotations=rot)
    build_wall = time.perf_counter() - t0
    f.search_nn(q[:256], max_leaves)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        got = f.search_nn(q, max_leaves)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, kernel_ms = dt, f.last_stats.kernel_ms
    recall = float(np.mean(got["index"][:, 0] == exact_index))
    info = f.info()
    line = {"config": "kd_forest (f4) on configs[3] data: %d trees, max_leaf_size %d, %d leaves per tree, search_nn"
                      % (n_trees, max_leaf_size, max_leaves),
            "value": len(q) / best, "unit": "q/s", "ms_per_step": best * 1e3, "kernel_ms": kernel_ms, "steps": 3,
            "n_tree": int(len(pts)), "sdim": int(pts.shape[1]), "n_query": int(len(q)), "recall_vs_exact_nn": recall,
            "reference_precision_note": "examples/kd_forest/kd_forest.cpp:118-121 quotes ~0.884 (8 trees) / ~0.940 "
                                        "(16 trees) on real SIFT with these settings; this is synthetic sift-shape data",
            "build_ms_device": info["build_ms"], "build_wall_s": build_wall, "forest_device_bytes": info["device_bytes"],
            "e2e": {"value": len(q) / best, "unit": "q/s", "ms_per_step": best * 1e3, "h2d_bytes_per_step": int(q.nbytes),
                    "d2h_bytes_per_step": int(len(q) * 8), "note": "the timed call takes host arrays"}}
    del f
    # bit parity + CPU baseline on a reduced set (a CPU forest over 1M x 128 takes minutes to build)
    rp, rq = np.ascontiguousarray(pts[:reduced]), np.ascontiguousarray(q[:512])
    nt = min(n_trees, 4)
    fr = pt.KdForest(rp, max_leaf_size, nt, rotations=rot[:nt])
    t0 = time.perf_counter()
    of = O.OracleForest(rp, rot[:nt], max_leaf_size)
    cpu_build = time.perf_counter() - t0
    threads = O.max_threads()
    t0 = time.perf_counter()
    want = of.search_knn(rq, 1, max_leaves, threads=threads)
    cpu_s = time.perf_counter() - t0
    gotr = fr.search_nn(rq, max_leaves)
    line["parity"] = {"reduced_n_tree": int(reduced), "trees": nt, "queries": int(len(rq)),
                      "indices_equal": bool(np.array_equal(gotr["index"], want["index"])),
                      "distances_bit_equal": bool(np.array_equal(gotr["distance"], want["distance"]))}
    line["cpu_baseline"] = {"value": len(rq) / cpu_s, "unit": "q/s", "cores": threads, "kind": "port",
                            "sample": "oracle forest of %d trees over the first %d points (built in %.1f s on one "
                                      "thread), %d queries" % (nt, reduced, cpu_build, len(rq))}
    return line


def run_all(ctx, tree, tree_pts, q, only=None):
    """ctx: {"torch", "lib" (pico_tree_b200._lib), "oracle" (oracle.oracle), "pt" (pico_tree_b200), "peak", "peak_kind"}.
    `tree` / `tree_pts` / `q` are the headline's cfg2 objects (reused: one build, one reference tree)."""
    O, pt = ctx["oracle"], ctx["pt"]
    from pico_tree_b200 import datasets as D
    cache = {}

    def oracle_tree(pts):
        key = id(pts)
        if key not in cache:
            cache.clear()
            cache[key] = O.OracleTree(pts, 10)
        return cache[key]
    ctx = dict(ctx, oracle_tree=oracle_tree)
    out = {}
    want = only or ["cfg1", "cfg3_knn16", "cfg3_radius", "box", "cfg4_exact", "cfg4_approx", "cfg4_forest"]

    def guarded(key, fn):
        if key not in want and not (key.startswith("cfg4_forest") and "cfg4_forest" in want):
            return
        t0 = time.perf_counter()
        try:
            out[key] = fn()
        except Exception as ex:  # a failing side configuration must not take the headline line down
            out[key] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        out[key]["wall_s"] = time.perf_counter() - t0

    if "cfg1" in want:
        p1, q1 = D.uniform(100_000, 3, seed=1), D.uniform(100_000, 3, seed=2)
        t1 = pt.KdTree(p1, pt.Metric.L2Squared, 10)
        r1, k1 = _cpu_tree(O, p1)

        def cfg1():
            line = knn_config(ctx, "configs[0]: uniform 100k / 100k, knn=1", t1, p1, q1, 1, r1, k1, reps=50, e2e_reps=20,
                              kernel_name="nn_fat_kernel<float,3>")
            t0 = time.perf_counter()
            r1.search_knn(q1, 1, threads=1)
            line["cpu_baseline"]["single_thread_value"] = len(q1) / (time.perf_counter() - t0) / 1e6
            return line
        guarded("cfg1", cfg1)
        del t1, r1
    if any(kname in want for kname in ("cfg3_knn16", "cfg3_radius", "box")):
        ref, kind = _cpu_tree(O, tree_pts)
        guarded("cfg3_knn16", lambda: knn_config(ctx, "configs[2]: cfg2 cloud, knn=16", tree, tree_pts, q, 16, ref, kind,
                                                 reps=10, e2e_reps=3, kernel_name="knn_thread_kernel<float,3,16,FAST>"))
        guarded("cfg3_radius", lambda: radius_config(ctx, tree, tree_pts, q, ref, kind))
        guarded("box", lambda: box_config(ctx, tree, tree_pts, q, ref, kind))
        del ref
    if "cfg4_forest" in want:
        p4, q4 = D.sift_shape(1_000_000, seed=1), D.sift_shape(10_000, seed=2)
        exact = pt.KdTree(p4, pt.Metric.L2Squared, 10).search_knn(q4, 1)["index"][:, 0]
        for nt in (8, 16):
            guarded("cfg4_forest_%d" % nt, lambda nt=nt: forest_config(ctx, p4, q4, exact, nt))
    if "cfg4_exact" in want or "cfg4_approx" in want:
        p4, q4 = D.sift_shape(1_000_000, seed=1), D.sift_shape(10_000, seed=2)
        t0 = time.perf_counter()
        t4 = pt.KdTree(p4, pt.Metric.L2Squared, 10)
        build_wall = time.perf_counter() - t0
        r4, k4 = _cpu_tree(O, p4)
        same_tree = None
        if k4 == "reference" and hasattr(r4, "saved"):
            same_tree = bool(t4._saved_stream() == bytes(r4.saved()))
        for key, e in (("cfg4_exact", 0.0), ("cfg4_approx", 2.25)):
            def cfg4(e=e):
                line = knn_config(ctx, "configs[3]: sift-shape 1M x 128, 10k queries, knn=10, runtime-dim path, " +
                                  ("search_approximate_knn e=2.25" if e else "exact"), t4, p4, q4, 10, r4, k4, e=e,
                                  reps=1, e2e_reps=1, cpu_sample=96, counter_sample=16, unit_scale=1.0, unit="q/s",
                                  kernel_name="knn_warp_kernel<float,ROWS,REGLIST>", warmup=1)
                line["tree_height"] = t4.info()["height"]
                line["build_ms_device"] = t4.info()["build_ms"]
                line["build_wall_s"] = build_wall
                line["parity"]["device_built_tree_stream_equals_reference"] = same_tree
                return line
            guarded(key, cfg4)
    return out

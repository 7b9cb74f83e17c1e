#!/usr/bin/env python
"""Measures BASELINE.json's other configurations on one B200 (bench.py covers configs[1], the headline):

  cfg1   uniform 100k / 100k, knn=1                (the reference's CPU-runnable case)
  cfg2s  cfg2 with the queries SHUFFLED            (worst case for caches; the batch is Z-ordered on device)
  cfg3a  cfg2 cloud, knn=16
  cfg3b  cfg2 cloud, search_radius r^2 = 0.01      (mean hits/query reported)
  cfg4   sift-shape 1M x 128, 10k queries, knn=10  (runtime-dim path), exact and e = metric(1.5)
  build  device build of the cfg2 tree, three rules

Every line: resident Mq/s (device pointers, CUDA events) where the call supports it, end to end Mq/s through
the public host API, the reference on the box's host threads for a bounded sample, and a parity check of the
GPU result against that reference sample. One JSON object per line on stdout.

    python profiles/run_configs.py [cfg1 cfg2s cfg3a cfg3b cfg4 build]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import pico_tree_b200 as pt  # noqa: E402
from oracle import oracle as O  # noqa: E402
from pico_tree_b200 import _lib, datasets as D  # noqa: E402


def cpu_tree(pts, leaf=10):
    return (O.RefTree(pts, leaf), "reference") if O.ref_available() else (O.OracleTree(pts, leaf), "port")


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def resident_knn(tree, q, k, e=0.0, reps=20, flags=0):
    """Device pointers in, device pointers out, CUDA events around `reps` calls."""
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    qd = torch.from_numpy(q).to(dev)
    rec = 2 if q.dtype == np.float32 else 4
    out = torch.empty((len(q), k, rec), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    _lib.check(L.pico_b200_set_stream(C.c_void_p(stream.cuda_stream)))
    fl = _lib.FLAG_DEVICE_POINTERS | _lib.FLAG_ASYNC | flags

    def step():
        _lib.check(L.pico_b200_knn(tree._h, C.c_void_p(qd.data_ptr()), len(q), q.shape[1], k, float(e),
                                   C.c_void_p(out.data_ptr()), fl, None))
    with torch.cuda.stream(stream):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
    _lib.check(L.pico_b200_set_stream(None))
    ms = e0.elapsed_time(e1) / reps
    res = out.cpu().numpy()
    return ms, res


def knn_line(name, tree_pts, q, k, e=0.0, cpu_sample=500_000, reps=20, note=""):
    t0 = time.perf_counter()
    tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
    build_wall = time.perf_counter() - t0
    info = tree.info()
    ms, res = resident_knn(tree, q, k, e, reps)
    qp = torch.from_numpy(q).pin_memory().numpy()
    outp = torch.empty((len(q), k, 2), dtype=torch.int32).pin_memory().numpy().view(tree.dtype_neighbor).reshape(len(q), k)
    e2e_s = timed(lambda: tree.search_knn(qp, k, *( [e] if e else [] ), outp), max(3, reps // 4))
    # CPU reference on a bounded sample + parity of that sample
    ref, kind = cpu_tree(tree_pts)
    threads = O.max_threads()
    ns = min(len(q), cpu_sample)
    qs = np.ascontiguousarray(q[:ns])
    t0 = time.perf_counter()
    want = ref.search_knn(qs, k, e=e, threads=threads)
    cpu_s = time.perf_counter() - t0
    n1 = min(ns, max(1000, cpu_sample // 10))
    t0 = time.perf_counter()
    ref.search_knn(np.ascontiguousarray(qs[:n1]), k, e=e, threads=1)
    cpu1_s = time.perf_counter() - t0
    got = res.reshape(len(q), k, 2)[:ns]
    dist_equal = bool(np.array_equal(got[..., 1].view(np.float32), want["distance"]))
    idx_diff = int(np.count_nonzero(got[..., 0] != want["index"]))
    return {"config": name, "note": note, "n_tree": len(tree_pts), "sdim": tree_pts.shape[1], "n_query": len(q), "k": k,
            "e": e, "resident_mqs": len(q) / ms / 1e3, "resident_ms": ms, "e2e_mqs": len(q) / e2e_s / 1e6,
            "e2e_ms": e2e_s * 1e3, "build_ms_device": info["build_ms"], "build_wall_s": build_wall,
            "tree_nodes": info["n_nodes"], "tree_height": info["height"],
            "cpu": {"kind": kind, "threads": threads, "sample": ns, "mqs_all_threads": ns / cpu_s / 1e6,
                    "mqs_one_thread": n1 / cpu1_s / 1e6},
            "parity": {"queries": ns, "distances_bit_equal": dist_equal, "index_mismatches": idx_diff}}


def cfg1():
    return knn_line("cfg1 uniform 100k/100k knn=1", D.uniform(100_000, 3, seed=1), D.uniform(100_000, 3, seed=2), 1,
                    cpu_sample=100_000, reps=50)


def cfg2s():
    tree_pts, q = D.bench_clouds()
    rng = np.random.default_rng(5)
    q = np.ascontiguousarray(q[rng.permutation(len(q))])
    return knn_line("cfg2 shuffled queries knn=1", tree_pts, q, 1, note="queries in random order on input")


def cfg3a():
    tree_pts, q = D.bench_clouds()
    return knn_line("cfg3 knn=16", tree_pts, q, 16, cpu_sample=len(q), reps=10)


def cfg3b():
    tree_pts, q = D.bench_clouds()
    tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
    qp = torch.from_numpy(q).pin_memory().numpy()
    r2 = 0.01
    box = {}

    def run():
        box["nns"] = tree.search_radius(qp, r2)
    e2e_s = timed(run, 3)
    nns = box["nns"]
    st = tree.last_stats
    hits = int(nns._offsets[-1])
    ref, kind = cpu_tree(tree_pts)
    threads = O.max_threads()
    ns = 1_000_000
    qs = np.ascontiguousarray(q[:ns])
    t0 = time.perf_counter()
    offs, flat = ref.search_radius(qs, r2)
    cpu_s = time.perf_counter() - t0
    same_counts = bool(np.array_equal(nns._offsets[:ns + 1], offs))
    a = np.sort(nns._flat[:int(offs[-1])], order=["distance", "index"]) if same_counts else None
    same_dist = bool(same_counts and np.array_equal(np.sort(nns._flat["distance"][:int(offs[-1])]),
                                                    np.sort(flat["distance"])))
    return {"config": "cfg3 search_radius r^2=0.01", "n_tree": len(tree_pts), "n_query": len(q), "radius": r2,
            "mean_hits_per_query": hits / len(q), "total_hits": hits, "e2e_mqs": len(q) / e2e_s / 1e6,
            "e2e_ms": e2e_s * 1e3, "device_ms": {"h2d": st.h2d_ms, "reorder": st.reorder_ms,
                                                  "count+scan+fill": st.kernel_ms, "d2h": st.d2h_ms},
            "cpu": {"kind": kind, "threads": 1, "sample": ns, "mqs_one_thread": ns / cpu_s / 1e6,
                    "note": "the reference's radius loop is serial in oracle/ref_driver.cpp"},
            "parity": {"queries": ns, "hit_counts_equal": same_counts, "sorted_distances_equal": same_dist}}


def cfgbox():
    """search_box on the cfg2 cloud: 1M axis-aligned boxes of 0.4 m edge centred on query points."""
    tree_pts, q = D.bench_clouds()
    tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
    nb = 1_000_000
    boxes = np.empty((2 * nb, 3), np.float32)
    boxes[0::2] = q[:nb] - np.float32(0.2)
    boxes[1::2] = q[:nb] + np.float32(0.2)
    box = {}

    def run():
        box["r"] = tree.search_box(boxes)
    e2e_s = timed(run, 3)
    res = box["r"]
    st = tree.last_stats
    ref, kind = cpu_tree(tree_pts)
    ns = 100_000
    t0 = time.perf_counter()
    offs, flat = ref.search_box(np.ascontiguousarray(boxes[0:2 * ns:2]), np.ascontiguousarray(boxes[1:2 * ns:2]))
    cpu_s = time.perf_counter() - t0
    same_counts = bool(np.array_equal(res._offsets[:ns + 1], offs))
    same_order = bool(same_counts and np.array_equal(res._flat[:int(offs[-1])], flat))
    return {"config": "search_box, 1M boxes of 0.4 m edge on the cfg2 cloud", "n_tree": len(tree_pts), "n_boxes": nb,
            "mean_hits_per_box": int(res._offsets[-1]) / nb, "e2e_mboxes_s": nb / e2e_s / 1e6, "e2e_ms": e2e_s * 1e3,
            "device_ms": {"h2d": st.h2d_ms, "count+scan+fill": st.kernel_ms, "d2h": st.d2h_ms},
            "cpu": {"kind": kind, "threads": 1, "sample": ns, "mboxes_s_one_thread": ns / cpu_s / 1e6},
            "parity": {"boxes": ns, "hit_counts_equal": same_counts, "indices_equal_in_dfs_order": same_order}}


def cfg4():
    pts = D.sift_shape(1_000_000, seed=1)
    q = D.sift_shape(10_000, seed=2)
    lines = []
    t0 = time.perf_counter()
    tree = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    build_wall = time.perf_counter() - t0
    info = tree.info()
    ref, kind = cpu_tree(pts)
    threads = O.max_threads()
    ns = 200
    for e in (0.0, 2.25):
        ms, res = resident_knn(tree, q, 10, e, reps=2)
        e2e_s = timed(lambda: tree.search_knn(q, 10, *([e] if e else [])), 1)
        qs = np.ascontiguousarray(q[:ns])
        t0 = time.perf_counter()
        want = ref.search_knn(qs, 10, e=e, threads=threads)
        cpu_s = time.perf_counter() - t0
        got = res.reshape(len(q), 10, 2)[:ns]
        lines.append({"config": "cfg4 sift-shape 1M x 128, 10k queries, knn=10" + (" approx e=2.25" if e else " exact"),
                      "n_tree": len(pts), "sdim": 128, "n_query": len(q), "k": 10, "e": e,
                      "resident_qps": len(q) / ms * 1e3, "resident_ms": ms, "e2e_qps": len(q) / e2e_s,
                      "build_ms_device": info["build_ms"], "build_wall_s": build_wall, "tree_height": info["height"],
                      "cpu": {"kind": kind, "threads": threads, "sample": ns, "qps_all_threads": ns / cpu_s},
                      "parity": {"queries": ns,
                                 "distances_bit_equal": bool(np.array_equal(got[..., 1].view(np.float32),
                                                                            want["distance"])),
                                 "index_mismatches": int(np.count_nonzero(got[..., 0] != want["index"]))}})
    return lines


def build():
    tree_pts, _ = D.bench_clouds()
    out = []
    for rule in (pt.kd_tree.Rule.SlidingMidpointMaxSide, pt.kd_tree.Rule.MidpointMaxSide,
                 pt.kd_tree.Rule.MedianMaxSide):
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10, rule=rule)
            wall = time.perf_counter() - t0
            info = tree.info()
            if best is None or info["build_ms"] < best["build_ms_device"]:
                best = {"config": "build 7.7M lidar-shape, max_leaf_size=10, " + rule.name,
                        "build_ms_device": info["build_ms"], "build_wall_s": wall, "tree_nodes": info["n_nodes"],
                        "tree_height": info["height"], "mpts_per_s_device": len(tree_pts) / info["build_ms"] / 1e3}
            del tree
        out.append(best)
    ref_t0 = time.perf_counter()
    cpu_tree(tree_pts)
    out.append({"config": "build 7.7M lidar-shape, reference on 1 host thread", "build_wall_s":
                time.perf_counter() - ref_t0})
    return out


def main():
    want = sys.argv[1:] or ["cfg1", "cfg2s", "cfg3a", "cfg3b", "cfgbox", "cfg4", "build"]
    fns = {"cfg1": cfg1, "cfg2s": cfg2s, "cfg3a": cfg3a, "cfg3b": cfg3b, "cfgbox": cfgbox, "cfg4": cfg4, "build": build}
    for w in want:
        r = fns[w]()
        for line in (r if isinstance(r, list) else [r]):
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

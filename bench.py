#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: Mqueries/s, knn=1, float x3, 7.7M-point tree / 7.2M queries.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (Z-order the batch, traverse, write neighbours) over one
batch of 7,200,863 synthetic LiDAR-shape queries against the resident 7,733,372-point tree
(configs[1]). N > 1: the tree is built on rank 0 and broadcast once over NVLink (NCCL through
torch.distributed), every rank searches its own 7.2M-query cloud (configs[4], weak scaling, no
collective on the query path).

  value     whole-job Mq/s with queries and results resident in HBM (CUDA events, max over ranks)
  e2e       the same through the public API with pinned HOST buffers (H2D + D2H inside the region)
  roofline  traversal kernel: algorithmic bytes (SURVEY.md §8d model, counters from the oracle's
            instrumented reference traversal) / its CUDA-event duration, vs the measured HBM peak
  cpu_baseline  the reference (oracle/_ref, unmodified headers) or the oracle port on the host cores
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mqueries/s knn=1 float x3 7.7M-tree/7.2M-query"
UNIT = "Mq/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-tree", type=int, default=None)
    ap.add_argument("--n-query", type=int, default=None)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--warp-per-query", action="store_true")
    ap.add_argument("--no-reorder", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip BASELINE.json's other configurations (bench_configs.py; N=1 only, adds ~2 minutes)")
    ap.add_argument("--configs", default=None,
                    help="comma list out of cfg1,cfg3_knn16,cfg3_radius,box,cfg4_exact,cfg4_approx,cfg4_forest (default: all)")
    return ap.parse_args()


def config_dict(n_tree, n_query, k, n_gpus, extra=None):
    cfg = {
        "workload": "configs[1]: lidar-shape synthetic (9 poses x 64-ring scanner in a 50 m room, scan order), "
                    "tree %d pts, %d queries/GPU, knn=%d, max_leaf_size=10, metric_l2_squared, "
                    "sliding_midpoint_max_side" % (n_tree, n_query, k),
        "n_tree": n_tree, "n_query_per_gpu": n_query, "k": k, "max_leaf_size": 10,
        "query_order": "scan order on input; Z-ordered on device inside the timed step (tile by tile once the tree has "
                       "measured that batches arrive locally coherent: pico_b200_tree_order_state)",
        "parallelism": "tree replicated (NCCL broadcast once), queries sharded x%d" % n_gpus,
        "l2": "inputs exceed L2 (tree 167 MB + queries 86 MB + results 58 MB vs 126 MB L2); no flush between steps",
    }
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------- reference arm
def cpu_reference_tree(tree_pts):
    from oracle import oracle as O
    if O.ref_available():
        return O.RefTree(tree_pts, 10), "reference"
    return O.OracleTree(tree_pts, 10), "port"


def run_reference(args, n_tree, n_query):
    """The reference's own CPU implementation, all host threads, OpenMP dynamic,128 like its
    Python binding (_pyco_tree/kd_tree.hpp:128); each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from pico_tree_b200 import datasets as D
    tree_pts, q = D.lidar_shape(n_tree, seed=1), D.lidar_shape(n_query, seed=2, pose_shift=0.35)
    t0 = time.perf_counter()
    tree, kind = cpu_reference_tree(tree_pts)
    build_s = time.perf_counter() - t0
    threads = O.max_threads()
    sample = min(n_query, 2_000_000)
    qs = np.ascontiguousarray(q[:sample])
    for _ in range(args.warmup):
        tree.search_knn(qs, args.k, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tree.search_knn(qs, args.k, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(n_tree, n_query, args.k, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "first %d of %d queries per step, all %d host threads, tree build %.2f s "
                                   "(1 thread) not included" % (sample, n_query, threads, build_s)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def kernel_sources_sha16():
    """Hash of the sources of the traversal kernels (what an ncu capture is valid for)."""
    import hashlib
    h = hashlib.sha256()
    for f in ("traverse.cuh", "search.cu", "common.cuh"):
        h.update(open(os.path.join(ROOT, "pico_tree_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken between the two time.time() stamps (all samples if fewer
        than three fall inside, e.g. for a very short timed region)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[5:9]]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except OSError:
            pass
        inside = [r for r in rows if t_begin is not None and t_begin <= r[0] <= t_end]
        window = "timed region"
        if len(inside) < 3:
            inside, window = rows, "warm-up + timed region"
        if inside:
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            reasons = sorted({n for r in inside for n, on in zip(names, r[3]) if on})
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=reasons, samples=len(inside), window=window)
        return out


def tree_leaf_scan(L, handle, q_dev, n_query, dev, repeats):
    """pico_b200_profile_leaf_scan: (neighbours, {descend_ms, scan_ms, scan_bytes}) per launch."""
    import torch
    from pico_tree_b200 import _lib
    out = torch.empty((n_query, 1, 2), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    d_ms, s_ms, nbytes = C.c_double(), C.c_double(), C.c_uint64()
    _lib.check(L.pico_b200_profile_leaf_scan(handle, C.c_void_p(q_dev.data_ptr()), n_query, 3,
                                             C.c_void_p(out.data_ptr()), repeats, C.byref(d_ms), C.byref(s_ms),
                                             C.byref(nbytes)))
    return out, {"descend_ms": d_ms.value, "scan_ms": s_ms.value, "scan_bytes": int(nbytes.value)}


def pcie_floor_ms(q_pin, out_pin, dev, reps=10, barrier=None):
    """One full-size H2D of the queries and D2H of the results, issued together on two streams (PCIe is
    full duplex): the time no host-buffer call can beat on this box. With `barrier` (N > 1) every repetition
    starts behind a barrier, so all ranks copy at the same time — what the shared uplinks and the one host memory
    system allow when every GPU moves its step's bytes at once."""
    import torch
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    d_in = torch.empty(q_pin.shape, dtype=q_pin.dtype, device=dev)
    d_out = torch.empty(out_pin.shape, dtype=out_pin.dtype, device=dev)
    scratch = torch.empty_like(out_pin).pin_memory()
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            d_in.copy_(q_pin, non_blocking=True)
        with torch.cuda.stream(s_out):
            scratch.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best


# --------------------------------------------------------------------------- our arm
def run_ours(args, n_tree, n_query):
    import torch
    import torch.distributed as dist

    import pico_tree_b200 as pt
    from pico_tree_b200 import _lib, datasets as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host buffers next to the GPU: first-touch puts the pinned pages of this rank on the GPU's NUMA node
    # (pico_tree_b200/hostmem.py); the CPU baseline further down gets all host threads back
    from pico_tree_b200 import hostmem
    cpus_all = os.sched_getaffinity(0)
    binding = hostmem.bind_to_gpu_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    k = args.k

    # ---- data: rank r searches its own cloud (cfg5: cfg2 queries + reseeded clouds)
    q_host = D.lidar_shape(n_query, seed=2 + rank, pose_shift=0.35)
    tree_pts = D.lidar_shape(n_tree, seed=1) if (rank == 0 or world == 1) else None

    # ---- tree: built on rank 0, broadcast once
    t_build0 = time.perf_counter()
    first_build = None
    if rank == 0:
        # the first build of a process also loads the build kernels' modules (lazy loading: 130 ms inside the
        # CUDA-event span of BENCH_r01); a small throw-away build takes that out of the reported build time
        t_first = time.perf_counter()
        first = pt.KdTree(tree_pts[:200_000], pt.Metric.L2Squared, 10, device=local)
        first_build = {"wall_s": time.perf_counter() - t_first, "build_ms_device": first.info()["build_ms"],
                       "n_tree": 200_000}
        del first
        # ... and the first full-size build grows the stream-ordered pool its workspaces come from; the reference's use
        # is a rebuild per frame, so the tree is built twice and the second build is the one reported
        t_build0 = time.perf_counter()
        tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10, device=local)
        first_build["first_full_size_build"] = {"wall_s": time.perf_counter() - t_build0,
                                                "build_ms_device": tree.info()["build_ms"]}
        del tree
        t_build0 = time.perf_counter()
        tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10, device=local)
        handle = tree._h
    build_wall = time.perf_counter() - t_build0
    bcast_ms = 0.0
    if world > 1:
        from pico_tree_b200 import distributed as pd
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        handle = pd.replicate_tree(tree if rank == 0 else None, 0, local)  # NCCL over NVLink / NVSwitch
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
    info = _lib.TreeInfo()
    _lib.check(L.pico_b200_tree_info_get(handle, C.byref(info)))
    # the same replication through the C-ABI's own entry (pico_b200_tree_broadcast with a plain ncclComm_t created
    # with NCCL's API): the replica it makes must answer exactly like the one above
    raw_nccl = None
    if world > 1:
        nccl_lib, comm = pd.raw_nccl_comm(local)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        h2 = pd.replicate_tree_raw_nccl(handle if rank == 0 else None, comm, 0, local)
        torch.cuda.synchronize()
        raw_ms = (time.perf_counter() - t0) * 1e3
        ns = min(n_query, 200_000)
        qs = np.ascontiguousarray(q_host[:ns])
        a = np.empty((ns, 2), dtype=np.dtype([("index", "<i4"), ("distance", "<f4")]))
        b = np.empty_like(a)
        for hh, o in ((handle, a), (h2, b)):
            _lib.check(L.pico_b200_knn(hh, C.c_void_p(qs.ctypes.data), ns, 3, 2, 0.0, C.c_void_p(o.ctypes.data), 0, None))
        same = torch.tensor([int(np.array_equal(a, b))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if rank != 0:
            L.pico_b200_tree_destroy(h2)
        nccl_lib.ncclCommDestroy(comm)
        raw_nccl = {"ms": raw_ms, "replica_answers_equal_on_all_ranks": bool(same.item()), "queries_checked": ns}

    # ---- resident inputs / outputs
    stream = torch.cuda.Stream(device=dev)
    q_dev = torch.from_numpy(q_host).to(dev)
    out_dev = torch.empty((n_query, k, 2), dtype=torch.int32, device=dev)  # {int32 index, f32 distance}
    flags = _lib.FLAG_DEVICE_POINTERS | _lib.FLAG_ASYNC
    if args.warp_per_query:
        flags |= _lib.FLAG_WARP_PER_QUERY
    if args.no_reorder:
        flags |= _lib.FLAG_NO_REORDER
    _lib.check(L.pico_b200_set_stream(C.c_void_p(stream.cuda_stream)))

    def step_resident():
        _lib.check(L.pico_b200_knn(handle, C.c_void_p(q_dev.data_ptr()), n_query, 3, k, 0.0,
                                   C.c_void_p(out_dev.data_ptr()), flags, None))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local) if rank == 0 else None
        for _ in range(args.warmup):
            step_resident()
        barrier()
        if sampler:
            time.sleep(0.3)  # let nvidia-smi come up before the timed region
        t_begin = time.time()
        _lib.check(L.pico_b200_profile_begin())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        trav_ms, trav_n = C.c_double(), C.c_uint64()
        _lib.check(L.pico_b200_profile_end(C.byref(trav_ms), C.byref(trav_n)))
        clocks = sampler.stop(t_begin, time.time()) if sampler else None
    ms_step = ms_total / args.steps
    kernel_ms = trav_ms.value / max(trav_n.value, 1)

    # ---- the leaf scan in isolation: first-leaf ranges, then a kernel that only streams them
    leaf_scan = None
    if k == 1 and not args.warp_per_query:
        _lib.check(L.pico_b200_set_stream(None))
        _, ls = tree_leaf_scan(L, handle, q_dev, n_query, dev, repeats=20)
        leaf_scan = ls

    # ---- e2e: public API, pinned host buffers, H2D + D2H inside the timed region
    q_pin = torch.from_numpy(q_host).pin_memory()
    out_pin = torch.empty((n_query, k, 2), dtype=torch.int32).pin_memory()
    host_flags = flags & ~(_lib.FLAG_DEVICE_POINTERS | _lib.FLAG_ASYNC)

    def step_e2e():
        _lib.check(L.pico_b200_knn(handle, C.c_void_p(q_pin.data_ptr()), n_query, 3, k, 0.0,
                                   C.c_void_p(out_pin.data_ptr()), host_flags, None))

    _lib.check(L.pico_b200_set_stream(None))  # library-owned streams: chunked H2D / traverse / D2H pipeline
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    floor_ms = pcie_floor_ms(q_pin, out_pin, dev)
    floor_concurrent_ms = None
    if world > 1:  # all ranks at once: the floor of the whole job's step
        barrier()
        floor_concurrent_ms = pcie_floor_ms(q_pin, out_pin, dev, reps=6, barrier=dist.barrier)
    os.sched_setaffinity(0, cpus_all)

    # resident and host paths must agree
    res_dev = out_dev.cpu().numpy()
    assert np.array_equal(res_dev, out_pin.numpy()), "resident and host-buffer paths disagree"

    # ---- max over ranks
    if world > 1:
        tmax = torch.tensor([ms_step, e2e_s, kernel_ms, floor_ms, floor_concurrent_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_step, e2e_s, kernel_ms, floor_ms, floor_concurrent_ms = [float(x) for x in tmax.tolist()]
    total_q = n_query * world
    value = total_q / (ms_step * 1e-3) / 1e6
    e2e_value = total_q / e2e_s / 1e6

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline + algorithmic-bytes counters (rank 0, N=1 only; bounded sample)
    cpu = None
    bytes_per_query = None
    counters = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        if world == 1:
            ref_tree, kind = cpu_reference_tree(tree_pts)
            threads = O.max_threads()
            t0 = time.perf_counter()
            ref_all = ref_tree.search_knn(q_host, k, threads=threads)
            all_s = time.perf_counter() - t0
            one_n = min(n_query, 500_000)
            t0 = time.perf_counter()
            ref_tree.search_knn(np.ascontiguousarray(q_host[:one_n]), k, threads=1)
            one_s = time.perf_counter() - t0
            cpu = {"value": n_query / all_s / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": "all %d queries once on %d threads (OpenMP dynamic,128); single thread: first %d "
                             "queries -> %.3f Mq/s" % (n_query, threads, one_n, one_n / one_s / 1e6),
                   "single_thread_value": one_n / one_s / 1e6}
            # parity of the benchmarked run itself against the CPU reference (full size)
            got = res_dev.reshape(n_query, k, 2)
            same_d = np.array_equal(got[..., 1].view(np.float32), ref_all["distance"])
            idx_diff = int(np.count_nonzero(got[..., 0] != ref_all["index"]))
            cpu["parity_vs_this_run"] = {"distances_bit_equal": bool(same_d), "index_mismatches": idx_diff,
                                         "queries": n_query}
        # algorithmic bytes: counters from the oracle's instrumented reference traversal (sample of
        # rank 0's queries; every N, the roofline describes the kernel, not the host)
        oc = O.OracleTree(tree_pts, 10)
        cs = min(n_query, 400_000)
        sel = np.ascontiguousarray(q_host[:: max(n_query // cs, 1)][:cs])
        _, cnt = oc.search_knn(sel, k, counters=True, threads=O.max_threads())
        counters = (cnt / len(sel)).tolist()
        bytes_per_query = 4 * 3 + 8 * k + counters[0] * 16 + counters[2] * 16

    peak = peaks.get("hbm_gbs")
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)"
    if peak is None:
        peak, peak_kind = 6650.0, "fallback (B200_PROFILING.md)"
    roofline = None
    if bytes_per_query is not None and kernel_ms > 0:
        achieved = bytes_per_query * n_query / (kernel_ms * 1e-3) / 1e9
        # DRAM bytes per launch from the last ncu --set full capture of this kernel — only if it was taken with the
        # kernel sources that are being timed now (profiles/traffic.json records their hash)
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("sources_sha16") == kernel_sources_sha16() and k == 1:
                traffic = tj.get("knn1_dram_bytes_per_launch")
                traffic_note = tj.get("source")
            else:
                traffic_note = "profiles/traffic.json was captured for other kernel sources or another k: not reported"
        except (OSError, ValueError):
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_note": traffic_note,
                    "kernel": ("knn_thread_kernel<float,3,1,FAST>" if k == 1 and os.environ.get("PICO_B200_NN", "0") == "0"
                               else "nn_kernel<float,3,...> (PICO_B200_NN)" if k == 1 else
                               "knn_thread_kernel<float,3,K,FAST>") if not args.warp_per_query
                    else "knn_warp_kernel<float,PACKED,REG>", "kernel_ms": kernel_ms,
                    "algorithmic_bytes_per_query": bytes_per_query,
                    "per_query_branches_leaves_points": counters, "peak_source": peak_kind}
    if leaf_scan is not None:
        # same peak; bytes = what one launch of the isolated kernel reads and writes (queries, slot ids,
        # leaf ranges, the leaves' float4 records, results)
        gbs = leaf_scan["scan_bytes"] / (leaf_scan["scan_ms"] * 1e-3) / 1e9
        ls_traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("sources_sha16") == kernel_sources_sha16():
                ls_traffic = tj.get("leaf_scan_dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        leaf_scan = dict(leaf_scan, kernel="leaf_scan_kernel<float,3>", achieved=gbs, peak=peak, unit="GB/s",
                         frac=gbs / peak, traffic=ls_traffic, note="first leaf of every query, Z-ordered slots; ranges written by "
                         "first_leaf_kernel (descend_ms)")
        if roofline is not None:
            roofline["leaf_scan"] = leaf_scan
        else:
            roofline = {"bound": "hbm", "leaf_scan": leaf_scan}

    # ---- BASELINE.json's other configurations (N=1 only): each with value, e2e, parity vs the reference, roofline
    configs = None
    if world == 1 and not args.no_configs and not args.no_cpu_baseline and k == 1 and not args.warp_per_query:
        import bench_configs
        from oracle import oracle as O
        ctx = {"torch": torch, "lib": _lib, "oracle": O, "pt": pt, "peak": peak, "peak_kind": peak_kind}
        configs = bench_configs.run_all(ctx, tree, tree_pts, q_host, args.configs.split(",") if args.configs else None)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(n_tree, n_query, k, world, {
            "tree_nodes": int(info.n_nodes), "tree_height": int(info.height), "build_ms_device": info.build_ms,
            "build_wall_s": build_wall, "first_build_of_the_process": first_build, "tree_broadcast_ms": bcast_ms,
            "tree_broadcast_raw_nccl": raw_nccl, "tree_device_bytes": int(info.device_bytes)}),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(q_host.nbytes),
                "d2h_bytes_per_step": int(n_query * k * 8), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "pcie_floor_ms": floor_ms, "pcie_floor_all_ranks_at_once_ms": floor_concurrent_ms,
                "e2e_over_concurrent_floor": (floor_concurrent_ms / (e2e_s * 1e3)) if floor_concurrent_ms else None,
                "host_binding": binding,
                "pcie_floor_note": "one H2D of all queries + one D2H of all results issued together, best of 10"},
        "gpu_launches": int(args.steps * (1 + (0 if args.no_reorder else 1))),
        "gpu_launches_note": "own kernels per step: the ordering kernel (tile_order_kernel once the tree has seen that "
                             "batches arrive locally coherent, else morton_kernel + CUB's radix-sort passes, which are "
                             "not counted) + the traversal kernel",
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "configs": configs,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    from pico_tree_b200 import datasets as D
    n_tree = args.n_tree or D.N_TREE
    n_query = args.n_query or D.N_QUERY
    if args.impl == "reference":
        run_reference(args, n_tree, n_query)
    else:
        run_ours(args, n_tree, n_query)


if __name__ == "__main__":
    main()

"""Resident cfg2 knn=1 step and traversal-kernel time for the tuning hooks of the search-image nn kernel
(PICO_B200_NN, PICO_B200_FAT_LEAF, ...), each point in its own process (the hooks are read once), with the
full-size parity check against the unmodified reference (oracle/_ref) done once and every point compared with it.

    python profiles/nn_sweep.py [point ...]      point = KEY=VAL,KEY=VAL (PICO_B200_ prefix implied)
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/pico_b200_bench_clouds.npz"
WANT = "/tmp/pico_b200_bench_want.npy"

ONE = r"""
import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, %r)
import pico_tree_b200 as pt
from pico_tree_b200 import _lib
z = np.load(%r)
tree_pts, q = z["tree"], z["q"]
k = int(os.environ.get("SWEEP_K", "1"))
tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
L = _lib.lib()
qd = torch.from_numpy(q).cuda()
od = torch.empty((len(q), k, 2), dtype=torch.int32, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): tree.search_knn_device(qd, k, nns=od)
    torch.cuda.synchronize()
    _lib.check(L.pico_b200_profile_begin())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    reps = 30
    for _ in range(reps): tree.search_knn_device(qd, k, nns=od)
    e1.record(s); torch.cuda.synchronize()
    ms, n = C.c_double(), C.c_uint64()
    _lib.check(L.pico_b200_profile_end(C.byref(ms), C.byref(n)))
got = od.cpu().numpy()
note = ""
if k == 1 and os.path.exists(%r):
    want = np.load(%r)
    bad_i = int(np.count_nonzero(got[:, 0, 0] != want["index"][:, 0]))
    bad_d = int(np.count_nonzero(got[:, 0, 1].view(np.float32) != want["distance"][:, 0]))
    note = " | vs reference: %%d index / %%d distance mismatches of %%d" %% (bad_i, bad_d, len(q))
print("step %%.3f ms | traversal kernel(s) %%.3f ms (%%d spans) | build %%.2f ms%%s" %% (
    e0.elapsed_time(e1) / reps, ms.value / reps, n.value // reps, tree.info()["build_ms"], note))
""" % (ROOT, CACHE, WANT, WANT)


def run(code, env):
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env))
    return (r.stdout.strip() + ("\n" + r.stderr.strip()[-2600:] if r.stderr.strip() else "")).strip()


if __name__ == "__main__":
    from pico_tree_b200 import datasets as D
    if not os.path.exists(CACHE):
        tree_pts, q = D.bench_clouds()
        np.savez(CACHE, tree=tree_pts, q=q)
    else:
        z = np.load(CACHE)
        tree_pts, q = z["tree"], z["q"]
    if not os.path.exists(WANT):
        from oracle import oracle as O
        ref = O.RefTree(tree_pts, 10) if O.ref_available() else O.OracleTree(tree_pts, 10)
        np.save(WANT, ref.search_knn(q, 1, threads=O.max_threads()))
    points = sys.argv[1:] or ["NN=0", "NN=1", "NN=5", "FAT_LEAF=12,NN=1", "FAT_LEAF=12,NN=3", "FAT_LEAF=16,NN=1", "FAT_LEAF=16,NN=3", "FAT_LEAF=24,NN=3"]
    for p in points:
        env = {}
        for kv in p.split(","):
            if kv:
                key, val = kv.split("=")
                env[key if key.startswith("SWEEP_") else "PICO_B200_" + key] = val
        print("%-40s %s" % (p, run(ONE, env)), flush=True)

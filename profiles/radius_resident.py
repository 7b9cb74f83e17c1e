"""Resident search_radius (device pointers in, ragged result left in HBM), call by call: the 6.1 GB hit array comes
from the stream-ordered pool every call. PICO_B200_TREE_POOL=0 puts the tree's arrays back on cudaMalloc."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pico_tree_b200 as pt
from pico_tree_b200 import _lib, datasets as D

tree_pts, q = D.bench_clouds()
tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
L = _lib.lib()
dev = torch.device("cuda", 0)
qd = torch.from_numpy(q).to(dev)
offs = torch.empty(len(q) + 1, dtype=torch.int64, device=dev)
times = []
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hits = C.c_void_p()
    _lib.check(L.pico_b200_radius(tree._h, C.c_void_p(qd.data_ptr()), len(q), 3, 0.01, 0.0, C.c_void_p(offs.data_ptr()),
                                  C.byref(hits), _lib.FLAG_DEVICE_POINTERS, None))
    t1 = time.perf_counter()
    L.pico_b200_free_device(hits)
    torch.cuda.synchronize()
    times.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
print("TREE_POOL=%s | ms per call (search, free): %s" % (os.environ.get("PICO_B200_TREE_POOL", "-"),
      " ".join("%.1f+%.1f" % t for t in times)), flush=True)

"""Workload for ncu captures: cfg2 tree, then knn=1, knn=16, radius r^2=0.01 and the isolated leaf scan
("leaf": first_leaf_kernel + leaf_scan_kernel) with device-resident queries (see profiles/README.md for the
command lines)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pico_tree_b200 as pt  # noqa: E402
from pico_tree_b200 import _lib, datasets as D  # noqa: E402

which = sys.argv[1:] or ["knn1", "knn16", "radius"]
tree_pts, q = D.bench_clouds()
tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
L = _lib.lib()
dev = torch.device("cuda", 0)
qd = torch.from_numpy(q).to(dev)
for w in which:
    if w.startswith("knn"):
        k = int(w[3:])
        out = torch.empty((len(q), k, 2), dtype=torch.int32, device=dev)
        for _ in range(2):
            _lib.check(L.pico_b200_knn(tree._h, C.c_void_p(qd.data_ptr()), len(q), 3, k, 0.0,
                                       C.c_void_p(out.data_ptr()), _lib.FLAG_DEVICE_POINTERS, None))
    elif w == "leaf":
        tree.profile_leaf_scan(qd, repeats=2)
    elif w == "radius":
        n = 2_000_000
        offs = torch.empty(n + 1, dtype=torch.int64, device=dev)
        hits = C.c_void_p()
        _lib.check(L.pico_b200_radius(tree._h, C.c_void_p(qd.data_ptr()), n, 3, 0.01, 0.0,
                                      C.c_void_p(offs.data_ptr()), C.byref(hits), _lib.FLAG_DEVICE_POINTERS, None))
        L.pico_b200_free_device(hits)
torch.cuda.synchronize()
print("done", which)

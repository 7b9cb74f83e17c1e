"""pico_b200_knn on cfg2, knn=1, from PAGEABLE host memory (what a std::vector / numpy caller has): ms per call
after the pinned mirrors exist, and with PICO_B200_TIMELINE=1 PICO_B200_HOST_STREAMS=12 the per-chunk timeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D

tree_pts, q = D.bench_clouds()
tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
out = np.empty((len(q), 1), dtype=tree.dtype_neighbor)
out[:] = 0
for _ in range(3):
    tree.search_knn(q, 1, out)
times = []
for _ in range(10):
    t0 = time.perf_counter()
    tree.search_knn(q, 1, out)
    times.append((time.perf_counter() - t0) * 1e3)
print("pageable e2e ms per call: best %.2f median %.2f | %s" % (min(times), sorted(times)[5], " ".join("%.2f" % t for t in times)), flush=True)

import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D
tree_pts, q = D.bench_clouds()
q = np.ascontiguousarray(q[:2_000_000])
t = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
for k in (20, 32):
    for kw in ({}, {"warp_per_query": True}):
        r = t.search_knn(q, k, **kw)
        r = t.search_knn(q, k, **kw)
        print("k", k, kw, "kernel", round(t.last_stats.kernel_ms, 2), "ms ->", round(len(q) / t.last_stats.kernel_ms / 1e3), "Mq/s", flush=True)
    a = t.search_knn(q[:300000], k); b = t.search_knn(q[:300000], k, warp_per_query=True)
    print("  equal:", np.array_equal(a["index"], b["index"]), np.array_equal(a["distance"], b["distance"]))

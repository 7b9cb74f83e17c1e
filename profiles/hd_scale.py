import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D
pts = D.sift_shape(1_000_000, seed=1)
tree = pt.KdTree(pts, pt.Metric.L2Squared, 10)
print(tree.info())
for nq in (32, 256, 1184, 2368, 4736, 9472):
    q = D.sift_shape(nq, seed=2)
    tree.search_knn(q, 10, reorder=False)
    print(nq, "queries:", round(tree.last_stats.kernel_ms, 1), "ms  ->", round(nq / tree.last_stats.kernel_ms * 1e3), "q/s", flush=True)

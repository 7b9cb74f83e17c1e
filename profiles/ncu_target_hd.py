"""Workload for ncu: sift-shape 1M x 128 tree, knn=10 for 2048 queries (warp-per-query kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pico_tree_b200 as pt  # noqa: E402
from pico_tree_b200 import datasets as D  # noqa: E402

pts = D.sift_shape(1_000_000, seed=1)
q = D.sift_shape(int(sys.argv[1]) if len(sys.argv) > 1 else 2048, seed=2)
tree = pt.KdTree(pts, pt.Metric.L2Squared, 10)
print(tree.info())
r = tree.search_knn(q, 10)
print(tree.last_stats.kernel_ms, "ms")

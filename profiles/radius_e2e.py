"""End-to-end search_radius (r^2 = 0.01, 760 M neighbours = 6.1 GB of results) and search_box from pinned host
queries, call by call: the device part is ~25 ms, the rest is moving the ragged result into host memory.
Knobs: PICO_B200_COPY_THREADS, PICO_B200_RESULT_CACHE_MB (search.cu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D

tree_pts, q = D.bench_clouds()
tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
qp = torch.from_numpy(q).pin_memory().numpy()
times = []
nns = None
for rep in range(5):
    t0 = time.perf_counter()
    nns = tree.search_radius(qp, 0.01)
    times.append((time.perf_counter() - t0) * 1e3)
total = int(nns._offsets[-1])
print("COPY_THREADS=%s RESULT_CACHE_MB=%s | radius: %d neighbours, %.2f GB | ms per call: %s" % (
    os.environ.get("PICO_B200_COPY_THREADS", "-"), os.environ.get("PICO_B200_RESULT_CACHE_MB", "-"), total,
    total * 8 / 1e9, " ".join("%.0f" % t for t in times)), flush=True)

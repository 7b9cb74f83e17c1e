"""Device build time per splitter rule at the headline size (7.7 M LiDAR-shaped points), warmed: the time the
C-ABI reports (CUDA events around the kernels and the per-level round trips, build.cu) and the wall clock of the
whole constructor. Usage: python profiles/build_rules.py [n_points]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 7_700_000
pts = D.lidar_shape(n, seed=1)
pt.KdTree(pts[:100_000], pt.Metric.L2Squared, 10)  # module load, allocator
for name, rule in (("sliding_midpoint", pt.kd_tree.Rule.SlidingMidpointMaxSide),
                   ("midpoint", pt.kd_tree.Rule.MidpointMaxSide), ("median", pt.kd_tree.Rule.MedianMaxSide)):
    runs = []
    t = None
    for _ in range(4):
        del t  # the previous tree goes first: its cudaFree calls are not part of the next build
        t0 = time.perf_counter()
        t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=rule)
        wall = (time.perf_counter() - t0) * 1e3
        inf = t.info()
        runs.append((inf["build_ms"], wall))
    best = min(runs)
    print(f"{name:18s} n={n} device {best[0]:8.2f} ms | constructor wall {best[1]:8.2f} ms | height {inf['height']} | all runs "
          + " ".join(f"{a:.1f}" for a, _ in runs), flush=True)

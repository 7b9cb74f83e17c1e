"""Per-level host timeline of warmed builds (PICO_B200_BUILD_TIMELINE, build.cu). Usage: build_timeline.py [rule]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D

rule = {"median": pt.kd_tree.Rule.MedianMaxSide, "sliding": pt.kd_tree.Rule.SlidingMidpointMaxSide}[sys.argv[1] if len(sys.argv) > 1 else "median"]
pts = D.lidar_shape(7_700_000, seed=1)
pt.KdTree(pts[:100_000], pt.Metric.L2Squared, 10, rule=rule)
for rep in range(6):
    if rep >= 2:
        os.environ["PICO_B200_BUILD_TIMELINE"] = "1"
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=rule)
    print("build_ms", t.info()["build_ms"], file=sys.stderr, flush=True)
    del t

"""One warmed median-rule build of 7.7 M LiDAR-shaped points (target of the ncu launch list
profiles/r2/launches_build_median.csv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D

pts = D.lidar_shape(7_700_000, seed=1)
pt.KdTree(pts[:100_000], pt.Metric.L2Squared, 10, rule=pt.kd_tree.Rule.MedianMaxSide)
t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=pt.kd_tree.Rule.MedianMaxSide)
print(t.info())

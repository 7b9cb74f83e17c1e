import sys
sys.path.insert(0, "/root/repo")
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D
pts = D.lidar_shape(D.N_TREE, seed=1)
t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
print(t.info())

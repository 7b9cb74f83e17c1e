"""compute-sanitizer target for the build paths added late in round 2: the median rule's multi-CTA introselect
(median_huge_level: cooperative launch, barriers in global memory, exchanges between CTAs), the one-kernel root
set-up and the pool-backed tree arrays; every permutation is checked against the oracle.

    compute-sanitizer --tool memcheck|racecheck python profiles/sanitize_build_target.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PICO_B200_HUGE_MIN", "2048")

import pico_tree_b200 as pt  # noqa: E402
from oracle import oracle as O  # noqa: E402
from pico_tree_b200 import datasets as D  # noqa: E402

rng = np.random.default_rng(5)
cases = [D.lidar_shape(50_000, seed=1), rng.random((30_000, 2)), D.sift_shape(12_000, 16, seed=4)]
dup = D.lidar_shape(40_000, seed=2)
dup[::7] = dup[3]
cases.append(dup)
for pts in cases:
    for rule, oname in ((pt.kd_tree.Rule.MedianMaxSide, "median"), (pt.kd_tree.Rule.SlidingMidpointMaxSide, "sliding_midpoint")):
        t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=rule)
        o = O.OracleTree(pts, 10, rule=oname)
        _, idx, _ = t.export()
        assert np.array_equal(idx, o.indices), (pts.shape, oname)
        got, want = t.search_knn(pts[:2000], 3), o.search_knn(pts[:2000], 3)
        assert np.array_equal(got["index"], want["index"])
data = np.load(os.path.join(ROOT, "tests", "data", "introselect_killers.npz"))
for key in data.files:
    v = data[key]
    pts = np.ascontiguousarray(np.stack([v, np.zeros_like(v)], axis=1))
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=pt.kd_tree.Rule.MedianMaxSide)
    assert np.array_equal(t.export()[1], O.OracleTree(pts, 10, rule="median").indices), key
print("build paths ok")

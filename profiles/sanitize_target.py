"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every build path (warp, CTA
and — with PICO_B200_HUGE_MIN lowered — the grid-wide chunked passes with their in-place exchanges), the three
rules, both traversal families, k in registers and in memory, radius two-pass (count, scan, fill, sort), box, the
topological metrics, float64, the kd_forest build and search, the pinned host pipeline and (de)serialisation.
Results are checked against the oracle so that a silent corruption cannot hide behind a clean sanitizer log.

    compute-sanitizer --tool memcheck python profiles/sanitize_target.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PICO_B200_HUGE_MIN", "4096")  # nodes above 4096 points take the grid-wide path

import pico_tree_b200 as pt  # noqa: E402
from oracle import oracle as O  # noqa: E402
from pico_tree_b200 import datasets as D  # noqa: E402

rng = np.random.default_rng(3)
n = int(os.environ.get("SANITIZE_N", "60000"))
pts = D.lidar_shape(n, seed=1)
q = D.lidar_shape(n // 4, seed=2, pose_shift=0.35)
for rule, oname in ((pt.kd_tree.Rule.SlidingMidpointMaxSide, "sliding_midpoint"), (pt.kd_tree.Rule.MidpointMaxSide, "midpoint"),
                    (pt.kd_tree.Rule.MedianMaxSide, "median")):
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=rule)
    o = O.OracleTree(pts, 10, rule=oname)
    _, idx, _ = t.export()
    assert np.array_equal(idx, o.indices), rule
    for k in (1, 8, 40):
        for kw in ({}, {"warp_per_query": True}):
            got, want = t.search_knn(q, k, **kw), o.search_knn(q, k)
            assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["distance"], want["distance"])
    assert np.array_equal(t.search_knn(q, 4, 1.5)["index"], o.search_knn(q, 4, e=1.5)["index"])
    for kw in ({}, {"warp_per_query": True}):
        r = t.search_radius(q, 0.05, **kw)
        offs, flat = o.search_radius(q, 0.05)
        assert np.array_equal(r._offsets, offs) and np.array_equal(r._flat["index"][:len(flat)], flat["index"])
    r = t.search_radius(q, 0.05, True)
    boxes = np.empty((2000, 3), np.float32)
    boxes[0::2], boxes[1::2] = q[:1000] - 0.4, q[:1000] + 0.4
    b = t.search_box(boxes)
    offs, flat = o.search_box(boxes[0::2], boxes[1::2])
    assert np.array_equal(b._offsets, offs) and np.array_equal(b._flat[:len(flat)], flat)
    blob = t.serialize()
print("euclidean f32 ok")
# float64, other metrics, 2-D, high-D rows, topological
p64 = rng.random((20000, 3))
t = pt.KdTree(p64, pt.Metric.L1, 7)
o = O.OracleTree(p64, 7, metric="l1")
assert np.array_equal(t.search_knn(p64[:3000], 5)["index"], o.search_knn(p64[:3000], 5)["index"])
p16 = rng.random((20000, 16), dtype=np.float32)
t = pt.KdTree(p16, pt.Metric.L2Squared, 10)
o = O.OracleTree(p16, 10)
assert np.array_equal(t.search_knn(p16[:2000], 10)["index"], o.search_knn(p16[:2000], 10)["index"])
assert np.array_equal(t.search_knn(p16[:500], 50)["index"], o.search_knn(p16[:500], 50)["index"])
se2 = rng.random((20000, 3), dtype=np.float32)
t = pt.KdTree(se2, pt.Metric.SE2Squared, 10)
o = O.OracleTree(se2, 10, metric="se2_squared")
assert np.array_equal(t.search_knn(se2[:3000], 4)["index"], o.search_knn(se2[:3000], 4)["index"])
print("f64 / rows / topological ok")
# kd_forest
rot = rng.normal(size=(3, 16))
rot = (rot / np.linalg.norm(rot, axis=1, keepdims=True)).astype(np.float32)
f = pt.KdForest(p16, 10, 3, rotations=rot)
of = O.OracleForest(p16, rot, 10)
for k, leaves in ((1, 8), (10, 40), (40, 900)):
    got, want = f.search_knn(p16[:1500], k, leaves), of.search_knn(p16[:1500], k, leaves)
    assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["distance"], want["distance"])
print("forest ok")
# pinned host pipeline (chunked, copy-ahead, high-priority ordering): needs >= 2 Mi queries
import torch  # noqa: E402
big_q = np.ascontiguousarray(np.tile(q, (2 * (1 << 20) // len(q) + 2, 1))[: 2 * (1 << 20) + 12345])
t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
qp = torch.from_numpy(big_q).pin_memory().numpy()
out = torch.empty((len(big_q), 1, 2), dtype=torch.int32).pin_memory().numpy().view(t.dtype_neighbor).reshape(len(big_q), 1)
t.search_knn(qp, 1, out)
want = O.OracleTree(pts, 10).search_knn(q, 1)
assert np.array_equal(out["index"][:len(q)], want["index"])
print("host pipeline ok")

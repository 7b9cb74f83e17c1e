"""Parity of the CUDA path (through the C-ABI / Python host mirror) against the CPU oracle and
against the committed reference fixtures. Everything here needs a B200: `-m gpu`."""
import glob
import os

import numpy as np
import pytest

from parity import (assert_knn_parity, assert_radius_parity, assert_same_structure, leaf_sets_equal,
                    nodes_from_export, split_ragged)

pytestmark = pytest.mark.gpu

METRIC = {"l2_squared": "L2Squared", "l1": "L1", "lpinf": "LPInf", "lninf": "LNInf", "so2": "SO2",
          "se2_squared": "SE2Squared"}
RULE = {"sliding_midpoint": "SlidingMidpointMaxSide", "midpoint": "MidpointMaxSide", "median": "MedianMaxSide"}


def make_tree(pt, pts, metric="l2_squared", rule="sliding_midpoint", stop="max_leaf_size", stop_value=10, bounds=None):
    kw = {"rule": pt.kd_tree.Rule[RULE[rule]], "bounds": bounds}
    if stop == "max_leaf_depth":
        kw["max_leaf_depth"] = stop_value
        return pt.KdTree(pts, pt.Metric[METRIC[metric]], **kw)
    return pt.KdTree(pts, pt.Metric[METRIC[metric]], stop_value, **kw)


@pytest.fixture(scope="module")
def pt():
    import pico_tree_b200
    return pico_tree_b200


def _golden_files():
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(glob.glob(os.path.join(d, "*.npz")))


# ---------------------------------------------------------------- reference fixtures
@pytest.mark.parametrize("path", _golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_fixture_build_and_search(pt, path):
    """Device build + every search kind against outputs of the unmodified reference."""
    g = np.load(path)
    pts, q = g["pts"], g["q"]
    metric, rule, stop = str(g["metric"]), str(g["rule"]), str(g["stop"])
    t = make_tree(pt, pts, metric, rule, stop, int(g["stop_value"]))
    nodes, indices, box = t.export()
    assert np.array_equal(box, g["root_box"])
    got = nodes_from_export(nodes, pts.dtype)
    want = {f: g["node_" + f] for f in ("left_max", "right_min", "split_dim", "begin", "end", "left", "right")}
    # node for node (split positions, split dims, tight bounds, leaf ranges, links) AND index for index:
    # std::partition's exchange order and libstdc++'s nth_element (slides, median rule) are reproduced
    # exactly by the device build, duplicated coordinates included
    assert_same_structure(got, want)
    assert np.array_equal(indices, g["indices"]), "index permutation differs from the reference's"
    if pts.dtype == np.float32:  # (the f64 stream has uninitialised padding bytes in the reference)
        assert t._saved_stream() == g["saved_stream"].tobytes(), "kd_tree::save stream differs from the reference's"
    assert sorted(indices.tolist()) == list(range(len(pts)))
    if "node_outer" in g:  # topological metrics: the two extra bounds of kd_tree_branch_double
        assert np.array_equal(t.export_outer_bounds(), g["node_outer"])

    # identical trees -> identical answers, ties, approximate search and visit order included
    k = g["knn_index"].shape[1]
    for kw in ({}, {"warp_per_query": True}):
        for name, kk, e in (("nn", 1, 0.0), ("knn", k, 0.0), ("aknn", k, 1.5)):
            r = t.search_knn(q, kk, **kw) if e == 0.0 else t.search_knn(q, kk, e, **kw)
            assert np.array_equal(r["index"], g[name + "_index"]), (name, kw)
            assert np.array_equal(r["distance"], g[name + "_distance"]), (name, kw)
    nns = t.search_radius(q, float(g["radius"]))
    assert np.array_equal(nns._offsets, g["radius_offsets"])
    n = len(g["radius_index"])
    assert np.array_equal(nns._flat["index"][:n], g["radius_index"])
    assert np.array_equal(nns._flat["distance"][:n], g["radius_distance"])
    boxes = np.empty((2 * len(q), pts.shape[1]), pts.dtype)
    boxes[0::2], boxes[1::2] = g["box_min"], g["box_max"]
    res = t.search_box(boxes)
    assert np.array_equal(res._offsets, g["box_offsets"])
    assert np.array_equal(res._flat[:len(g["box_index"])], g["box_index"])  # DFS order


@pytest.mark.parametrize("path", _golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_fixture_reference_tree_uploaded(pt, oracle, path, tmp_path):
    """The reference's own saved stream loads into the engine (kd_tree::load path) and then every
    result — including approximate search and visit ORDER — must be identical, ties included."""
    g = np.load(path)
    pts, q = g["pts"], g["q"]
    metric = str(g["metric"])
    header = b"\x89PKD" + (1).to_bytes(4, "little") + len(METRIC[metric]).to_bytes(8, "little") + METRIC[metric].encode()
    f = tmp_path / "ref.pkd"
    f.write_bytes(header + g["saved_stream"].tobytes())
    t = pt.load_kd_tree(pts, str(f))
    _, indices, box = t.export()
    assert np.array_equal(indices, g["indices"]) and np.array_equal(box, g["root_box"])
    k = g["knn_index"].shape[1]
    for kw in ({}, {"warp_per_query": True}, {"reorder": False}):
        for name, kk, e in (("nn", 1, 0.0), ("knn", k, 0.0), ("aknn", k, 1.5)):
            r = t.search_knn(q, kk, **kw) if e == 0.0 else t.search_knn(q, kk, e, **kw)
            assert np.array_equal(r["index"], g[name + "_index"]), (name, kw)
            assert np.array_equal(r["distance"], g[name + "_distance"]), (name, kw)
    for kw in ({}, {"warp_per_query": True}):
        nns = t.search_radius(q, float(g["radius"]), **kw)
        assert np.array_equal(nns._offsets, g["radius_offsets"])
        n = len(g["radius_index"])
        assert np.array_equal(nns._flat["index"][:n], g["radius_index"])
        assert np.array_equal(nns._flat["distance"][:n], g["radius_distance"])
        nns = t.search_radius(q, float(g["radius"]), 1.5, True, **kw)
        assert np.array_equal(nns._offsets, g["aradius_offsets"])
        assert np.array_equal(nns._flat["distance"][:len(g["aradius_distance"])], g["aradius_distance"])
    boxes = np.empty((2 * len(q), pts.shape[1]), pts.dtype)
    boxes[0::2], boxes[1::2] = g["box_min"], g["box_max"]
    res = t.search_box(boxes)
    assert np.array_equal(res._offsets, g["box_offsets"])
    assert np.array_equal(res._flat[:len(g["box_index"])], g["box_index"])  # DFS order
    # and the engine writes the same bytes back
    out = tmp_path / "mine.pkd"
    pt.save_kd_tree(t, str(out))
    if pts.dtype == np.float32:
        assert out.read_bytes() == f.read_bytes()
    else:
        # the reference writes its raw branch struct {int; double; double}: 4 padding bytes per
        # branch are uninitialised there, so compare the decoded content instead
        topo = metric in oracle.TOPOLOGICAL
        a = oracle.parse_saved_tree(out.read_bytes()[len(header):], pts.dtype, topo)
        b = oracle.parse_saved_tree(f.read_bytes()[len(header):], pts.dtype, topo)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        for fld in a[3].dtype.names:
            assert np.array_equal(a[3][fld], b[3][fld])
        if topo:
            assert np.array_equal(a[4], b[4])
        t2 = pt.load_kd_tree(pts, str(out))
        assert np.array_equal(t2.search_knn(q, k)["index"], g["knn_index"])


# ---------------------------------------------------------------- oracle on seeded inputs
@pytest.mark.parametrize("n,sdim,leaf", [(100_000, 3, 10), (50_000, 2, 1), (30_000, 3, 64), (20_000, 1, 7)])
def test_uniform_vs_oracle(pt, oracle, n, sdim, leaf):
    """cfg1-like: uniform clouds, build + knn 1/4/16 + radius + box against the oracle."""
    from pico_tree_b200 import datasets as D
    pts = D.uniform(n, sdim, seed=1)
    q = D.uniform(n // 2, sdim, seed=2)
    o = oracle.OracleTree(pts, leaf)
    t = pt.KdTree(pts, pt.Metric.L2Squared, leaf)
    nodes, indices, box = t.export()
    on = o.nodes
    assert_same_structure(nodes_from_export(nodes, pts.dtype), on)
    assert np.array_equal(box, o.root_box)
    assert np.array_equal(indices, o.indices)
    info = t.info()
    assert info["n_nodes"] == o.num_nodes and info["height"] == o.height
    ties = 0
    for k in (1, 4, 16, 40):
        ties += assert_knn_parity(t.search_knn(q, k), o.search_knn(q, k), pts, q)
    assert ties == 0
    r = 0.0004 if sdim > 1 else 1e-8
    nns = t.search_radius(q, r)
    offs, flat = o.search_radius(q, r)
    assert_radius_parity(nns._offsets, nns._flat[:len(flat)], offs, flat, ordered=False)
    nns = t.search_radius(q, r, True)
    offs, flat = o.search_radius(q, r, sort=True)
    assert np.array_equal(nns._offsets, offs) and np.array_equal(nns._flat["distance"][:len(flat)], flat["distance"])
    boxes = np.empty((2000, sdim), np.float32)
    boxes[0::2] = q[:1000] - 0.03
    boxes[1::2] = q[:1000] + 0.02
    res = t.search_box(boxes)
    offs, flat = o.search_box(boxes[0::2], boxes[1::2])
    assert np.array_equal(res._offsets, offs)
    for a, b in zip(split_ragged(res._offsets, res._flat), split_ragged(offs, flat)):
        assert np.array_equal(np.sort(a), np.sort(b))


@pytest.mark.parametrize("leaf", [1, 2, 10])
def test_exact_knn_bound_priming_stress(pt, oracle, leaf):
    """k > 1 exact search prunes with an upper bound taken from stored points next to the first leaf
    (traverse.cuh, kPrimeBound). Worst cases for it: one-point leaves (the bound point's own node is a
    far child), queries that coincide with tree points, duplicated points, planes of equal coordinates,
    k from 2 to 16 — answers must still be those of the reference traversal."""
    rng = np.random.default_rng(100 + leaf)
    for sdim in (2, 3):
        pts = rng.random((60_000, sdim)).astype(np.float32)
        pts[:5_000, 0] = 0.25                      # a plane
        pts[5_000:7_000] = pts[7_000:9_000]        # exact duplicates
        pts[9_000:12_000] = np.round(pts[9_000:12_000] * 64) / 64   # lattice: many equal distances
        q = np.concatenate([pts[rng.integers(0, len(pts), 15_000)],            # on tree points
                            rng.random((15_000, sdim), dtype=np.float32),
                            (np.round(rng.random((5_000, sdim)) * 64) / 64).astype(np.float32)])
        q = np.ascontiguousarray(q.astype(np.float32))
        o = oracle.OracleTree(pts, leaf)
        t = pt.KdTree(pts, pt.Metric.L2Squared, leaf)
        for k in (2, 3, 7, 16):
            want = o.search_knn(q, k, threads=oracle.max_threads())
            assert_knn_parity(t.search_knn(q, k), want, pts, q)
        t1 = pt.KdTree(pts, pt.Metric.L1, leaf)
        o1 = oracle.OracleTree(pts, leaf, metric="l1")
        assert_knn_parity(t1.search_knn(q, 5), o1.search_knn(q, 5, threads=oracle.max_threads()), pts, q, "l1")


def test_lidar_shape_vs_oracle(pt, oracle):
    """cfg2's cloud at a size the oracle finishes in seconds: slides in both directions, deep tree."""
    from pico_tree_b200 import datasets as D
    pts = D.lidar_shape(400_000, seed=1)
    q = D.lidar_shape(200_000, seed=2, pose_shift=0.35)
    o = oracle.OracleTree(pts, 10)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    nodes, indices, box = t.export()
    assert_same_structure(nodes_from_export(nodes, pts.dtype), o.nodes)
    assert np.array_equal(indices, o.indices)  # slides in both directions, reproduced index for index
    want = o.search_knn(q, 1, threads=oracle.max_threads())
    for kw in ({}, {"reorder": False}, {"warp_per_query": True}):
        ties = assert_knn_parity(t.search_knn(q, 1, **kw), want, pts, q)
        assert ties == 0
    assert_knn_parity(t.search_knn(q, 16), o.search_knn(q, 16, threads=oracle.max_threads()), pts, q)
    nns = t.search_radius(q[:50_000], 0.01)
    offs, flat = o.search_radius(q[:50_000], 0.01)
    assert_radius_parity(nns._offsets, nns._flat[:len(flat)], offs, flat, ordered=False)


@pytest.mark.parametrize("metric", ["l1", "lpinf", "lninf"])
def test_other_metrics_and_approximate(pt, oracle, metric):
    from pico_tree_b200 import datasets as D
    pts = D.uniform(40_000, 3, seed=5)
    q = D.uniform(5_000, 3, seed=6)
    o = oracle.OracleTree(pts, 8, metric=metric)
    t = make_tree(pt, pts, metric, stop_value=8)
    assert_knn_parity(t.search_knn(q, 5), o.search_knn(q, 5), pts, q, metric)
    assert_knn_parity(t.search_knn(q, 24), o.search_knn(q, 24), pts, q, metric)
    # approximate results depend on the traversal; identical trees -> identical answers
    got, want = t.search_knn(q, 3, 1.7), o.search_knn(q, 3, e=1.7)
    assert np.array_equal(got["distance"], want["distance"]) and np.array_equal(got["index"], want["index"])


@pytest.mark.parametrize("metric,sdim,dtype", [("so2", 1, np.float32), ("se2_squared", 3, np.float32),
                                               ("se2_squared", 3, np.float64)])
def test_topological_metrics(pt, oracle, metric, sdim, dtype):
    """search_nearest_topological (kd_tree_search.hpp:122-229) and the wrapped box search on device:
    circle S1 = [0, 1) and R2 x S1, against the oracle (itself pinned to the reference's fixtures)."""
    rng = np.random.default_rng(17)
    pts = rng.random((60_000, sdim)).astype(dtype)
    q = rng.random((8_000, sdim)).astype(dtype)
    q[:50] = pts[:50]
    q[50:60, -1] = dtype(0.0)
    q[60:70, -1] = np.nextafter(dtype(1.0), dtype(0.0))
    o = oracle.OracleTree(pts, 10, metric=metric)
    t = make_tree(pt, pts, metric)
    nodes, indices, box = t.export()
    assert_same_structure(nodes_from_export(nodes, pts.dtype), o.nodes)
    assert np.array_equal(t.export_outer_bounds(), o.outer_bounds)
    assert np.array_equal(indices, o.indices)
    for kw in ({}, {"warp_per_query": True}):
        for k in (1, 8, 20):
            assert_knn_parity(t.search_knn(q, k, **kw), o.search_knn(q, k), pts, q, metric)
        got, want = t.search_knn(q, 4, 1.8, **kw), o.search_knn(q, 4, e=1.8)
        assert_knn_parity(got, want, pts, q, metric, e=1.8)
        r = 0.002 if metric == "so2" else 0.001
        nns = t.search_radius(q, r, **kw)
        offs, flat = o.search_radius(q, r)
        assert_radius_parity(nns._offsets, nns._flat[:len(flat)], offs, flat, ordered=False)
    mins = (q[:3000] - dtype(0.03)).astype(dtype)
    maxs = (q[:3000] + dtype(0.02)).astype(dtype)
    s1 = slice(0, None) if metric == "so2" else slice(2, None)
    mins[:, s1] = np.mod(mins[:, s1], 1)
    maxs[:, s1] = np.mod(maxs[:, s1], 1)
    assert np.any(mins[:, -1] > maxs[:, -1])  # some boxes wrap around the circle
    boxes = np.empty((2 * len(mins), sdim), dtype)
    boxes[0::2], boxes[1::2] = mins, maxs
    res = t.search_box(boxes)
    offs, flat = o.search_box(mins, maxs)
    assert np.array_equal(res._offsets, offs)
    for a, b in zip(split_ragged(res._offsets, res._flat), split_ragged(offs, flat)):
        assert np.array_equal(np.sort(a), np.sort(b))
    with pytest.raises(pt._lib.PicoB200Error):
        pt.KdTree(np.zeros((10, 2), dtype), pt.Metric.SO2, 10)


def test_high_dim_runtime_path(pt, oracle):
    """cfg4-like (runtime-dim path): 128-D descriptors, knn=10 exact and approximate."""
    from pico_tree_b200 import datasets as D
    pts = D.sift_shape(20_000, seed=1)
    q = D.sift_shape(64, seed=2)
    o = oracle.OracleTree(pts, 10)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    assert_knn_parity(t.search_knn(q, 10), o.search_knn(q, 10), pts, q)
    assert_knn_parity(t.search_knn(q, 100), o.search_knn(q, 100), pts, q)
    # approximate search is traversal dependent: compare on the reference-built tree
    import ctypes as C
    got = t.search_knn(q, 10, 2.25)
    exact = o.search_knn(q, 10)
    assert np.all(got["distance"][:, 0] * np.float32(2.25) >= exact["distance"][:, 0])


def test_float64(pt, oracle):
    rng = np.random.default_rng(11)
    pts = rng.random((30_000, 3))
    q = rng.random((4_000, 3))
    o = oracle.OracleTree(pts, 10)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    assert t.dtype_neighbor.itemsize == 16
    nodes, indices, _ = t.export()
    assert_same_structure(nodes_from_export(nodes, pts.dtype), o.nodes)
    assert np.array_equal(indices, o.indices)
    assert_knn_parity(t.search_knn(q, 1), o.search_knn(q, 1), pts, q)
    assert_knn_parity(t.search_knn(q, 12), o.search_knn(q, 12), pts, q)
    nns = t.search_radius(q, 0.0005)
    offs, flat = o.search_radius(q, 0.0005)
    assert_radius_parity(nns._offsets, nns._flat[:len(flat)], offs, flat, ordered=False)


# ---------------------------------------------------------------- edge cases
def test_edge_cases(pt, oracle):
    # a single point, k larger than n, all points identical, queries far outside the root box
    one = np.array([[0.5, 0.25, 0.125]], np.float32)
    t = pt.KdTree(one, pt.Metric.L2Squared, 10)
    r = t.search_knn(np.array([[0, 0, 0], [9, 9, 9]], np.float32), 1)
    assert r["index"].ravel().tolist() == [0, 0]
    same = np.tile(np.array([[1.0, 2.0, 3.0]], np.float32), (100, 1))
    t = pt.KdTree(same, pt.Metric.L2Squared, 10)
    o = oracle.OracleTree(same, 10)
    q = np.array([[1, 2, 3], [0, 0, 0]], np.float32)
    assert_knn_parity(t.search_knn(q, 7), o.search_knn(q, 7), same, q)
    assert len(t.search_radius(q, 1e-12)[0]) == 100  # all duplicates at distance 0 < r
    assert len(t.search_radius(q, 0.0)[0]) == 0      # strict '<'
    # k == n returns everything, sorted
    pts = np.random.default_rng(0).random((50, 2), dtype=np.float32)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 3)
    o = oracle.OracleTree(pts, 3)
    q = np.random.default_rng(1).random((20, 2), dtype=np.float32) * 3 - 1
    got = t.search_knn(q, 50)
    assert_knn_parity(got, o.search_knn(q, 50), pts, q)
    assert np.all(np.diff(got["distance"], axis=1) >= 0)
    # inclusive box bounds: a box that is exactly one point
    boxes = np.stack([pts[3], pts[3]])
    assert t.search_box(boxes)[0].tolist() == [3]
    # leaf_ranges with max_leaf_depth(2): 4 leaves covering every index once (kd_tree_test.cpp:148-178)
    t = pt.KdTree(pts, pt.Metric.L2Squared, max_leaf_depth=2)
    ranges = t.leaf_ranges()
    assert len(ranges) == 4 and sorted(np.concatenate(ranges).tolist()) == list(range(50))
    # empty query batch
    assert t.search_knn(np.empty((0, 2), np.float32), 1).shape == (0, 1)
    # column-major queries give (k, n) output (kd_tree_test.py:71-88)
    qf = np.asfortranarray(q.T)
    assert qf.flags["F_CONTIGUOUS"]
    got_f = t.search_knn(qf, 2)
    assert got_f.shape == (2, 20)
    assert np.array_equal(got_f.T["index"], t.search_knn(q, 2)["index"])


def test_empty_and_degenerate_batches(pt, oracle):
    """Empty and ragged inputs: no queries, no hits at all, k beyond the point count, boxes that hold
    nothing / everything, one-point leaves."""
    rng = np.random.default_rng(4)
    pts = rng.random((300, 3), dtype=np.float32)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 1)          # every leaf holds exactly one point
    assert t.info()["n_leaves"] == 300
    none = np.empty((0, 3), np.float32)
    assert len(t.search_radius(none, 1.0)) == 0
    assert len(t.search_box(none)) == 0
    far = np.full((5, 3), 100.0, np.float32)
    rad = t.search_radius(far, 1e-3)
    assert len(rad) == 5 and all(len(r) == 0 for r in rad)  # ragged result with zero total hits
    everything = np.array([[-1, -1, -1], [2, 2, 2], [5, 5, 5], [6, 6, 6]], np.float32)
    res = t.search_box(everything)
    assert sorted(res[0].tolist()) == list(range(300)) and len(res[1]) == 0
    # k beyond the point count through the C-ABI: the first n slots are the sorted answer, the rest stay
    # {-1, max} (the reference's iterator overload leaves them unspecified, search_visitor.hpp:98-103)
    q = rng.random((40, 3), dtype=np.float32)
    for kw in ({}, {"warp_per_query": True}):
        got = t.search_knn(q, 320, **kw)
        want = oracle.OracleTree(pts, 1).search_knn(q, 300)
        assert np.array_equal(got["distance"][:, :300], want["distance"])
        assert np.all(got["index"][:, 300:] == -1) and np.all(got["distance"][:, 300:] == np.finfo(np.float32).max)
    # all queries identical, all points identical in one coordinate
    pts2 = rng.random((5000, 2), dtype=np.float32)
    pts2[:, 1] = 0.5
    t2 = pt.KdTree(pts2, pt.Metric.L2Squared, 10)
    o2 = oracle.OracleTree(pts2, 10)
    q2 = np.tile(np.array([[0.3, 0.5]], np.float32), (3000, 1))
    assert_knn_parity(t2.search_knn(q2, 3), o2.search_knn(q2, 3), pts2, q2)


def test_python_api_three_point_cases(pt):
    # test/pyco_tree/kd_tree_test.py:53-69,90-118,151-192
    a = np.array([[2, 1], [4, 3], [8, 7]], np.float32)
    t = pt.KdTree(a, pt.Metric.L2Squared, 10)
    assert (t.sdim, t.npts) == (2, 3) and t.metric(-2.0) == 4
    assert pt.KdTree(a, pt.Metric.L1, 10).metric(-2.0) == 2
    nns = t.search_knn(a, 2)
    assert nns.shape == (3, 2) and nns["index"][:, 0].tolist() == [0, 1, 2] and np.all(nns["distance"][:, 0] == 0)
    data = nns.ctypes.data
    t.search_knn(a, 2, nns)
    assert nns.ctypes.data == data
    nns = t.search_knn(a, 2, 1.0)
    assert nns["index"][:, 0].tolist() == [0, 1, 2]
    rad = t.search_radius(a, t.metric(2.5))
    assert len(rad) == 3 and rad.dtype == t.dtype_neighbor
    assert [len(n) for n in rad] == [1, 1, 1] and [int(n[0][0]) for n in rad] == [0, 1, 2]
    boxes = np.array([[0, 0], [3, 3], [2, 2], [3, 3], [0, 0], [9, 9], [6, 6], [9, 9]], np.float32)
    res = t.search_box(boxes)
    assert res.dtype == t.dtype_index and [len(n) for n in res] == [1, 0, 3, 1]
    assert [len(n) for n in res[0:4:2]] == [1, 3]
    with pytest.raises(ValueError):
        t.search_box(boxes[:3])
    with pytest.raises(ValueError):
        t.search_knn(a.astype(np.float64), 1)


def test_deep_tree_uses_global_stack(pt, oracle):
    """Duplicated coordinates make the sliding midpoint peel one point per level: the tree gets
    deeper than the local traversal stack and the workspace variant must take over."""
    rng = np.random.default_rng(3)
    # exponentially spaced points: every split of the sliding midpoint peels off a few points only
    scale = np.float32(2.0) ** -np.arange(110, dtype=np.float32)
    pts = (rng.random((110, 40, 3), dtype=np.float32) * np.float32(0.4) + np.float32(0.6)) * scale[:, None, None]
    pts = np.ascontiguousarray(pts.reshape(-1, 3))
    rng.shuffle(pts, axis=0)
    o = oracle.OracleTree(pts, 2)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 2)
    assert t.info()["height"] == o.height and o.height >= 64
    q = np.ascontiguousarray((rng.random((3000, 3), dtype=np.float32) *
                              np.float32(2.0) ** -rng.integers(0, 100, (3000, 1)).astype(np.float32)))
    for kw in ({}, {"warp_per_query": True}):
        assert_knn_parity(t.search_knn(q, 1, **kw), o.search_knn(q, 1), pts, q)
        assert_knn_parity(t.search_knn(q, 6, **kw), o.search_knn(q, 6), pts, q)
    nns = t.search_radius(q, 0.001)
    offs, flat = o.search_radius(q, 0.001)
    assert_radius_parity(nns._offsets, nns._flat[:len(flat)], offs, flat, ordered=False)


def test_python_device_arrays(pt):
    """KdTree.search_knn with torch CUDA tensors: no host copies, enqueued on torch's current stream."""
    import torch
    from pico_tree_b200 import datasets as D
    pts = D.uniform(50_000, 3, seed=1)
    q = D.uniform(30_000, 3, seed=2)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    want = t.search_knn(q, 5)
    qd = torch.from_numpy(q).cuda()
    got = t.search_knn(qd, 5)
    assert isinstance(got, torch.Tensor) and got.shape == (30_000, 5, 2) and got.dtype == torch.int32
    torch.cuda.synchronize()
    assert np.array_equal(got[..., 0].cpu().numpy(), want["index"])
    assert np.array_equal(got[..., 1].contiguous().view(torch.float32).cpu().numpy(), want["distance"])
    out = torch.empty((30_000, 1, 2), dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        r = t.search_knn(qd, 1, 1.5, out)
    s.synchronize()
    assert r is out
    assert np.array_equal(out[..., 0].cpu().numpy(), t.search_knn(q, 1, 1.5)["index"])
    with pytest.raises(ValueError):
        t.search_knn(qd.double(), 1)
    with pytest.raises(ValueError):
        t.search_knn(qd[:, :2], 1)
    # queries PRODUCED on torch's default stream (handle 0) right before the call: the search has to be ordered
    # after the producer (the library must not fall back to a private stream for handle 0)
    big = torch.from_numpy(D.uniform(4_000_000, 3, seed=3)).cuda()
    torch.cuda.synchronize()
    for _ in range(3):
        qd2 = torch.zeros_like(big)
        for _ in range(8):
            qd2 = qd2 * 0.5 + big * 0.5  # a chain of async producer kernels; ends close to `big`
        got2 = t.search_knn(qd2, 1)
        torch.cuda.synchronize()
        want2 = t.search_knn(qd2.cpu().numpy(), 1)
        assert np.array_equal(got2[..., 0].cpu().numpy(), want2["index"])


def test_device_resident_ragged_results(pt):
    """PICO_B200_DEVICE_POINTERS for radius / box: queries, offsets and hits stay in HBM; the bytes must
    equal what the host-buffer path returns."""
    import ctypes as C
    import torch
    from pico_tree_b200 import _lib, datasets as D
    L = _lib.lib()
    pts = D.uniform(80_000, 3, seed=3)
    q = D.uniform(20_000, 3, seed=4)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    want = t.search_radius(q, 0.0008, True)
    dev = torch.device("cuda", 0)
    qd = torch.from_numpy(q).to(dev)
    offs = torch.empty(len(q) + 1, dtype=torch.int64, device=dev)
    hits = C.c_void_p()
    _lib.check(L.pico_b200_radius(t._h, C.c_void_p(qd.data_ptr()), len(q), 3, 0.0008, 0.0, C.c_void_p(offs.data_ptr()),
                                  C.byref(hits), _lib.FLAG_DEVICE_POINTERS | _lib.FLAG_SORT_RESULTS, None))
    offs_h = offs.cpu().numpy().astype(np.uint64)
    assert np.array_equal(offs_h, want._offsets)
    total = int(offs_h[-1])
    got = torch.empty(total * 2, dtype=torch.int32, device=dev)
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(got.data_ptr()), hits, C.c_size_t(total * 8), C.c_int(3))
    L.pico_b200_free_device(hits)
    got = got.cpu().numpy().view(t.dtype_neighbor)
    assert np.array_equal(got["distance"], want._flat["distance"][:total])
    boxes_min = torch.from_numpy(np.ascontiguousarray(q[:4000] - np.float32(0.02))).to(dev)
    boxes_max = torch.from_numpy(np.ascontiguousarray(q[:4000] + np.float32(0.02))).to(dev)
    boffs = torch.empty(4001, dtype=torch.int64, device=dev)
    bhits = C.c_void_p()
    _lib.check(L.pico_b200_box(t._h, C.c_void_p(boxes_min.data_ptr()), C.c_void_p(boxes_max.data_ptr()), 4000, 3,
                               C.c_void_p(boffs.data_ptr()), C.byref(bhits), _lib.FLAG_DEVICE_POINTERS, None))
    both = np.empty((8000, 3), np.float32)
    both[0::2], both[1::2] = boxes_min.cpu().numpy(), boxes_max.cpu().numpy()
    wantb = t.search_box(both)
    assert np.array_equal(boffs.cpu().numpy().astype(np.uint64), wantb._offsets)
    nb = int(wantb._offsets[-1])
    gotb = torch.empty(nb, dtype=torch.int32, device=dev)
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(gotb.data_ptr()), bhits, C.c_size_t(nb * 4), C.c_int(3))
    L.pico_b200_free_device(bhits)
    assert np.array_equal(gotb.cpu().numpy(), wantb._flat[:nb])


def test_nth_element_heap_select_fallback(pt, oracle):
    """Adversarial inputs (oracle/make_killers.py) push libstdc++'s introselect past its depth limit into
    std::__heap_select; the device build's emulation must follow it there (warp and CTA flavours)."""
    data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "introselect_killers.npz"))
    for key in data.files:
        v = data[key]
        pts = np.ascontiguousarray(np.stack([v, np.zeros_like(v)], axis=1))
        o = oracle.OracleTree(pts, 10, rule="median")
        t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=pt.kd_tree.Rule.MedianMaxSide)
        nodes, indices, _ = t.export()
        assert np.array_equal(indices, o.indices), key
        assert_same_structure(nodes_from_export(nodes, pts.dtype), o.nodes)


def test_small_host_batches_with_full_lists(pt, oracle):
    """Small host batches are answered through a pinned, device-mapped buffer that the kernel writes over PCIe
    (search.cu SmallBuffer); with k equal to a list size of the kernel (4, 8, 16, 32) the finished list leaves as
    whole 32-byte sectors (traverse.cuh store_sector) — into host memory here. Other k take the record-wise path."""
    from pico_tree_b200 import datasets as D
    pts = D.lidar_shape(30_000, seed=4)
    q = D.lidar_shape(64, seed=5, pose_shift=0.3)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    o = oracle.OracleTree(pts, 10)
    for k in (1, 3, 4, 8, 9, 16, 32):
        for nq in (1, 5, 64):
            got, want = t.search_knn(q[:nq], k), o.search_knn(q[:nq], k)
            assert np.array_equal(got["index"], want["index"]), (k, nq)
            assert np.array_equal(got["distance"], want["distance"]), (k, nq)


def test_big_ragged_result_block_is_reused_correctly(pt):
    """A ragged host result of 64 MiB or more lives in a huge-page block that pico_b200_free keeps for the next big
    result (search.cu alloc_result / release_result). Results written into a reused block — bigger than, equal
    to and smaller than what the next call needs — must equal the ones written into fresh memory."""
    import gc
    from pico_tree_b200 import datasets as D
    pts = D.lidar_shape(400_000, seed=2)
    q = D.lidar_shape(300_000, seed=3, pose_shift=0.2)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)

    def snapshot(r2):
        res = t.search_radius(q, r2)
        off, flat = res._offsets.copy(), res._flat[:int(res._offsets[-1])].copy()
        del res
        gc.collect()  # the block goes back to the library here
        return off, flat

    r2s = (0.25, 0.25, 0.16, 0.25, 0.36, 0.25)
    first = {}
    for r2 in r2s:
        off, flat = snapshot(r2)
        assert flat.nbytes >= 64 << 20, "the case must reach the huge-page path"
        if r2 in first:
            assert np.array_equal(off, first[r2][0]) and np.array_equal(flat, first[r2][1]), r2
        else:
            first[r2] = (off, flat)


def test_median_rule_huge_nodes_match_oracle(pt, oracle, monkeypatch):
    """Median rule above the one-CTA threshold (build.cu median_huge_level: several CTAs per node, barriers in
    global memory): node table and index permutation equal to the oracle's std::nth_element build — at the
    default threshold, with a low one (many iterations by many groups), and on the introselect killers, whose
    depth limit runs out while the range is still in the multi-CTA phase."""
    from pico_tree_b200 import datasets as D
    pts = D.lidar_shape(400_000, seed=11)
    o = oracle.OracleTree(pts, 10, rule="median")
    for huge_min, coop in ((None, None), ("2048", None), (None, "0")):
        if huge_min:
            monkeypatch.setenv("PICO_B200_HUGE_MIN", huge_min)
        if coop:  # the fallback of a device that cannot hold the cooperative grid: one CTA per huge node
            monkeypatch.setenv("PICO_B200_MEDIAN_COOP", coop)
        t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=pt.kd_tree.Rule.MedianMaxSide)
        nodes, indices, _ = t.export()
        assert np.array_equal(indices, o.indices), (huge_min, coop)
        assert_same_structure(nodes_from_export(nodes, pts.dtype), o.nodes)
        monkeypatch.delenv("PICO_B200_HUGE_MIN", raising=False)
        monkeypatch.delenv("PICO_B200_MEDIAN_COOP", raising=False)
    data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "introselect_killers.npz"))
    monkeypatch.setenv("PICO_B200_HUGE_MIN", "1024")
    for key in data.files:
        v = data[key]
        pts = np.ascontiguousarray(np.stack([v, np.zeros_like(v)], axis=1))
        o = oracle.OracleTree(pts, 10, rule="median")
        t = pt.KdTree(pts, pt.Metric.L2Squared, 10, rule=pt.kd_tree.Rule.MedianMaxSide)
        nodes, indices, _ = t.export()
        assert np.array_equal(indices, o.indices), key
        assert_same_structure(nodes_from_export(nodes, pts.dtype), o.nodes)
    monkeypatch.delenv("PICO_B200_HUGE_MIN")


def test_build_paths_agree(pt, monkeypatch):
    """The three ways a node can be split on the device — one warp, one CTA, grid-wide chunked passes
    (build.cu, PICO_B200_HUGE_MIN is the test hook for the threshold) — must leave the very same
    tree AND the very same index permutation (std::partition's exchange order), slides included."""
    from pico_tree_b200 import datasets as D
    rng = np.random.default_rng(21)
    clustered = rng.random((150_000, 3)).astype(np.float32)
    clustered[:50_000, 2] = 0.25            # a plane: slides towards it
    clustered[50_000:70_000, 0] = 0.5
    clustered[::13] = clustered[7]          # duplicates of one point
    cases = [
        (D.lidar_shape(300_000, seed=3), {}),
        (clustered, {}),
        (rng.random((120_000, 2)), {}),  # float64
        (D.uniform(100_000, 3, seed=9), {"rule": pt.kd_tree.Rule.MidpointMaxSide}),
        (D.uniform(90_000, 3, seed=9), {"max_leaf_depth": 6}),
        (D.sift_shape(40_000, 16, seed=4), {}),
        # median rule: one CTA per node against groups of CTAs running introselect together (median_huge_level)
        (D.lidar_shape(300_000, seed=5), {"rule": pt.kd_tree.Rule.MedianMaxSide}),
        (clustered, {"rule": pt.kd_tree.Rule.MedianMaxSide}),
        (rng.random((120_000, 2)), {"rule": pt.kd_tree.Rule.MedianMaxSide}),
    ]
    for pts, kw in cases:
        got = []
        for huge_min in ("2147483647", "1024", "20000"):
            monkeypatch.setenv("PICO_B200_HUGE_MIN", huge_min)
            t = pt.KdTree(pts, pt.Metric.L2Squared, 10, **kw) if "max_leaf_depth" not in kw else \
                pt.KdTree(pts, pt.Metric.L2Squared, **kw)
            nodes, indices, box = t.export()
            got.append((nodes.tobytes(), indices.copy(), box.copy()))
        for other in got[1:]:
            assert other[0] == got[0][0], "node tables differ between build paths"
            assert np.array_equal(other[1], got[0][1]), "index permutations differ between build paths"
            assert np.array_equal(other[2], got[0][2])
    monkeypatch.delenv("PICO_B200_HUGE_MIN")


# ---------------------------------------------------------------- BASELINE sizes: properties
def test_isolated_leaf_scan(pt):
    """pico_b200_profile_leaf_scan (the leaf-scan measurement of SURVEY.md §8d): the neighbour it returns is
    the first nearest point, in leaf order, of the leaf a numpy descent over the exported nodes ends in
    (kd_tree_search.hpp:60-88 then :54-59), with bit-equal distance; the byte count matches the ranges."""
    import torch
    from parity import ref_distance
    from pico_tree_b200 import datasets as D
    pts = D.lidar_shape(200_000, seed=1)
    q = D.lidar_shape(100_000, seed=2, pose_shift=0.35)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    nodes, indices, _ = t.export()
    a, b = nodes["a"].view(np.float32), nodes["b"].view(np.float32)
    node = np.zeros(len(q), dtype=np.int64)
    rows = np.arange(len(q))
    while True:
        sd = nodes["split_dim"][node]
        live = sd != 0xFFFFFFFF
        if not live.any():
            break
        v = q[rows, np.where(live, sd, 0)]
        left = (a[node] + b[node] - v - v) > 0  # float32, left to right
        node = np.where(live, np.where(left, node + 1, nodes["right"][node]), node)
    lb = nodes["a"][node].astype(np.int64)
    le = nodes["b"][node].astype(np.int64)
    width = int((le - lb).max())
    pos = lb[:, None] + np.arange(width)[None, :]
    valid = pos < le[:, None]
    cand = indices[np.minimum(pos, len(indices) - 1)]
    d = np.stack([ref_distance(pts, q, cand[:, j]) for j in range(width)], axis=1)
    d[~valid] = np.inf
    best = np.argmin(d, axis=1)  # first minimum = the strict `max() > d` of search_nn
    nns, st = t.profile_leaf_scan(torch.from_numpy(q).cuda(), repeats=2)
    got = nns.cpu().numpy()
    assert np.array_equal(got[:, 0, 0], cand[rows, best])
    assert np.array_equal(got[:, 0, 1].view(np.float32), d[rows, best].astype(np.float32))
    assert st["scan_bytes"] == len(q) * (4 + 12 + 8 + 8) + int((le - lb).sum()) * 16
    assert st["scan_ms"] > 0 and st["descend_ms"] > 0
    # never better than the true nearest neighbour, and equal to it for most queries
    nn = t.search_knn(q, 1)
    assert np.all(got[:, 0, 1].view(np.float32) >= nn["distance"][:, 0])
    assert np.mean(got[:, 0, 0] == nn["index"][:, 0]) > 0.3


def test_pinned_host_pipeline(pt):
    """Pinned host buffers take the copy-ahead pipeline of pico_b200_knn (all H2D copies up front on a copy
    stream, head chunks, per-chunk Z-order / traversal / D2H): same neighbours as the resident path, for packed
    rows, for rows with padding (stride 4: the 2-D copy) and for pageable buffers (per-chunk copies)."""
    import ctypes as C

    import torch
    from pico_tree_b200 import _lib, datasets as D
    pts = D.uniform(300_000, 3, seed=3)
    nq = 2_600_000  # > 2 chunks of 1 Mi: the chunked path
    q = D.uniform(nq, 3, seed=4)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    want = t.search_knn(torch.from_numpy(q).cuda(), 1).cpu().numpy().reshape(nq, 2)
    L = _lib.lib()
    for k in (1, 3):
        ref = want if k == 1 else t.search_knn(torch.from_numpy(q).cuda(), k).cpu().numpy().reshape(nq, k * 2)
        # packed pinned rows
        qp = torch.from_numpy(q).pin_memory()
        out = torch.zeros((nq, k * 2), dtype=torch.int32).pin_memory()
        _lib.check(L.pico_b200_knn(t._h, C.c_void_p(qp.data_ptr()), nq, 3, k, 0.0, C.c_void_p(out.data_ptr()), 0, None))
        assert np.array_equal(out.numpy(), ref)
        # padded pinned rows (stride 4)
        q4 = torch.zeros((nq, 4), dtype=torch.float32).pin_memory()
        q4[:, :3] = torch.from_numpy(q)
        q4[:, 3] = 1e30
        out.zero_()
        _lib.check(L.pico_b200_knn(t._h, C.c_void_p(q4.data_ptr()), nq, 4, k, 0.0, C.c_void_p(out.data_ptr()), 0, None))
        assert np.array_equal(out.numpy(), ref)
        # pageable buffers, twice (the second big pageable batch of a thread goes through the pinned mirrors)
        for _ in range(2):
            got = t.search_knn(q, k)
            assert np.array_equal(got["index"], ref[:, 0::2])
            assert np.array_equal(got["distance"], ref[:, 1::2].view(np.float32))


def test_full_size_properties(pt):
    """cfg2 at full size (7,733,372 / 7,200,863): size-independent properties instead of the oracle —
    (a) nn of a tree point is itself at distance 0; (b) nn distance is a lower bound of the distance
    to a random sample of points and is attained by the returned index; (c) results do not depend on
    the traversal kernel or on query reordering; (d) knn=16 rows are sorted and start with the nn."""
    from parity import ref_distance
    from pico_tree_b200 import datasets as D
    tree_pts, q = D.bench_clouds()
    t = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
    info = t.info()
    assert info["n_points"] == D.N_TREE and info["n_leaves"] * 10 >= D.N_TREE
    sub = np.ascontiguousarray(tree_pts[::97])
    r = t.search_knn(sub, 1)
    assert np.all(r["distance"] == 0)
    assert np.array_equal(tree_pts[r["index"][:, 0]], sub)  # duplicates may answer for each other
    nn = t.search_knn(q, 1)
    d = ref_distance(tree_pts, q, nn["index"][:, 0])
    assert np.array_equal(d, nn["distance"][:, 0])
    rng = np.random.default_rng(0)
    for _ in range(4):
        other = rng.integers(0, D.N_TREE, size=len(q))
        assert np.all(ref_distance(tree_pts, q, other) >= nn["distance"][:, 0])
    nn2 = t.search_knn(q, 1, reorder=False)
    assert np.array_equal(nn2["distance"], nn["distance"]) and np.array_equal(nn2["index"], nn["index"])
    part = q[:500_000]
    nn3 = t.search_knn(part, 1, warp_per_query=True)
    assert np.array_equal(nn3["index"], nn["index"][:500_000])
    k16 = t.search_knn(part, 16)
    assert np.all(np.diff(k16["distance"], axis=1) >= 0)
    assert np.array_equal(k16["index"][:, 0], nn["index"][:500_000, 0])
    # radius search agrees with knn: count(d < r) >= 1 iff nn distance < r
    rad = t.search_radius(part, 0.01)
    counts = np.diff(rad._offsets.astype(np.int64))
    assert np.array_equal(counts > 0, nn["distance"][:500_000, 0] < np.float32(0.01))


def test_nccl_tree_broadcast_two_gpus(pt):
    """pico_b200_tree_broadcast with raw ncclComm_t handles (ncclCommInitAll, one host thread per GPU):
    the replica on GPU 1 must answer exactly like the tree built on GPU 0. Needs two visible GPUs."""
    import ctypes as C
    import glob
    import threading
    import torch
    from pico_tree_b200 import _lib, datasets as D
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    libs = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
    nccl = C.CDLL(libs[0] if libs else "libnccl.so.2", mode=C.RTLD_GLOBAL)
    comms = (C.c_void_p * 2)()
    devs = (C.c_int * 2)(0, 1)
    assert nccl.ncclCommInitAll(comms, 2, devs) == 0
    L = _lib.lib()
    pts = D.lidar_shape(200_000, seed=1)
    q = D.lidar_shape(50_000, seed=2, pose_shift=0.35)
    tree0 = pt.KdTree(pts, pt.Metric.L2Squared, 10, device=0)
    handles = [C.c_void_p(tree0._h.value), C.c_void_p()]
    rcs = [None, None]

    def run(rank):
        rcs[rank] = L.pico_b200_tree_broadcast(C.byref(handles[rank]), comms[rank], rank, 0, rank)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in threads]
    [t.join(timeout=120) for t in threads]
    assert rcs == [0, 0], (rcs, L.pico_b200_last_error())
    assert handles[1].value and handles[1].value != handles[0].value
    inf = _lib.TreeInfo()
    _lib.check(L.pico_b200_tree_info_get(handles[1], C.byref(inf)))
    assert inf.device == 1 and inf.n_points == len(pts) and inf.n_nodes == tree0.info()["n_nodes"]
    want = tree0.search_knn(q, 4)
    got = np.empty_like(want)
    _lib.check(L.pico_b200_knn(handles[1], C.c_void_p(q.ctypes.data), len(q), 3, 4, 0.0, C.c_void_p(got.ctypes.data), 0,
                               None))
    assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["distance"], want["distance"])
    L.pico_b200_tree_destroy(handles[1])
    for c in comms:
        nccl.ncclCommDestroy(C.c_void_p(c))


@pytest.mark.parametrize("dtype,sdim", [(np.float32, 3), (np.float32, 2), (np.float64, 3)])
def test_search_image_nn_with_ties(pt, oracle, dtype, sdim):
    """k = 1 goes through the search image (fat.cu, nn_fat_kernel). On an integer lattice most queries have
    several points at the minimum distance: those are detected and re-run by the order-exact kernel, so the
    INDEX is still the reference's first-visited one. Duplicated points included."""
    rng = np.random.default_rng(11)
    pts = rng.integers(0, 24, size=(60_000, sdim)).astype(dtype)
    q = (rng.integers(0, 24, size=(40_000, sdim)) + rng.choice([0.0, 0.5], size=(40_000, sdim))).astype(dtype)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    ora = oracle.OracleTree(pts, 10)
    want = ora.search_knn(q, 1)
    for kw in ({}, {"reorder": False}):
        got = t.search_knn(q, 1, **kw)
        assert np.array_equal(got["distance"], want["distance"])
        assert np.array_equal(got["index"], want["index"])
    # and a cloud without ties, far children of every depth: uniform points, queries partly outside the box
    pts = rng.random((200_000, sdim)).astype(dtype)
    q = (rng.random((100_000, sdim)) * 1.5 - 0.25).astype(dtype)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    want = oracle.OracleTree(pts, 10).search_knn(q, 1)
    got = t.search_knn(q, 1)
    assert np.array_equal(got["distance"], want["distance"]) and np.array_equal(got["index"], want["index"])


NN_VARIANT = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import pico_tree_b200 as pt
from oracle import oracle as O
rng = np.random.default_rng(11)
for dtype, sdim in ((np.float32, 3), (np.float32, 2), (np.float64, 3)):
    pts = rng.integers(0, 24, size=(60_000, sdim)).astype(dtype)                      # lattice: ties everywhere
    q = (rng.integers(0, 24, size=(40_000, sdim)) + rng.choice([0.0, 0.5], size=(40_000, sdim))).astype(dtype)
    want = O.OracleTree(pts, 10).search_knn(q, 1)
    got = pt.KdTree(pts, pt.Metric.L2Squared, 10).search_knn(q, 1)
    assert np.array_equal(got["distance"], want["distance"]) and np.array_equal(got["index"], want["index"])
    pts = rng.random((200_000, sdim)).astype(dtype)
    q = (rng.random((100_000, sdim)) * 3 - 1).astype(dtype)                           # many far children pending
    want = O.OracleTree(pts, 4).search_knn(q, 1)
    got = pt.KdTree(pts, pt.Metric.L2Squared, 4).search_knn(q, 1)
    assert np.array_equal(got["distance"], want["distance"]) and np.array_equal(got["index"], want["index"])
print("NN_VARIANT_OK")
'''


@pytest.mark.parametrize("env", [{"PICO_B200_NN": "1"}, {"PICO_B200_NN": "5"},
                                 {"PICO_B200_NN": "1", "PICO_B200_FAT_LEAF": "16"},
                                 {"PICO_B200_NN": "3", "PICO_B200_FAT_LEAF": "32"}, {"PICO_B200_NN": "32"}],
                         ids=["slot_stack", "slot_stack_no_restart", "search_image_16", "search_image_32_far_fat",
                              "far_subtrees_as_items"])
def test_nn_kernel_variants(env):
    """The selectable k = 1 kernels (PICO_B200_NN / PICO_B200_FAT_LEAF tuning hooks, read once per process): shared
    slot stack with restore records, prefix-minimum restart, search image with tie re-run. Same answers as the
    oracle, ties included."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", NN_VARIANT % {"root": root}], capture_output=True, text=True,
                       env=dict(os.environ, **env), timeout=600)
    assert r.returncode == 0 and "NN_VARIANT_OK" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_tree_image_round_trip_and_validation(pt):
    """pico_b200_tree_serialize / _deserialize (what a replica receives): the copy answers like the original, and
    an image with a damaged header, node link, leaf range or index is refused instead of being searched."""
    import ctypes as C
    from pico_tree_b200 import _lib, datasets as D
    L = _lib.lib()
    pts = D.lidar_shape(80_000, seed=1)
    q = D.lidar_shape(20_000, seed=2, pose_shift=0.35)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    image = t.serialize()
    want = t.search_knn(q, 4)

    def load(buf):
        h = C.c_void_p()
        rc = L.pico_b200_tree_deserialize(C.c_void_p(buf.ctypes.data), buf.size, 0, 0, C.byref(h))
        return rc, h

    rc, h = load(image)
    assert rc == 0
    got = np.empty_like(want)
    _lib.check(L.pico_b200_knn(h, C.c_void_p(q.ctypes.data), len(q), 3, 4, 0.0, C.c_void_p(got.ctypes.data), 0, None))
    assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["distance"], want["distance"])
    L.pico_b200_tree_destroy(h)
    info = t.info()
    header = 128  # ImageHeader padded to 16 bytes: magic, 4 x u32, 5 x u64, 8 doubles = 128 bytes
    nodes_at = header + 32  # root box: 6 floats padded to 32 bytes
    for what, off, value in (("scalar", 12, 7), ("metric", 16, 99), ("sdim", 32, 0), ("n_nodes", 40, 4),
                             ("right link of the root", nodes_at + 8, 0),
                             ("split_dim of the root", nodes_at + 12, 9),
                             ("an index", nodes_at + ((info["n_nodes"] * 16 + 15) // 16) * 16 + 4 * 100, -5)):
        bad = image.copy()
        bad[off:off + 4] = np.frombuffer(np.int32(value).tobytes(), np.uint8)
        rc, h = load(bad)
        assert rc != 0 and not h.value, what
    rc, h = load(image[:len(image) // 2].copy())
    assert rc != 0


def test_query_ordering_adapts_to_the_batch(pt, oracle):
    """Scan-order batches are Z-ordered tile by tile (tile_order_kernel), shuffled ones with the device-wide sort;
    the tree decides from the last measured batch (pico_b200_tree_order_state) and the answers never change."""
    import ctypes as C
    from pico_tree_b200 import _lib, datasets as D
    L = _lib.lib()
    pts = D.lidar_shape(400_000, seed=1)
    q = D.lidar_shape(300_000, seed=2, pose_shift=0.35)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    want = oracle.OracleTree(pts, 10).search_knn(q, 4)

    def state():
        s = C.c_int(-1)
        _lib.check(L.pico_b200_tree_order_state(t._h, C.byref(s)))
        return s.value

    assert state() == 0
    states = []
    for _ in range(3):  # single-neighbour searches measure and use the tile order
        got = t.search_knn(q, 1)
        assert np.array_equal(got["index"], want["index"][:, :1]) and np.array_equal(got["distance"], want["distance"][:, :1])
        states.append(state())
    assert states[-1] == 1, ("a scan-order batch should be ordered tile by tile", states)
    got = t.search_knn(q, 4)
    assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["distance"], want["distance"])
    perm = np.random.default_rng(0).permutation(len(q))
    qs = np.ascontiguousarray(q[perm])
    states = []
    for _ in range(3):
        got = t.search_knn(qs, 1)
        assert np.array_equal(got["index"], want["index"][perm][:, :1])
        states.append(state())
    assert states[-1] == 2, ("a shuffled batch should go back to the device-wide sort", states)
    for _ in range(40):  # the global path measures again every 16th call
        t.search_knn(q, 1)
    assert state() == 1
    r = t.search_radius(q[:50_000], 0.01)
    offs, flat = oracle.OracleTree(pts, 10).search_radius(q[:50_000], 0.01)
    assert np.array_equal(r._offsets, offs) and np.array_equal(r._flat["index"][:len(flat)], flat["index"])


def test_host_pipeline_orders_inside_the_traversal_for_coherent_batches(pt, oracle):
    """A scan-order batch from host memory: after the tree has measured one such batch, the chunks of the host pipeline
    are ordered and searched by ONE kernel each (nn_tile_kernel). Pinned, padded and pageable buffers, against the
    oracle; then a shuffled batch, which must fall back to the device-wide sort — same answers throughout."""
    import ctypes as C

    import torch
    from pico_tree_b200 import _lib, datasets as D
    L = _lib.lib()
    pts = D.lidar_shape(500_000, seed=1)
    nq = 2_700_000
    q = D.lidar_shape(nq, seed=2, pose_shift=0.35)
    t = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    want = oracle.OracleTree(pts, 10).search_knn(q, 1, threads=oracle.max_threads())

    def state():
        s = C.c_int(-1)
        _lib.check(L.pico_b200_tree_order_state(t._h, C.byref(s)))
        return s.value

    qp = torch.from_numpy(q).pin_memory()
    out = torch.zeros((nq, 2), dtype=torch.int32).pin_memory()
    for call in range(3):  # call 0 measures, calls 1.. use the fused kernel
        out.zero_()
        _lib.check(L.pico_b200_knn(t._h, C.c_void_p(qp.data_ptr()), nq, 3, 1, 0.0, C.c_void_p(out.data_ptr()), 0, None))
        got = out.numpy()
        assert np.array_equal(got[:, 0], want["index"][:, 0]), call
        assert np.array_equal(got[:, 1].view(np.float32), want["distance"][:, 0]), call
    assert state() == 1
    q4 = torch.zeros((nq, 4), dtype=torch.float32).pin_memory()
    q4[:, :3] = torch.from_numpy(q)
    out.zero_()
    _lib.check(L.pico_b200_knn(t._h, C.c_void_p(q4.data_ptr()), nq, 4, 1, 0.0, C.c_void_p(out.data_ptr()), 0, None))
    assert np.array_equal(out.numpy()[:, 0], want["index"][:, 0])
    for _ in range(2):  # pageable
        got = t.search_knn(q, 1)
        assert np.array_equal(got["index"], want["index"]) and np.array_equal(got["distance"], want["distance"])
    perm = np.random.default_rng(1).permutation(nq)
    qs = torch.from_numpy(np.ascontiguousarray(q[perm])).pin_memory()
    for call in range(20):  # the pipeline measures its first chunk every 16th call
        out.zero_()
        _lib.check(L.pico_b200_knn(t._h, C.c_void_p(qs.data_ptr()), nq, 3, 1, 0.0, C.c_void_p(out.data_ptr()), 0, None))
        if call in (0, 19):
            assert np.array_equal(out.numpy()[:, 0], want["index"][perm, 0]), call
    assert state() == 2

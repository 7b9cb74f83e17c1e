"""kd_forest on the device (csrc/forest.cu) against the outputs of the UNMODIFIED reference
kd_forest (tests/golden/forest/*.npz, oracle/make_golden_forest.py) and against the oracle's forest on larger
seeded inputs. Needs a B200: `-m gpu`."""
import glob
import os

import numpy as np
import pytest

from test_forest_core import flat_nodes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _forest_files():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "forest", "*.npz")))


@pytest.fixture(scope="module")
def pt():
    import pico_tree_b200
    return pico_tree_b200


@pytest.mark.parametrize("path", _forest_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_forest_matches_reference_fixture(pt, oracle, path):
    g = np.load(path)
    pts, q, rot = g["pts"], g["q"], np.ascontiguousarray(g["rotations"])
    f = pt.KdForest(pts, int(g["max_leaf_size"]), len(rot), rotations=rot)
    assert np.array_equal(f.rotations, rot)
    inf = f.info()
    assert inf["n_trees"] == len(rot) and inf["n_points"] == len(pts) and inf["sdim"] == pts.shape[1]
    # every tree: node for node, bound for bound and index for index the oracle's tree over the reflected copy
    of = oracle.OracleForest(pts, rot, int(g["max_leaf_size"]))
    for t in range(len(rot)):
        tv = of.tree(t)
        nodes, indices, _, outer = f.export_tree(t)
        want = flat_nodes(tv.nodes, pts.dtype)
        for fld in ("a", "b", "right", "split_dim"):
            assert np.array_equal(nodes[fld], want[fld]), (t, fld)
        assert np.array_equal(indices, tv.indices)
        branch = nodes["split_dim"] != 0xFFFFFFFF
        assert np.array_equal(outer[branch], np.asarray(tv.outer_bounds).reshape(-1, 2)[branch])
    for k, ml in g["searches"]:
        got = f.search_knn(q, int(k), int(ml))
        assert np.array_equal(got["index"], g[f"index_k{k}_m{ml}"]), (k, ml)
        assert np.array_equal(got["distance"], g[f"distance_k{k}_m{ml}"]), (k, ml)


@pytest.mark.parametrize("n,sdim,dtype,trees,k,leaves", [
    (60_000, 3, np.float32, 4, 1, 8), (40_000, 16, np.float32, 6, 10, 32), (20_000, 128, np.float32, 4, 10, 64),
    (30_000, 8, np.float64, 3, 5, 16), (5_000, 5, np.float32, 2, 40, 700), (30_000, 16, np.float32, 2, 4, 5000)])
def test_forest_vs_oracle(pt, oracle, n, sdim, dtype, trees, k, leaves):
    """Seeded clouds of several dimensions (packed float4 leaves, lane-per-point rows, staged row tiles), k in the
    register list and in the memory list, leaf budgets in the shared-memory queue and in the global one."""
    rng = np.random.default_rng(n + sdim)
    pts = rng.random((n, sdim)).astype(dtype)
    q = rng.random((2_000, sdim)).astype(dtype)
    rot = rng.normal(size=(trees, sdim))
    rot = (rot / np.linalg.norm(rot, axis=1, keepdims=True)).astype(dtype)
    f = pt.KdForest(pts, 10, trees, rotations=rot)
    of = oracle.OracleForest(pts, rot, 10)
    want = of.search_knn(q, k, leaves, threads=oracle.max_threads())
    got = f.search_knn(q, k, leaves)
    assert np.array_equal(got["distance"], want["distance"])
    assert np.array_equal(got["index"], want["index"])


def test_forest_random_rotations_and_recall(pt):
    """Without given vectors the forest draws its own (unit length); more trees / more leaves -> recall grows
    towards the exact answer (the figures examples/kd_forest/kd_forest.cpp:113-123 prints)."""
    rng = np.random.default_rng(7)
    pts = rng.random((50_000, 32)).astype(np.float32)
    q = rng.random((1_000, 32)).astype(np.float32)
    exact = pt.KdTree(pts, pt.Metric.L2Squared, 10).search_knn(q, 1)["index"][:, 0]
    recalls = []
    for trees, leaves in ((1, 4), (4, 32), (8, 256)):
        f = pt.KdForest(pts, 10, trees)
        r = f.rotations
        assert np.allclose(np.linalg.norm(r, axis=1), 1.0, atol=1e-5)
        got = f.search_nn(q, leaves)["index"][:, 0]
        recalls.append(float(np.mean(got == exact)))
    assert recalls[0] <= recalls[1] <= recalls[2] and recalls[2] > 0.9, recalls


def test_forest_argument_errors(pt):
    pts = np.random.default_rng(0).random((100, 3)).astype(np.float32)
    with pytest.raises(ValueError):
        pt.KdForest(pts, 0, 2)
    with pytest.raises(ValueError):
        pt.KdForest(pts, 10, 0)
    with pytest.raises(ValueError):
        pt.KdForest(pts, 10, 2, rotations=np.zeros((3, 3), np.float32))
    f = pt.KdForest(pts, 10, 2)
    with pytest.raises(ValueError):
        f.search_knn(pts.astype(np.float64), 1, 4)
    assert f.search_knn(pts[:0], 3, 4).shape == (0, 3)
    r = f.search_knn(pts, 200, 1 << 30)   # k clamps to n; an unbounded budget visits everything
    assert r.shape == (100, 100) and np.all(np.diff(r["distance"], axis=1) >= 0)

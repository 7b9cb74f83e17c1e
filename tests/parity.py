"""Comparison helpers shared by the parity tests.

Bar (BASELINE.json north_star): indices bit-exact, squared distances within 1e-6 relative.
The CUDA path keeps the reference's operation order without FMA contraction, so the tests
demand BIT-EQUAL distances. The helpers still recognise the one class of index differences the
reference's own tests exclude ("Index is not tested in case it happens points have an equal
distance", test/pico_tree/common.hpp:195-196) — the returned point must then attain the very same
distance — and return how many such ties they saw; since the device build reproduces the
reference's index permutation exactly, the tests that build on the device assert that count to be 0.
"""
import numpy as np

GOLDEN_NODE_FIELDS = ("left_max", "right_min", "split_dim", "begin", "end", "left", "right")


def ref_distance(pts, q, idx, metric="l2_squared"):
    """metric(q, pts[idx]) with the reference's sequential accumulation (metric.hpp:36-51)
    in the scalar type of `pts`; q: (m, d), idx: (m,)."""
    p = pts[idx]
    t = q.astype(pts.dtype) - p
    one = pts.dtype.type(1.0)
    if metric == "so2":  # s1_distance of coordinate 0, distance.hpp:19-22
        d = np.abs(t[:, 0])
        return np.minimum(d, one - d)
    if metric == "se2_squared":  # metric.hpp:229-238
        d = np.zeros(len(q), dtype=pts.dtype)
        for j in range(2):
            d = d + t[:, j] * t[:, j]
        c = np.abs(t[:, 2])
        c = np.minimum(c, one - c)
        return d + c * c
    if metric == "l2_squared":
        terms = t * t
    else:
        terms = np.abs(t)
    if metric in ("l2_squared", "l1"):
        d = np.zeros(len(q), dtype=pts.dtype)
        for j in range(pts.shape[1]):
            d = d + terms[:, j]
        return d
    if metric == "lpinf":
        return terms.max(axis=1)
    return terms.min(axis=1)


def assert_knn_parity(got, want, pts, q, metric="l2_squared", e=0.0):
    """got / want: structured (nq, k) arrays. Returns the number of index ties."""
    assert got.shape == want.shape
    gd, wd = got["distance"], want["distance"]
    assert np.array_equal(gd, wd), f"distances differ in {np.count_nonzero(gd != wd)} slots"
    diff = got["index"] != want["index"]
    ties = int(np.count_nonzero(diff))
    if ties:
        rows, cols = np.nonzero(diff)
        d = ref_distance(pts, q[rows], got["index"][rows, cols], metric)
        if e and e > 0:
            d = d * (pts.dtype.type(1.0) / pts.dtype.type(e))
        assert np.array_equal(d, gd[rows, cols]), "index differs and is not an equal-distance tie"
        # every row must still list distinct points
        for r in np.unique(rows):
            assert len(set(got["index"][r].tolist())) == got.shape[1]
    return ties


def split_ragged(offsets, flat):
    offsets = offsets.astype(np.int64)
    return [flat[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]


def assert_radius_parity(got_off, got, want_off, want, ordered=True):
    """Radius results: same counts; same (index, distance) records — in visit order when the
    trees are order-identical, else as multisets (SURVEY.md §8c hazard 5)."""
    assert np.array_equal(got_off, want_off), "hit counts differ"
    if ordered:
        assert np.array_equal(got["index"], want["index"])
        assert np.array_equal(got["distance"], want["distance"])
        return
    for g, w in zip(split_ragged(got_off, got), split_ragged(want_off, want)):
        gs = np.sort(g, order=["distance", "index"])
        ws = np.sort(w, order=["distance", "index"])
        assert np.array_equal(gs["index"], ws["index"]) and np.array_equal(gs["distance"], ws["distance"])


def nodes_from_export(nodes, scalar_dtype):
    """pico_b200_tree_export records -> dict of arrays named like the oracle's node fields."""
    leaf = nodes["split_dim"] == 0xFFFFFFFF
    utype = np.uint32 if scalar_dtype == np.float32 else np.uint64
    a = nodes["a"].astype(utype)
    b = nodes["b"].astype(utype)
    out = {
        "split_dim": np.where(leaf, -1, nodes["split_dim"].astype(np.int64)).astype(np.int32),
        "left_max": np.where(leaf, 0, a.view(scalar_dtype)),
        "right_min": np.where(leaf, 0, b.view(scalar_dtype)),
        "begin": np.where(leaf, a.astype(np.int64), 0).astype(np.int32),
        "end": np.where(leaf, b.astype(np.int64), 0).astype(np.int32),
        "left": np.where(leaf, -1, np.arange(len(nodes)) + 1).astype(np.int32),
        "right": np.where(leaf, -1, nodes["right"].astype(np.int64)).astype(np.int32),
    }
    return out


def assert_same_structure(got, want_nodes):
    for f in GOLDEN_NODE_FIELDS:
        w = want_nodes[f]
        assert np.array_equal(got[f], w), f"node field {f} differs at {np.flatnonzero(got[f] != w)[:5]}"


def leaf_sets_equal(nodes, idx_a, idx_b):
    """True when every leaf holds the same SET of point indices in both permutations."""
    leaf = nodes["split_dim"] < 0
    for b, e in zip(nodes["begin"][leaf], nodes["end"][leaf]):
        if not np.array_equal(np.sort(idx_a[b:e]), np.sort(idx_b[b:e])):
            return False
    return True

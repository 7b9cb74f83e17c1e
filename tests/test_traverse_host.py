"""The product's thread-per-query traversals (pico_tree_b200/csrc/traverse.cuh) compiled for the host
(tests/cpp/traverse_host.cpp) and run over the oracle's tree: control flow of the search-image nn traversal
(collapsed subtrees, prefix-minimum restart, tie detection) and of the order-exact traverse_packed, checked
without a GPU. The `-m gpu` tests check the same code as it runs on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_forest_core import flat_nodes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "_bin", "libtraverse_host.so")


@pytest.fixture(scope="module")
def host_lib():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s", LIB], check=True)
    return C.CDLL(LIB)


def fat_image(flat, limit):
    """What fat.cu's fat_nodes_kernel computes: every branch with at most `limit` points below it becomes a leaf."""
    fat = flat.copy()
    n = len(flat)
    begin = np.zeros(n, np.int64)
    end = np.zeros(n, np.int64)
    leaf = flat["split_dim"] == 0xFFFFFFFF
    for i in range(n - 1, -1, -1):  # pre-order: children have larger ids
        if leaf[i]:
            begin[i], end[i] = int(flat["a"][i]), int(flat["b"][i])
        else:
            begin[i], end[i] = begin[i + 1], end[int(flat["right"][i])]
    collapse = ~leaf & (end - begin <= limit)
    fat["a"][collapse] = begin[collapse].astype(fat["a"].dtype)
    fat["b"][collapse] = end[collapse].astype(fat["b"].dtype)
    fat["right"][collapse] = 0xFFFFFFFF
    fat["split_dim"][collapse] = 0xFFFFFFFF
    return fat, int(collapse.sum())


def leaf_order_points(pts, indices):
    """pts4 / double4 records in leaf order with the original index in .w."""
    n, sdim = pts.shape
    out = np.zeros((n, 4), pts.dtype)
    out[:, :sdim] = pts[indices]
    if pts.dtype == np.float32:
        out[:, 3] = indices.astype(np.int32).view(np.float32)
    else:
        out[:, 3] = indices.astype(np.int64).view(np.float64)
    return np.ascontiguousarray(out)


def run_nn_fat(lib, fat, far, pts4, q, nrec):
    f64 = q.dtype == np.float64
    fn = lib.host_nn_fat_f64 if f64 else lib.host_nn_fat_f32
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                   C.c_void_p]
    fn.restype = None
    idx = np.empty(len(q), np.int32)
    dist = np.empty(len(q), q.dtype)
    tie = np.empty(len(q), np.uint8)
    fn(fat.ctypes.data, far.ctypes.data, pts4.ctypes.data, q.ctypes.data, len(q), q.shape[1], nrec, idx.ctypes.data,
       dist.ctypes.data, tie.ctypes.data)
    return idx, dist, tie


def clouds(kind, n, nq, sdim, dtype, seed):
    from pico_tree_b200 import datasets as D
    if kind == "lidar":
        return D.lidar_shape(n, seed=seed, dtype=dtype), D.lidar_shape(nq, seed=seed + 1, pose_shift=0.35, dtype=dtype)
    rng = np.random.default_rng(seed)
    if kind == "grid":  # integer lattice: exact distance ties everywhere
        pts = rng.integers(0, 12, size=(n, sdim)).astype(dtype)
        q = rng.integers(0, 12, size=(nq, sdim)).astype(dtype) + dtype(0.5)
        return pts, q
    return rng.random((n, sdim)).astype(dtype), (rng.random((nq, sdim)) * 1.2 - 0.1).astype(dtype)


@pytest.mark.parametrize("kind,n,sdim,dtype,leaf,limit", [
    ("uniform", 20_000, 3, np.float32, 10, 32), ("uniform", 5_000, 2, np.float32, 1, 8),
    ("lidar", 30_000, 3, np.float32, 10, 32), ("lidar", 30_000, 3, np.float32, 10, 64),
    ("uniform", 8_000, 3, np.float64, 4, 16), ("grid", 6_000, 3, np.float32, 10, 32),
    ("uniform", 300, 3, np.float32, 10, 1000)])
def test_search_image_nn_matches_oracle(oracle, host_lib, kind, n, sdim, dtype, leaf, limit):
    pts, q = clouds(kind, n, 4_000, sdim, dtype, seed=3)
    tree = oracle.OracleTree(pts, leaf)
    flat = flat_nodes(tree.nodes, dtype)
    fat, collapsed = fat_image(flat, limit)
    assert collapsed > 0
    pts4 = leaf_order_points(pts, tree.indices)
    want = tree.search_knn(q, 1)
    for far in (flat, fat):          # far children in the real tree / in the search image
        for nrec in (0, 1, 3):       # without / with prefix-minimum restart records
            idx, dist, tie = run_nn_fat(host_lib, fat, far, pts4, q, nrec)
            assert np.array_equal(dist, want["distance"][:, 0]), (nrec, far is fat)
            ok = tie == 0
            # one point at the minimum distance: the index is the reference's; ties are flagged, never silent
            assert np.array_equal(idx[ok], want["index"][ok, 0]), (nrec, far is fat)
            diff = idx != want["index"][:, 0]
            assert not np.any(diff & ok)
            if kind == "grid":
                assert tie.sum() > len(q) // 4
            elif kind != "lidar":
                assert tie.sum() == 0
    # without a search image (real tree throughout, the reference's own prune test): the reference's visit
    # order, hence its index even where several points share the minimum distance
    for nrec in (100, 103):
        idx, dist, _ = run_nn_fat(host_lib, flat, flat, pts4, q, nrec)
        assert np.array_equal(dist, want["distance"][:, 0]) and np.array_equal(idx, want["index"][:, 0]), nrec


def test_slot_stack_spills_beyond_shared_slots(oracle, host_lib):
    """Queries far outside the cloud keep many far children pending: the slot stack goes past its four shared
    slots into the local spill, and the restore records keep the offsets right on the way back."""
    rng = np.random.default_rng(2)
    pts = rng.random((30_000, 3)).astype(np.float32)
    q = (rng.random((3_000, 3)) * 8 - 4).astype(np.float32)
    tree = oracle.OracleTree(pts, 2)
    flat = flat_nodes(tree.nodes, np.float32)
    pts4 = leaf_order_points(pts, tree.indices)
    want = tree.search_knn(q, 1)
    st = (C.c_ulonglong * 11)()
    host_lib.host_take_stack_stats(st)
    for nrec in (100, 103):
        idx, dist, _ = run_nn_fat(host_lib, flat, flat, pts4, q, nrec)
        assert np.array_equal(dist, want["distance"][:, 0]) and np.array_equal(idx, want["index"][:, 0])
    host_lib.host_take_stack_stats(st)
    assert st[3 + 7] > 0 or st[3 + 5] > 0 or st[3 + 6] > 0, "no query went deeper than four slots"


def test_tie_flag_is_complete(oracle, host_lib):
    """Every query whose minimum distance is attained by two or more points is flagged (brute force)."""
    pts, q = clouds("grid", 3_000, 1_500, 3, np.float32, seed=9)
    tree = oracle.OracleTree(pts, 10)
    flat = flat_nodes(tree.nodes, np.float32)
    fat, _ = fat_image(flat, 32)
    pts4 = leaf_order_points(pts, tree.indices)
    _, dist, tie = run_nn_fat(host_lib, fat, flat, pts4, q, 3)
    d = np.zeros((len(q), len(pts)), np.float32)
    for j in range(3):
        t = q[:, j:j + 1] - pts[None, :, j]
        d = d + t * t
    at_min = (d == d.min(axis=1, keepdims=True)).sum(axis=1)
    assert np.array_equal(dist, d.min(axis=1))
    assert np.all(tie[at_min > 1] == 1)
    assert np.all(tie[at_min == 1] == 0)


@pytest.mark.parametrize("k", [1, 4, 16])
def test_order_exact_traversal_matches_oracle(oracle, host_lib, k):
    """traverse_packed with the kernels' priming (first leaf for k = 1, bound for k > 1): index for index."""
    pts, q = clouds("lidar", 25_000, 3_000, 3, np.float32, seed=5)
    tree = oracle.OracleTree(pts, 10)
    flat = flat_nodes(tree.nodes, np.float32)
    pts4 = leaf_order_points(pts, tree.indices)
    fn = host_lib.host_knn_packed_f32
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    fn.restype = None
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.float32)
    fn(flat.ctypes.data, pts4.ctypes.data, q.ctypes.data, len(q), 3, k, len(pts), idx.ctypes.data, dist.ctypes.data)
    want = tree.search_knn(q, k)
    assert np.array_equal(dist, want["distance"]) and np.array_equal(idx, want["index"])


@pytest.mark.parametrize("sdim,leaf", [(3, 10), (2, 1), (3, 64)])
def test_box_thread_traversal_matches_oracle(oracle, host_lib, sdim, leaf):
    """traverse_box_thread (one thread per box): counts and indices in the reference's depth-first report order,
    whole cells reported as one run included; boxes of every size, some empty, some covering everything."""
    rng = np.random.default_rng(21)
    pts = rng.random((30_000, sdim)).astype(np.float32)
    tree = oracle.OracleTree(pts, leaf)
    flat = flat_nodes(tree.nodes, np.float32)
    pts4 = leaf_order_points(pts, tree.indices)
    c = rng.random((1_500, sdim)).astype(np.float32) * 1.2 - 0.1
    h = (rng.random((1_500, 1)) ** 3 * 0.5).astype(np.float32)
    mins, maxs = np.ascontiguousarray(c - h), np.ascontiguousarray(c + h)
    mins[0], maxs[0] = -1.0, 2.0            # everything
    mins[1], maxs[1] = 5.0, 6.0             # nothing
    mins[2], maxs[2] = pts[7], pts[7]       # a single point, bounds inclusive
    w_offs, w_flat = tree.search_box(mins, maxs)
    fn = host_lib.host_box_f32
    fn.restype = C.c_ulonglong
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                   C.c_void_p, C.c_void_p]
    root_box = np.ascontiguousarray(tree.root_box, dtype=np.float32).ravel()
    idx = np.ascontiguousarray(tree.indices)
    offs = np.zeros(len(mins) + 1, np.uint64)
    total = fn(flat.ctypes.data, pts4.ctypes.data, idx.ctypes.data, root_box.ctypes.data, mins.ctypes.data,
               maxs.ctypes.data, len(mins), sdim, offs.ctypes.data, None)
    assert total == int(w_offs[-1]) and np.array_equal(offs, w_offs)
    out = np.empty(int(total), np.int32)
    fn(flat.ctypes.data, pts4.ctypes.data, idx.ctypes.data, root_box.ctypes.data, mins.ctypes.data, maxs.ctypes.data,
       len(mins), sdim, offs.ctypes.data, out.ctypes.data)
    assert np.array_equal(out, w_flat)
    assert offs[1] - offs[0] == len(pts) and offs[2] == offs[1] and offs[3] - offs[2] >= 1

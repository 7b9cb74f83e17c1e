"""kd_forest traversal core (pico_tree_b200/csrc/forest.cuh) checked on the CPU.

The header is `__host__ __device__`; tests/cpp/forest_host.cpp compiles the same source for the host and drives
it the way kd_forest::search_nearest does. Here it runs over the flat node arrays of the oracle's forest (built
from the Householder vectors stored in the fixtures) and must reproduce the outputs of the UNMODIFIED reference
kd_forest (tests/golden/forest/*.npz) bit for bit. The kernel / handle / C-ABI around the core do not exist yet,
so there is no `-m gpu` counterpart of this file."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "_bin", "libforest_host.so")


def _forest_files():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "forest", "*.npz")))


@pytest.fixture(scope="module")
def host_lib():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s", LIB], check=True)
    return C.CDLL(LIB)


def flat_nodes(nodes, dtype):
    """oracle nodes (allocation order = pre-order, explicit links) -> pico_b200_node_f32 / _f64 records."""
    branch = nodes["split_dim"] >= 0
    assert np.array_equal(nodes["left"][branch], np.nonzero(branch)[0] + 1)  # left child = i + 1
    if dtype == np.float32:
        out = np.zeros(len(nodes), dtype=np.dtype([("a", "<u4"), ("b", "<u4"), ("right", "<u4"), ("split_dim", "<u4")]))
        out["a"] = np.where(branch, nodes["left_max"].view(np.uint32), nodes["begin"].astype(np.uint32))
        out["b"] = np.where(branch, nodes["right_min"].view(np.uint32), nodes["end"].astype(np.uint32))
    else:
        out = np.zeros(len(nodes), dtype=np.dtype([("a", "<u8"), ("b", "<u8"), ("right", "<u4"), ("split_dim", "<u4"),
                                                   ("pad", "<u8")]))
        out["a"] = np.where(branch, nodes["left_max"].view(np.uint64), nodes["begin"].astype(np.int64).view(np.uint64))
        out["b"] = np.where(branch, nodes["right_min"].view(np.uint64), nodes["end"].astype(np.int64).view(np.uint64))
    out["right"] = np.where(branch, nodes["right"], -1).astype(np.int64).astype(np.uint32)
    out["split_dim"] = np.where(branch, nodes["split_dim"], -1).astype(np.int64).astype(np.uint32)
    return out


@pytest.mark.parametrize("path", _forest_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_forest_core_matches_reference_fixture(oracle, host_lib, path):
    g = np.load(path)
    pts, q, rot = g["pts"], g["q"], np.ascontiguousarray(g["rotations"])
    dtype = pts.dtype
    f = oracle.OracleForest(pts, rot, int(g["max_leaf_size"]))
    n_trees, sdim = rot.shape
    keep, height = [], 0
    arrs = {"nodes": [], "outer": [], "indices": [], "points": []}
    for t in range(n_trees):
        tv = f.tree(t)
        height = max(height, tv.height)
        for name, a in (("nodes", flat_nodes(tv.nodes, dtype)), ("outer", np.ascontiguousarray(tv.outer_bounds)),
                        ("indices", tv.indices), ("points", f.rotated_space(t))):
            keep.append(a)
            arrs[name].append(a.ctypes.data)
    ptrs = {k: (C.c_void_p * n_trees)(*v) for k, v in arrs.items()}
    nb = np.dtype([("index", "<i4"), ("distance", "<f4")]) if dtype == np.float32 else \
        np.dtype([("index", "<i4"), ("distance", "<f8")], align=True)
    fn = host_lib.forest_host_knn_f32 if dtype == np.float32 else host_lib.forest_host_knn_f64
    fn.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                   C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32, C.c_void_p]
    for k, ml in g["searches"]:
        out = np.zeros((len(q), int(k)), dtype=nb)
        overflow = fn(n_trees, ptrs["nodes"], ptrs["outer"], ptrs["indices"], ptrs["points"], sdim, height,
                      rot.ctypes.data, q.ctypes.data, len(q), int(k), int(ml), 1 << 16, out.ctypes.data)
        assert overflow == 0
        assert np.array_equal(out["index"], g[f"index_k{k}_m{ml}"]), (k, ml)
        assert np.array_equal(out["distance"], g[f"distance_k{k}_m{ml}"]), (k, ml)
    # a queue that is too small is reported, never silently truncated
    out = np.zeros((len(q), 1), dtype=nb)
    assert fn(n_trees, ptrs["nodes"], ptrs["outer"], ptrs["indices"], ptrs["points"], sdim, height, rot.ctypes.data,
              q.ctypes.data, len(q), 1, 1 << 30, 2, out.ctypes.data) == 1

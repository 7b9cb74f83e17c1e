"""ctypes binding of libpico_b200.so (the C-ABI in include/pico_b200.h).

The shared library is the product; this module only loads it. There is no
Python or CPU fallback: if the library is missing or no B200 is visible the
calls raise.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpico_b200.so")

F32, F64 = 0, 1
FLAG_DEVICE_POINTERS = 1 << 0
FLAG_NO_REORDER = 1 << 1
FLAG_SORT_RESULTS = 1 << 2
FLAG_WARP_PER_QUERY = 1 << 3
FLAG_ASYNC = 1 << 4

EXPORTS = [
    "pico_b200_last_error", "pico_b200_abi_version", "pico_b200_device_count", "pico_b200_tree_create",
    "pico_b200_tree_create_from_nodes", "pico_b200_tree_destroy", "pico_b200_tree_info_get", "pico_b200_tree_export",
    "pico_b200_tree_export_outer_bounds",
    "pico_b200_knn", "pico_b200_radius", "pico_b200_box", "pico_b200_tree_broadcast",
    "pico_b200_tree_serialize_size", "pico_b200_tree_serialize", "pico_b200_tree_deserialize", "pico_b200_free",
    "pico_b200_free_device",
    "pico_b200_tree_save_size", "pico_b200_tree_save", "pico_b200_tree_load", "pico_b200_set_stream",
    "pico_b200_profile_begin", "pico_b200_profile_end", "pico_b200_profile_leaf_scan",
    "pico_b200_forest_create", "pico_b200_forest_destroy", "pico_b200_forest_info_get", "pico_b200_forest_rotations",
    "pico_b200_forest_tree", "pico_b200_forest_knn", "pico_b200_tree_order_state",
]


class TreeInfo(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("sdim", C.c_uint64), ("n_nodes", C.c_uint64), ("n_leaves", C.c_uint64),
                ("height", C.c_uint64), ("scalar", C.c_int32), ("metric", C.c_int32), ("device", C.c_int32),
                ("reserved_", C.c_int32), ("build_ms", C.c_double), ("device_bytes", C.c_uint64)]


class SearchStats(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("reorder_ms", C.c_double), ("kernel_ms", C.c_double),
                ("d2h_ms", C.c_double), ("kernel_launches", C.c_uint64)]


class ForestInfo(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("sdim", C.c_uint64), ("n_trees", C.c_uint64), ("max_leaf_size", C.c_uint64),
                ("height", C.c_uint64), ("scalar", C.c_int32), ("device", C.c_int32), ("build_ms", C.c_double),
                ("device_bytes", C.c_uint64)]


class PicoB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pico_b200 error {code}: {msg}")
        self.code = code


def build(verbose=False):
    """Compile the CUDA sources for sm_100a into pico_tree_b200/libpico_b200.so."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"] + ([] if verbose else ["-s"]), check=True)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(pico_tree_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u32, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_double
    L.pico_b200_last_error.restype = C.c_char_p
    L.pico_b200_device_count.argtypes = [C.POINTER(C.c_int)]
    L.pico_b200_tree_create.argtypes = [vp, sz, sz, sz, i32, i32, i32, i32, sz, vp, vp, i32, C.POINTER(vp)]
    L.pico_b200_tree_create_from_nodes.argtypes = [vp, sz, sz, sz, i32, i32, vp, sz, vp, vp, vp, i32, C.POINTER(vp)]
    L.pico_b200_tree_export_outer_bounds.argtypes = [vp, vp]
    L.pico_b200_tree_destroy.argtypes = [vp]
    L.pico_b200_tree_destroy.restype = None
    L.pico_b200_tree_info_get.argtypes = [vp, C.POINTER(TreeInfo)]
    L.pico_b200_tree_export.argtypes = [vp, vp, vp, vp]
    L.pico_b200_knn.argtypes = [vp, vp, sz, sz, sz, dbl, vp, u32, C.POINTER(SearchStats)]
    L.pico_b200_radius.argtypes = [vp, vp, sz, sz, dbl, dbl, vp, C.POINTER(vp), u32, C.POINTER(SearchStats)]
    L.pico_b200_box.argtypes = [vp, vp, vp, sz, sz, vp, C.POINTER(vp), u32, C.POINTER(SearchStats)]
    L.pico_b200_tree_broadcast.argtypes = [C.POINTER(vp), vp, i32, i32, i32]
    L.pico_b200_tree_serialize_size.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.pico_b200_tree_serialize.argtypes = [vp, vp, i32]
    L.pico_b200_tree_deserialize.argtypes = [vp, C.c_uint64, i32, i32, C.POINTER(vp)]
    L.pico_b200_free.argtypes = [vp]
    L.pico_b200_free.restype = None
    L.pico_b200_free_device.argtypes = [vp]
    L.pico_b200_free_device.restype = None
    L.pico_b200_tree_save_size.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.pico_b200_tree_save.argtypes = [vp, vp]
    L.pico_b200_tree_load.argtypes = [vp, sz, sz, sz, i32, i32, vp, C.c_uint64, i32, C.POINTER(vp),
                                      C.POINTER(C.c_uint64)]
    L.pico_b200_set_stream.argtypes = [vp]
    L.pico_b200_profile_begin.argtypes = []
    L.pico_b200_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.pico_b200_profile_leaf_scan.argtypes = [vp, vp, sz, sz, vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                              C.POINTER(C.c_uint64)]
    L.pico_b200_tree_order_state.argtypes = [vp, C.POINTER(C.c_int)]
    L.pico_b200_forest_create.argtypes = [vp, sz, sz, sz, i32, sz, vp, sz, i32, C.POINTER(vp)]
    L.pico_b200_forest_destroy.argtypes = [vp]
    L.pico_b200_forest_destroy.restype = None
    L.pico_b200_forest_info_get.argtypes = [vp, C.POINTER(ForestInfo)]
    L.pico_b200_forest_rotations.argtypes = [vp, vp]
    L.pico_b200_forest_tree.argtypes = [vp, sz, C.POINTER(vp)]
    L.pico_b200_forest_knn.argtypes = [vp, vp, sz, sz, sz, sz, vp, u32, C.POINTER(SearchStats)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise PicoB200Error(rc, lib().pico_b200_last_error().decode("utf-8", "replace"))

"""ctypes front-end of the CPU oracle. TEST INFRASTRUCTURE ONLY.

Two back-ends behind one small interface:
  * ``OracleTree``  — the plain-C restatement (oracle/pico_oracle.c).
  * ``RefTree``     — the unmodified reference headers compiled into
                       oracle/_ref/libpico_ref.so (prebuilt in the dev container;
                       travels to the GPU box as a binary).
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this module; pico_tree_b200/ never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = {"l1": 0, "l2_squared": 1, "lpinf": 2, "lninf": 3, "so2": 4, "se2_squared": 5}
TOPOLOGICAL = ("so2", "se2_squared")
RULES = {"sliding_midpoint": 0, "midpoint": 1, "median": 2}
STOPS = {"max_leaf_size": 0, "max_leaf_depth": 1}

NEIGHBOR_F32 = np.dtype([("index", "<i4"), ("distance", "<f4")])
NEIGHBOR_F64 = np.dtype([("index", "<i4"), ("distance", "<f8")], align=True)
NODE_F32 = np.dtype([("left_max", "<f4"), ("right_min", "<f4"), ("split_dim", "<i4"), ("begin", "<i4"),
                     ("end", "<i4"), ("left", "<i4"), ("right", "<i4")])
NODE_F64 = np.dtype([("left_max", "<f8"), ("right_min", "<f8"), ("split_dim", "<i4"), ("begin", "<i4"),
                     ("end", "<i4"), ("left", "<i4"), ("right", "<i4")], align=True)


def build(quiet=True):
    """(Re)build libpico_oracle.so and, where /root/reference exists, _ref/."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _sfx(dtype):
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libpico_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        for s in ("f32", "f64"):
            sc = C.c_float if s == "f32" else C.c_double
            getattr(L, f"po_build_{s}").restype = C.c_void_p
            getattr(L, f"po_build_{s}").argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                                    C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
            getattr(L, f"po_free_{s}").argtypes = [C.c_void_p]
            for f in ("po_num_nodes", "po_height"):
                getattr(L, f"{f}_{s}").restype = C.c_size_t
                getattr(L, f"{f}_{s}").argtypes = [C.c_void_p]
            for f in ("po_nodes", "po_indices", "po_root_box", "po_outer_bounds"):
                getattr(L, f"{f}_{s}").restype = C.c_void_p
                getattr(L, f"{f}_{s}").argtypes = [C.c_void_p]
            getattr(L, f"po_knn_batch_{s}").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                        C.c_double, C.c_void_p, C.c_int, C.c_void_p]
            getattr(L, f"po_radius_batch_{s}").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, sc,
                                                           C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
            getattr(L, f"po_box_batch_{s}").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                                        C.c_void_p, C.POINTER(C.c_void_p)]
            getattr(L, f"po_radius_counters_{s}").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, sc,
                                                              C.c_double, C.c_void_p]
            getattr(L, f"po_box_counters_{s}").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                                           C.c_void_p]
            getattr(L, f"po_splitter_once_{s}").restype = C.c_size_t
            getattr(L, f"po_splitter_once_{s}").argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                                            C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                                            C.POINTER(C.c_size_t), C.c_void_p]
            getattr(L, f"po_forest_build_{s}").restype = C.c_void_p
            getattr(L, f"po_forest_build_{s}").argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                                           C.c_void_p, C.c_size_t]
            getattr(L, f"po_forest_free_{s}").argtypes = [C.c_void_p]
            getattr(L, f"po_forest_tree_{s}").restype = C.c_void_p
            getattr(L, f"po_forest_tree_{s}").argtypes = [C.c_void_p, C.c_size_t]
            getattr(L, f"po_forest_space_{s}").restype = C.c_void_p
            getattr(L, f"po_forest_space_{s}").argtypes = [C.c_void_p, C.c_size_t]
            getattr(L, f"po_forest_knn_batch_{s}").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                                               C.c_size_t, C.c_size_t, C.c_void_p, C.c_int]
        L.po_free_buffer.argtypes = [C.c_void_p]
        L.po_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads():
    """Host threads the batch loops may use: the CPUs this process is allowed on. (Not omp_get_max_threads():
    torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently make the all-cores reference arm
    single-threaded; the loops take their thread count explicitly, `num_threads(threads)`.)"""
    if int(lib().po_max_threads()) < 1:
        return 1
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _take(ptr, count, dtype, free):
    """Copy `count` records out of a malloc'd buffer and free it."""
    if count == 0:
        out = np.empty(0, dtype=dtype)
    else:
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=dtype).copy()
    if ptr.value:
        free(ptr)
    return out


class _Base:
    def _prep(self, pts):
        pts = np.ascontiguousarray(pts)
        assert pts.ndim == 2 and pts.dtype in (np.float32, np.float64)
        self.pts = pts
        self.n, self.sdim = pts.shape
        self.dtype = pts.dtype
        self.nb_dtype = NEIGHBOR_F32 if pts.dtype == np.float32 else NEIGHBOR_F64

    def _q(self, q):
        q = np.ascontiguousarray(q, dtype=self.dtype)
        assert q.ndim == 2 and q.shape[1] == self.sdim
        return q


class OracleTree(_Base):
    """Plain-C restatement (oracle/pico_oracle.c)."""

    def __init__(self, pts, max_leaf_size=10, metric="l2_squared", rule="sliding_midpoint", stop="max_leaf_size",
                 bounds=None):
        self._prep(pts)
        self.s = _sfx(self.dtype)
        self.metric = metric
        bmin = bmax = None
        if bounds is not None:
            bmin = np.ascontiguousarray(bounds[0], dtype=self.dtype)
            bmax = np.ascontiguousarray(bounds[1], dtype=self.dtype)
        self._h = getattr(lib(), f"po_build_{self.s}")(_ptr(self.pts), self.n, self.sdim, self.sdim, METRICS[metric],
                                                       RULES[rule], STOPS[stop], int(max_leaf_size), _ptr(bmin),
                                                       _ptr(bmax))

    def __del__(self):
        if getattr(self, "_h", None):
            getattr(lib(), f"po_free_{self.s}")(self._h)
            self._h = None

    def _f(self, name):
        return getattr(lib(), f"{name}_{self.s}")

    @property
    def num_nodes(self):
        return int(self._f("po_num_nodes")(self._h))

    @property
    def height(self):
        return int(self._f("po_height")(self._h))

    @property
    def nodes(self):
        dt = NODE_F32 if self.s == "f32" else NODE_F64
        m = self.num_nodes
        buf = (C.c_char * (m * dt.itemsize)).from_address(self._f("po_nodes")(self._h))
        return np.frombuffer(buf, dtype=dt).copy()

    @property
    def outer_bounds(self):
        """(n_nodes, 2): left_min, right_max of every branch (kd_tree_branch_double)."""
        m = self.num_nodes
        buf = (C.c_char * (m * 2 * self.dtype.itemsize)).from_address(self._f("po_outer_bounds")(self._h))
        return np.frombuffer(buf, dtype=self.dtype).copy().reshape(m, 2)

    @property
    def indices(self):
        buf = (C.c_char * (self.n * 4)).from_address(self._f("po_indices")(self._h))
        return np.frombuffer(buf, dtype=np.int32).copy()

    @property
    def root_box(self):
        buf = (C.c_char * (2 * self.sdim * self.dtype.itemsize)).from_address(self._f("po_root_box")(self._h))
        return np.frombuffer(buf, dtype=self.dtype).copy().reshape(2, self.sdim)

    def search_knn(self, q, k, e=0.0, threads=1, counters=False):
        q = self._q(q)
        k = min(int(k), self.n)
        out = np.empty((len(q), k), dtype=self.nb_dtype)
        cnt = np.zeros(3, dtype=np.uint64) if counters else None
        self._f("po_knn_batch")(self._h, _ptr(q), len(q), self.sdim, k, float(e), _ptr(out), int(threads), _ptr(cnt))
        return (out, cnt) if counters else out

    def search_radius(self, q, radius, e=0.0, sort=False):
        q = self._q(q)
        offs = np.zeros(len(q) + 1, dtype=np.uint64)
        p = C.c_void_p()
        self._f("po_radius_batch")(self._h, _ptr(q), len(q), self.sdim, float(radius), float(e), int(sort), _ptr(offs),
                                   C.byref(p))
        return offs, _take(p, int(offs[-1]), self.nb_dtype, lib().po_free_buffer)

    def radius_counters(self, q, radius, e=0.0):
        """{branch nodes, leaves, points tested, hits} summed over the batch (SURVEY.md §8d byte model)."""
        q = self._q(q)
        c = np.zeros(4, dtype=np.uint64)
        self._f("po_radius_counters")(self._h, _ptr(q), len(q), self.sdim, float(radius), float(e), _ptr(c))
        return c

    def box_counters(self, mins, maxs):
        mins, maxs = self._q(mins), self._q(maxs)
        c = np.zeros(4, dtype=np.uint64)
        self._f("po_box_counters")(self._h, _ptr(mins), _ptr(maxs), len(mins), self.sdim, _ptr(c))
        return c

    def search_box(self, mins, maxs):
        mins, maxs = self._q(mins), self._q(maxs)
        offs = np.zeros(len(mins) + 1, dtype=np.uint64)
        p = C.c_void_p()
        self._f("po_box_batch")(self._h, _ptr(mins), _ptr(maxs), len(mins), self.sdim, _ptr(offs), C.byref(p))
        return offs, _take(p, int(offs[-1]), np.int32, lib().po_free_buffer)


def splitter_once(pts, rule, idx, begin, end, box_min, box_max):
    """One splitter call (the KATs of test/pico_tree/kd_tree_builder_test.cpp)."""
    pts = np.ascontiguousarray(pts)
    s = _sfx(pts.dtype)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    bmin = np.ascontiguousarray(box_min, dtype=pts.dtype)
    bmax = np.ascontiguousarray(box_max, dtype=pts.dtype)
    sd = C.c_size_t()
    sv = np.zeros(1, dtype=pts.dtype)
    split = getattr(lib(), f"po_splitter_once_{s}")(_ptr(pts), pts.shape[1], pts.shape[1], RULES[rule], _ptr(idx),
                                                    begin, end, _ptr(bmin), _ptr(bmax), C.byref(sd), _ptr(sv))
    return int(split), int(sd.value), sv[0], idx


# ---------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpico_ref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libpico_ref.so"))
        L.ref_build.restype = C.c_void_p
        L.ref_build.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_size_t, C.c_void_p, C.c_void_p]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_save.restype = C.c_void_p
        L.ref_save.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.ref_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_double, C.c_void_p, C.c_int]
        L.ref_radius.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_void_p,
                                 C.POINTER(C.c_void_p)]
        L.ref_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]
        L.ref_radius_mt.argtypes = L.ref_radius.argtypes + [C.c_int]
        L.ref_box_mt.argtypes = L.ref_box.argtypes + [C.c_int]
        L.ref_free_buffer.argtypes = [C.c_void_p]
        _ref = L
    return _ref


def parse_saved_tree(blob, scalar_dtype, topological=False):
    """Decode a kd_tree::save stream (internal/kd_tree_data.hpp:43-58,89-135) into
    (sdim, indices, root_box[2,sdim], nodes) with nodes in the oracle's NODE_* dtype. With
    ``topological`` the branches are kd_tree_branch_double records (kd_tree_node.hpp:52-59) and a
    fifth value is returned: outer[n_nodes, 2] = (left_min, right_max)."""
    scalar_dtype = np.dtype(scalar_dtype)
    mv = memoryview(blob)
    pos = 0
    sdim = int(np.frombuffer(mv[pos:pos + 8], dtype="<u8")[0]); pos += 8
    n = int(np.frombuffer(mv[pos:pos + 8], dtype="<u8")[0]); pos += 8
    indices = np.frombuffer(mv[pos:pos + 4 * n], dtype="<i4").copy(); pos += 4 * n
    box = np.frombuffer(mv[pos:pos + 2 * sdim * scalar_dtype.itemsize], dtype=scalar_dtype).copy().reshape(2, sdim)
    pos += 2 * sdim * scalar_dtype.itemsize
    f32 = scalar_dtype == np.float32
    if topological:
        branch_dt = np.dtype([("split_dim", "<i4"), ("left_min", scalar_dtype), ("left_max", scalar_dtype),
                              ("right_min", scalar_dtype), ("right_max", scalar_dtype)], align=True)
    else:
        branch_dt = np.dtype([("split_dim", "<i4"), ("left_max", scalar_dtype), ("right_min", scalar_dtype)],
                             align=True)
    outer = []
    out_dt = NODE_F32 if f32 else NODE_F64
    raw = np.frombuffer(mv[pos:], dtype=np.uint8)
    # Pre-order stream: 1-byte is_leaf flag + 8-byte leaf or sizeof(branch) bytes.
    bs = branch_dt.itemsize
    nodes = []
    p = 0
    while p < len(raw):
        is_leaf = raw[p]; p += 1
        if is_leaf:
            b, e = np.frombuffer(raw[p:p + 8].tobytes(), dtype="<i4"); p += 8
            nodes.append((0, 0, -1, b, e, -1, -1))
            outer.append((0, 0))
        else:
            br = np.frombuffer(raw[p:p + bs].tobytes(), dtype=branch_dt)[0]; p += bs
            nodes.append((br["left_max"], br["right_min"], br["split_dim"], 0, 0, 0, 0))
            outer.append((br["left_min"], br["right_max"]) if topological else (0, 0))
    arr = np.array(nodes, dtype=out_dt)
    # Link children: pre-order, left = self+1, right = first node after the left subtree.
    stack = []
    for i in range(len(arr)):
        while stack and stack[-1][1] == 2:
            stack.pop()
        if stack:
            parent, seen = stack[-1]
            if seen == 0:
                arr["left"][parent] = i
            else:
                arr["right"][parent] = i
            stack[-1][1] += 1
        if arr["split_dim"][i] >= 0:
            stack.append([i, 0])
    if topological:
        return sdim, indices, box, arr, np.array(outer, dtype=scalar_dtype).reshape(-1, 2)
    return sdim, indices, box, arr


class RefTree(_Base):
    """The unmodified reference (pico_tree::kd_tree<space_map<point_map<T const,D>>>)."""

    def __init__(self, pts, max_leaf_size=10, metric="l2_squared", rule="sliding_midpoint", stop="max_leaf_size",
                 bounds=None, force_dynamic=False):
        self._prep(pts)
        self.metric = metric
        bmin = bmax = None
        if bounds is not None:
            bmin = np.ascontiguousarray(bounds[0], dtype=self.dtype)
            bmax = np.ascontiguousarray(bounds[1], dtype=self.dtype)
        self._h = ref_lib().ref_build(_ptr(self.pts), self.n, self.sdim, int(self.dtype == np.float64),
                                      int(force_dynamic), METRICS[metric], RULES[rule], STOPS[stop],
                                      int(max_leaf_size), _ptr(bmin), _ptr(bmax))

    def __del__(self):
        if getattr(self, "_h", None):
            ref_lib().ref_free(self._h)
            self._h = None

    def saved(self):
        size = C.c_size_t()
        p = C.c_void_p(ref_lib().ref_save(self._h, C.byref(size)))
        blob = C.string_at(p, size.value)
        ref_lib().ref_free_buffer(p)
        return blob

    def structure(self):
        return parse_saved_tree(self.saved(), self.dtype, self.metric in TOPOLOGICAL)

    def search_knn(self, q, k, e=0.0, threads=1):
        q = self._q(q)
        k = min(int(k), self.n)
        out = np.empty((len(q), k), dtype=self.nb_dtype)
        ref_lib().ref_knn(self._h, _ptr(q), len(q), k, float(e), _ptr(out), int(threads))
        return out

    def search_radius(self, q, radius, e=0.0, sort=False, threads=1):
        q = self._q(q)
        offs = np.zeros(len(q) + 1, dtype=np.uint64)
        p = C.c_void_p()
        ref_lib().ref_radius_mt(self._h, _ptr(q), len(q), float(radius), float(e), int(sort), _ptr(offs), C.byref(p),
                                int(threads))
        return offs, _take(p, int(offs[-1]), self.nb_dtype, ref_lib().ref_free_buffer)

    def search_box(self, mins, maxs, threads=1):
        mins, maxs = self._q(mins), self._q(maxs)
        offs = np.zeros(len(mins) + 1, dtype=np.uint64)
        p = C.c_void_p()
        ref_lib().ref_box_mt(self._h, _ptr(mins), _ptr(maxs), len(mins), _ptr(offs), C.byref(p), int(threads))
        return offs, _take(p, int(offs[-1]), np.int32, ref_lib().ref_free_buffer)


# ---------------------------------------------------------------------------------------------- kd_forest
# SURVEY.md §8 f4 (examples/pico_understory/pico_understory/kd_forest.hpp): oracle first, product next round.
_ref_forest_lib = None


def ref_forest_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpico_ref_forest.so"))


def ref_forest_lib():
    global _ref_forest_lib
    if _ref_forest_lib is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libpico_ref_forest.so"))
        L.ref_forest_create.restype = C.c_void_p
        L.ref_forest_create.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t]
        L.ref_forest_free.argtypes = [C.c_void_p]
        L.ref_forest_size.restype = C.c_size_t
        L.ref_forest_size.argtypes = [C.c_void_p]
        L.ref_forest_rotations.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_forest_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int]
        _ref_forest_lib = L
    return _ref_forest_lib


class RefForest(_Base):
    """The unmodified reference kd_forest<space_map<point_map<T const, dynamic_extent>>> (its Householder vectors
    come from std::random_device; `rotations` returns the ones this instance drew)."""

    def __init__(self, pts, max_leaf_size=10, forest_size=4):
        self._prep(pts)
        self._h = ref_forest_lib().ref_forest_create(_ptr(self.pts), self.n, self.sdim, int(self.dtype == np.float64),
                                                     int(max_leaf_size), int(forest_size))

    def __del__(self):
        if getattr(self, "_h", None):
            ref_forest_lib().ref_forest_free(self._h)
            self._h = None

    @property
    def rotations(self):
        out = np.empty((int(ref_forest_lib().ref_forest_size(self._h)), self.sdim), dtype=self.dtype)
        ref_forest_lib().ref_forest_rotations(self._h, _ptr(out))
        return out

    def search_knn(self, q, k, max_leaves_visited, threads=1):
        q = self._q(q)
        out = np.empty((len(q), int(k)), dtype=self.nb_dtype)
        ref_forest_lib().ref_forest_knn(self._h, _ptr(q), len(q), int(k), int(max_leaves_visited), _ptr(out),
                                        int(threads))
        return out


class OracleForest(_Base):
    """Plain-C restatement of kd_forest (po_forest_*), Householder vectors given."""

    def __init__(self, pts, rotations, max_leaf_size=10):
        self._prep(pts)
        self.s = _sfx(self.dtype)
        self.rot = np.ascontiguousarray(rotations, dtype=self.dtype).reshape(-1, self.sdim)
        self._h = getattr(lib(), f"po_forest_build_{self.s}")(_ptr(self.pts), self.n, self.sdim, self.sdim,
                                                              int(max_leaf_size), _ptr(self.rot), len(self.rot))

    def __del__(self):
        if getattr(self, "_h", None):
            getattr(lib(), f"po_forest_free_{self.s}")(self._h)
            self._h = None

    def tree(self, t):
        """Tree t of the forest as an OracleTree view (nodes, outer_bounds, indices, height); owned by the forest."""
        return _ForestTreeView(self, int(t))

    def rotated_space(self, t):
        p = getattr(lib(), f"po_forest_space_{self.s}")(self._h, int(t))
        buf = (C.c_char * (self.n * self.sdim * self.dtype.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=self.dtype).copy().reshape(self.n, self.sdim)

    def search_knn(self, q, k, max_leaves_visited, threads=1):
        q = self._q(q)
        out = np.empty((len(q), int(k)), dtype=self.nb_dtype)
        getattr(lib(), f"po_forest_knn_batch_{self.s}")(self._h, _ptr(q), len(q), self.sdim, int(k),
                                                        int(max_leaves_visited), _ptr(out), int(threads))
        return out


class _ForestTreeView(OracleTree):
    def __init__(self, forest, t):  # noqa: super().__init__ not called: nothing is built here
        self._forest = forest  # keeps the owner alive
        self.s, self.dtype, self.nb_dtype = forest.s, forest.dtype, forest.nb_dtype
        self.n, self.sdim, self.metric = forest.n, forest.sdim, "l2_squared"
        self._h = getattr(lib(), f"po_forest_tree_{self.s}")(forest._h, t)

    def __del__(self):
        self._h = None

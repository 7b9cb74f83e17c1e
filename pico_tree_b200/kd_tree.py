"""Host-side mirror of the reference's Python API (``pico_tree.KdTree``).

Same names, argument meaning and error behaviour as the pybind11 module of the
reference (src/pyco_tree/pico_tree/_pyco_tree/def_kd_tree.cpp:14-195,
_pyco_tree/kd_tree.hpp:92-616, darray.hpp, py_array_map.hpp:36-62); the three
batch loops of ``kd_tree_impl`` (kd_tree.hpp:117-268) are each ONE call into
libpico_b200.so. Nothing here computes distances or walks a tree.
"""
import ctypes as C
import enum
import struct
import weakref

import numpy as np

from . import _lib

__all__ = ["Metric", "KdTree", "DArray", "load_kd_tree", "save_kd_tree"]


_CUDA_STREAM_LEGACY = 0x1  # driver_types.h: #define cudaStreamLegacy ((cudaStream_t)0x1)


class Metric(enum.Enum):
    """metric_t of the binding (def_kd_tree.cpp:14-17) plus the fourth euclidean metric of
    metric.hpp (metric_lninf), which the reference binding does not expose."""
    L1 = 0
    L2Squared = 1
    LPInf = 2
    LNInf = 3
    SO2 = 4          # the circle [0, 1), sdim 1 (metric.hpp:197-221; C++ API only in the reference)
    SE2Squared = 5   # R2 x S1, sdim 3 (metric.hpp:223-257)


class Rule(enum.Enum):
    """splitter rules, internal/kd_tree_builder.hpp:35-75 (C++ API only in the reference)."""
    SlidingMidpointMaxSide = 0
    MidpointMaxSide = 1
    MedianMaxSide = 2


_NEIGHBOR = {
    np.dtype(np.float32): np.dtype([("index", "<i4"), ("distance", "<f4")]),
    np.dtype(np.float64): np.dtype({"names": ["index", "distance"], "formats": ["<i4", "<f8"], "offsets": [0, 8],
                                    "itemsize": 16}),
}
_PKD_SIGNATURE = b"\x89PKD"  # _pyco_tree/kd_tree.hpp:548-550
_PKD_VERSION = 1


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _as_points(arr, sdim=None, dtype=None, what="array"):
    """py_array_map.hpp:36-62: ndim == 2, contiguous, inner dimension == sdim.
    Returns (C-contiguous (n, sdim) view, row_major)."""
    if not isinstance(arr, np.ndarray):
        raise TypeError(f"{what} must be a numpy.ndarray")
    if arr.ndim != 2:
        raise ValueError("array ndim not 2")
    if arr.flags["C_CONTIGUOUS"]:
        row_major, view = True, arr
    elif arr.flags["F_CONTIGUOUS"]:
        row_major, view = False, arr.T  # (sdim, n) column-major == (n, sdim) row-major
    else:
        raise ValueError("array not contiguous")
    if dtype is not None and view.dtype != dtype:
        raise ValueError("array dtype not " + ("float32" if dtype == np.float32 else "float64"))
    if view.dtype not in (np.float32, np.float64):
        raise ValueError("array dtype not float32 or float64")
    if sdim is not None and view.shape[1] != sdim:
        raise ValueError("incompatible kd_tree sdim and array inner stride")
    return view, row_major


class DArray:
    """Ragged result container: a list of ndarray views over one flat buffer (the reference's
    ``DArray`` wraps a std::vector<std::vector<T>>, darray.hpp). Supports len(), indexing,
    negative indices, slices, iteration, ``dtype`` and truthiness like the reference
    (def_darray.cpp:14-66, test/pyco_tree/kd_tree_test.py:151-219)."""

    def __init__(self, dtype=None, _flat=None, _offsets=None, _items=None):
        if isinstance(dtype, KdTree):
            raise TypeError("pass a dtype, e.g. tree.dtype_neighbor")
        self._dtype = np.dtype(dtype) if dtype is not None else None
        self._flat = _flat if _flat is not None else (np.empty(0, self._dtype) if self._dtype else None)
        self._offsets = _offsets if _offsets is not None else np.zeros(1, np.uint64)
        self._items = _items  # explicit list of views (result of slicing)

    @property
    def dtype(self):
        return self._dtype

    def _view(self, i):
        if self._items is not None:
            return self._items[i]
        b, e = int(self._offsets[i]), int(self._offsets[i + 1])
        return self._flat[b:e]

    def __len__(self):
        return len(self._items) if self._items is not None else len(self._offsets) - 1

    def __bool__(self):
        return len(self) > 0

    def __getitem__(self, i):
        n = len(self)
        if isinstance(i, slice):
            return DArray(self._dtype, _items=[self._view(j) for j in range(*i.indices(n))])
        if i < 0:
            i += n
        if i < 0 or i >= n:
            raise IndexError("DArray index out of range")
        return self._view(i)

    def __iter__(self):
        return (self._view(i) for i in range(len(self)))

    def _assign(self, offsets, flat):
        """Take new results. A fresh DArray adopts the library's buffer without a copy (results can be gigabytes).
        A DArray the caller passed in again is a request to re-use its memory, like the reference re-uses its
        vectors' storage (test/pyco_tree/kd_tree_test.py:107-118 compares addresses): the results are copied
        into the old buffer when they fit."""
        self._items = None
        if self._flat is not None and self._flat.size and self._flat.dtype == flat.dtype and self._flat.size >= flat.size:
            self._flat[:flat.size] = flat
        else:
            self._flat = flat
        self._offsets = offsets


class KdTree:
    """``KdTree(pts, metric, max_leaf_size)`` — def_kd_tree.cpp:19-29. Extra keyword
    arguments reach the parts of the C++ constructor the reference binding hides
    (kd_tree.hpp:72-88): ``rule``, ``max_leaf_depth``, ``bounds`` and the CUDA ``device``."""

    def __init__(self, pts, metric=Metric.L2Squared, max_leaf_size=10, *, rule=Rule.SlidingMidpointMaxSide,
                 max_leaf_depth=None, bounds=None, device=0, _stream=None):
        view, row_major = _as_points(pts, what="pts")
        if not isinstance(metric, Metric):
            raise TypeError("metric must be a pico_tree_b200.Metric")
        self._pts = pts  # keep alive, like py::keep_alive<1, 2>
        self._view = view
        self._row_major = row_major
        self._metric = metric
        self._dtype = view.dtype
        self._scalar = _lib.F32 if view.dtype == np.float32 else _lib.F64
        self._device = int(device)
        self._h = C.c_void_p()
        n, sdim = view.shape
        L = _lib.lib()
        if _stream is not None:
            consumed = C.c_uint64()
            buf = np.frombuffer(_stream, dtype=np.uint8)
            _lib.check(L.pico_b200_tree_load(_ptr(view), n, sdim, sdim, self._scalar, metric.value, _ptr(buf),
                                             buf.size, self._device, C.byref(self._h), C.byref(consumed)))
        else:
            if max_leaf_depth is not None:
                stop_kind, stop_value = 1, int(max_leaf_depth)
            else:
                stop_kind, stop_value = 0, int(max_leaf_size)
                if stop_value <= 0:
                    raise ValueError("max_leaf_size must be > 0")
            bmin = bmax = None
            if bounds is not None:
                bmin = np.ascontiguousarray(bounds[0], dtype=self._dtype)
                bmax = np.ascontiguousarray(bounds[1], dtype=self._dtype)
                if bmin.shape != (sdim,) or bmax.shape != (sdim,):
                    raise ValueError("bounds must be two points of the tree's spatial dimension")
            _lib.check(L.pico_b200_tree_create(_ptr(view), n, sdim, sdim, self._scalar, metric.value,
                                               Rule(rule).value, stop_kind, stop_value, _ptr(bmin), _ptr(bmax),
                                               self._device, C.byref(self._h)))
        self.last_stats = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().pico_b200_tree_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # ------------------------------------------------------------------ properties
    def __repr__(self):  # _pyco_tree/kd_tree.hpp:321-326
        return (f"KdTree(metric={self._metric.name}, dtype={self._dtype.name}, sdim={self.sdim}, "
                f"npts={self.npts})")

    @property
    def dtype_index(self):
        return np.dtype(np.int32)

    @property
    def dtype_scalar(self):
        return self._dtype

    @property
    def dtype_neighbor(self):
        return _NEIGHBOR[self._dtype]

    @property
    def sdim(self):
        return int(self._view.shape[1])

    @property
    def npts(self):
        return int(self._view.shape[0])

    def metric(self, scalar):
        """Metric applied to a scalar (metric.hpp:93-96,119-122,147-150): |x| or x*x."""
        x = self._dtype.type(scalar)
        return float(x * x) if self._metric in (Metric.L2Squared, Metric.SE2Squared) else float(abs(x))

    def info(self):
        inf = _lib.TreeInfo()
        _lib.check(_lib.lib().pico_b200_tree_info_get(self._h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in inf._fields_ if k != "reserved_"}

    def export(self):
        """(nodes, indices, root_box) copied back from the device (pico_b200_tree_export)."""
        inf = self.info()
        if self._scalar == _lib.F32:
            node_dt = np.dtype([("a", "<u4"), ("b", "<u4"), ("right", "<u4"), ("split_dim", "<u4")])
        else:
            node_dt = np.dtype([("a", "<u8"), ("b", "<u8"), ("right", "<u4"), ("split_dim", "<u4"), ("pad", "<u8")])
        nodes = np.empty(inf["n_nodes"], dtype=node_dt)
        indices = np.empty(self.npts, dtype=np.int32)
        box = np.empty((2, self.sdim), dtype=self._dtype)
        _lib.check(_lib.lib().pico_b200_tree_export(self._h, _ptr(nodes), _ptr(indices), _ptr(box)))
        return nodes, indices, box

    def export_outer_bounds(self):
        """(n_nodes, 2) = (left_min, right_max) per node — topological metrics only."""
        out = np.empty((self.info()["n_nodes"], 2), dtype=self._dtype)
        _lib.check(_lib.lib().pico_b200_tree_export_outer_bounds(self._h, _ptr(out)))
        return out

    def leaf_ranges(self):
        """Index ranges of all non-empty leaves in DFS order (kd_tree.hpp:325)."""
        nodes, indices, _ = self.export()
        leaf = nodes["split_dim"] == 0xFFFFFFFF
        b = nodes["a"][leaf].astype(np.int64)
        e = nodes["b"][leaf].astype(np.int64)
        return [indices[x:y] for x, y in zip(b, e) if y > x]

    # ------------------------------------------------------------------ searches
    def _queries(self, pts):
        return _as_points(pts, self.sdim, self._dtype, "pts")

    def _flags(self, **kw):
        f = 0
        if kw.get("sort"):
            f |= _lib.FLAG_SORT_RESULTS
        if kw.get("reorder") is False:
            f |= _lib.FLAG_NO_REORDER
        if kw.get("warp_per_query"):
            f |= _lib.FLAG_WARP_PER_QUERY
        return f

    # -- device-resident batches (skip the H2D / D2H copies) ---------------------
    @staticmethod
    def _cuda_view(obj, what):
        """(pointer, shape, typestr) of an object exposing __cuda_array_interface__ (torch.Tensor,
        cupy.ndarray, numba device array); it must be C-contiguous."""
        cai = obj.__cuda_array_interface__
        if cai.get("strides") is not None:
            itemsize = int(cai["typestr"][2:])
            expect = []
            acc = itemsize
            for n in reversed(cai["shape"]):
                expect.insert(0, acc)
                acc *= n
            if tuple(cai["strides"]) != tuple(expect):
                raise ValueError(f"{what} not contiguous")
        return int(cai["data"][0]), tuple(cai["shape"]), cai["typestr"]

    def search_knn_device(self, pts, k, e=None, nns=None, stream=None):
        """knn for queries that already live on the tree's GPU. `pts`: (n, sdim) device array of the
        tree's dtype. Returns (or fills) `nns`: a device array of n*k neighbour records seen as int32
        words — shape (n, k, 2) for float32 trees: [..., 0] = index, [..., 1] = bits of the float32
        distance (``nns[..., 1].view(torch.float32)``); (n, k, 4) for float64: word 0 = index, words
        2..3 = the float64 distance. The call is enqueued on `stream` (a CUDA stream handle; default:
        torch's current stream) and does not synchronise."""
        ptr, shape, typestr = self._cuda_view(pts, "pts")
        want = "<f4" if self._dtype == np.float32 else "<f8"
        if typestr != want:
            raise ValueError("array dtype not " + ("float32" if want == "<f4" else "float64"))
        if len(shape) != 2:
            raise ValueError("array ndim not 2")
        if shape[1] != self.sdim:
            raise ValueError("incompatible kd_tree sdim and array inner stride")
        k = int(k)
        if k <= 0:
            raise ValueError("k must be > 0")
        words = 2 if self._dtype == np.float32 else 4
        if nns is None:
            import torch
            nns = torch.empty((shape[0], k, words), dtype=torch.int32, device=torch.device("cuda", self._device))
        optr, oshape, otypestr = self._cuda_view(nns, "nns")
        if otypestr != "<i4" or int(np.prod(oshape)) != shape[0] * k * words:
            raise ValueError("array dtype not neighbor")
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self._device).cuda_stream
        # torch reports its default stream as handle 0, which pico_b200_set_stream reads as "no caller
        # stream" (a private non-blocking stream that is NOT ordered after pending default-stream work):
        # name the legacy default stream explicitly (cudaStreamLegacy).
        stream = int(stream) or _CUDA_STREAM_LEGACY
        L = _lib.lib()
        _lib.check(L.pico_b200_set_stream(C.c_void_p(stream)))
        try:
            _lib.check(L.pico_b200_knn(self._h, C.c_void_p(ptr), shape[0], self.sdim, k, float(e or 0.0),
                                       C.c_void_p(optr), _lib.FLAG_DEVICE_POINTERS | _lib.FLAG_ASYNC, None))
        finally:
            _lib.check(L.pico_b200_set_stream(None))
        return nns

    def profile_leaf_scan(self, pts, repeats=1):
        """Measurement only (pico_b200_profile_leaf_scan): the leaf scan in isolation. `pts` is a device
        array like in search_knn_device. Returns (nns, stats): nns[i] = nearest point inside the leaf the
        descent of query i ends in, stats = {descend_ms, scan_ms, scan_bytes} per launch."""
        import torch
        ptr, shape, typestr = self._cuda_view(pts, "pts")
        if typestr != ("<f4" if self._dtype == np.float32 else "<f8") or len(shape) != 2 or shape[1] != self.sdim:
            raise ValueError("pts must be an (n, sdim) device array of the tree's dtype")
        words = 2 if self._dtype == np.float32 else 4
        nns = torch.empty((shape[0], 1, words), dtype=torch.int32, device=torch.device("cuda", self._device))
        torch.cuda.synchronize(self._device)
        d_ms, s_ms, nbytes = C.c_double(), C.c_double(), C.c_uint64()
        _lib.check(_lib.lib().pico_b200_profile_leaf_scan(self._h, C.c_void_p(ptr), shape[0], self.sdim,
                                                          C.c_void_p(nns.data_ptr()), int(repeats), C.byref(d_ms),
                                                          C.byref(s_ms), C.byref(nbytes)))
        return nns, {"descend_ms": d_ms.value, "scan_ms": s_ms.value, "scan_bytes": int(nbytes.value)}

    def search_knn(self, pts, k, *args, **kw):
        """search_knn(pts, k[, e][, nns]) — def_kd_tree.cpp:59-110. Returns / fills an array
        of shape (npts, k) (or (k, npts) for column-major input, kd_tree.hpp:362-378). Device arrays
        (anything with __cuda_array_interface__) are routed to search_knn_device."""
        if hasattr(pts, "__cuda_array_interface__") and not isinstance(pts, np.ndarray):
            e = kw.pop("e", None)
            nns = kw.pop("nns", None)
            for a in args:
                if isinstance(a, (int, float, np.floating, np.integer)):
                    e = float(a)
                else:
                    nns = a
            return self.search_knn_device(pts, k, e, nns)
        e, nns = self._split_args(args, kw, np.ndarray)
        view, row_major = self._queries(pts)
        k = int(k)
        if k <= 0:
            raise ValueError("k must be > 0")
        nq = view.shape[0]
        shape = (nq, k) if row_major else (k, nq)
        order = "C" if row_major else "F"
        if nns is None:
            nns = np.empty(shape, dtype=self.dtype_neighbor, order=order)
            out = nns if row_major else nns.T
        else:
            # ensure_size of the reference binding (_pyco_tree/kd_tree.hpp:362-378): the array is only resized when
            # its SIZE differs (to (npts,) for k == 1, else (npts, k) / (k, npts)); otherwise its memory is filled
            # as it is, npts rows of k records one after the other, whatever its shape
            if nns.dtype != self.dtype_neighbor:
                raise ValueError("array dtype not neighbor")
            if nns.size != nq * k:
                nns.resize((nq,) if k == 1 else shape, refcheck=False)
            if not (nns.flags["C_CONTIGUOUS"] or nns.flags["F_CONTIGUOUS"]):
                raise ValueError("array not contiguous")
            out = nns.ravel(order="K")  # a view of the dense buffer, in memory order
            if not np.shares_memory(out, nns):
                raise ValueError("array not contiguous")
        stats = _lib.SearchStats()
        _lib.check(_lib.lib().pico_b200_knn(self._h, _ptr(view), nq, self.sdim, k, float(e or 0.0), _ptr(out),
                                            self._flags(**kw), C.byref(stats)))
        self.last_stats = stats
        return nns

    def search_radius(self, pts, radius, *args, **kw):
        """search_radius(pts, radius[, e][, nns][, sort=False]) — def_kd_tree.cpp:112-164."""
        sort = kw.pop("sort", False)
        args = list(args)
        if args and isinstance(args[-1], (bool, np.bool_)):
            sort = bool(args.pop())
        e, nns = self._split_args(args, kw, DArray)
        view, _ = self._queries(pts)
        if nns is not None and nns.dtype != self.dtype_neighbor:
            raise ValueError("array dtype not neighbor")
        nq = view.shape[0]
        offsets = np.zeros(nq + 1, dtype=np.uint64)
        p = C.c_void_p()
        stats = _lib.SearchStats()
        _lib.check(_lib.lib().pico_b200_radius(self._h, _ptr(view), nq, self.sdim, float(radius), float(e or 0.0),
                                               _ptr(offsets), C.byref(p), self._flags(sort=sort, **kw),
                                               C.byref(stats)))
        self.last_stats = stats
        flat = self._take(p, int(offsets[-1]), self.dtype_neighbor)
        if nns is None:
            nns = DArray(self.dtype_neighbor)
        nns._assign(offsets, flat)
        return nns

    def search_box(self, boxes, indices=None, **kw):
        """search_box(boxes[, indices]) — def_kd_tree.cpp:166-180; boxes are (min, max) row
        pairs, so their count must be even (_pyco_tree/kd_tree.hpp:251-253)."""
        view, _ = _as_points(boxes, self.sdim, self._dtype, "boxes")
        if view.shape[0] % 2 != 0:
            raise ValueError("query min and max don't have equal size")
        if indices is not None and indices.dtype != self.dtype_index:
            raise ValueError("array dtype not index")
        nb = view.shape[0] // 2
        mins = view[0::2]
        maxs = view[1::2]
        offsets = np.zeros(nb + 1, dtype=np.uint64)
        p = C.c_void_p()
        stats = _lib.SearchStats()
        # rows of one box are adjacent: min at 2i, max at 2i+1 -> stride of 2*sdim scalars
        _lib.check(_lib.lib().pico_b200_box(self._h, C.c_void_p(mins.ctypes.data), C.c_void_p(maxs.ctypes.data), nb,
                                            2 * self.sdim, _ptr(offsets), C.byref(p), self._flags(**kw),
                                            C.byref(stats)))
        self.last_stats = stats
        flat = self._take(p, int(offsets[-1]), np.dtype(np.int32))
        if indices is None:
            indices = DArray(np.int32)
        indices._assign(offsets, flat)
        return indices

    @staticmethod
    def _split_args(args, kw, out_type):
        """Positional forms of the reference: (out), (e), (e, out)."""
        e = kw.pop("e", None)
        out = kw.pop("nns", None)
        for a in args:
            if isinstance(a, out_type):
                out = a
            elif isinstance(a, (int, float, np.floating, np.integer)):
                e = float(a)
            else:
                raise TypeError(f"unexpected argument of type {type(a).__name__}")
        return e, out

    @staticmethod
    def _take(p, count, dtype):
        """Wrap a buffer handed out by the library as an ndarray WITHOUT copying (results can be
        gigabytes); the buffer is freed when the last view of it dies."""
        if count:
            buf = (C.c_char * (count * dtype.itemsize)).from_address(p.value)
            weakref.finalize(buf, _lib.lib().pico_b200_free, C.c_void_p(p.value))
            return np.frombuffer(buf, dtype=dtype)
        if p.value:
            _lib.lib().pico_b200_free(p)
        return np.empty(0, dtype=dtype)

    # ------------------------------------------------------------------ (de)serialisation
    def _saved_stream(self):
        size = C.c_uint64()
        _lib.check(_lib.lib().pico_b200_tree_save_size(self._h, C.byref(size)))
        buf = np.empty(size.value, dtype=np.uint8)
        _lib.check(_lib.lib().pico_b200_tree_save(self._h, _ptr(buf)))
        return buf.tobytes()

    def serialize(self):
        """Device-layout image of the tree (for broadcasting replicas)."""
        size = C.c_uint64()
        _lib.check(_lib.lib().pico_b200_tree_serialize_size(self._h, C.byref(size)))
        buf = np.empty(size.value, dtype=np.uint8)
        _lib.check(_lib.lib().pico_b200_tree_serialize(self._h, _ptr(buf), 0))
        return buf


def save_kd_tree(tree, filename):
    """.pkd writer — _pyco_tree/kd_tree.hpp:608-614: signature, version, metric string, then
    the kd_tree::save stream."""
    name = tree._metric.name.encode()
    try:
        f = open(filename, "wb")
    except OSError:
        raise RuntimeError("unable to open file: " + str(filename))
    with f:
        f.write(_PKD_SIGNATURE + struct.pack("<I", _PKD_VERSION) + struct.pack("<Q", len(name)) + name)
        f.write(tree._saved_stream())


def load_kd_tree(pts, filename, *, device=0):
    """.pkd reader — _pyco_tree/kd_tree.hpp:602-606,552-585."""
    try:
        f = open(filename, "rb")
    except OSError:
        raise RuntimeError("unable to open file: " + str(filename))
    with f:
        blob = f.read()
    if blob[:4] != _PKD_SIGNATURE:
        raise RuntimeError("unexpected header signature")
    (version,) = struct.unpack_from("<I", blob, 4)
    if version != _PKD_VERSION:
        raise RuntimeError("unsupported header version")
    (ln,) = struct.unpack_from("<Q", blob, 8)
    name = blob[16:16 + ln].decode()
    try:
        metric = Metric[name]
    except KeyError:
        raise RuntimeError("unexpected metric string")
    return KdTree(pts, metric, _stream=blob[16 + ln:], device=device)

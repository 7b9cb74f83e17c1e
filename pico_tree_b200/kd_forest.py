"""Host-side mirror of ``pico_tree::kd_forest`` (examples/pico_understory/pico_understory/kd_forest.hpp:15-138).

The reference has no Python binding for the forest; this follows the C++ class: constructor
``kd_forest(space, max_leaf_size, forest_size)`` (:44-52), ``search_nn(x, max_leaves_visited, nn)`` (:83-88) and
``search_nearest(x, max_leaves_visited, visitor)`` with a ``search_knn`` visitor (:76-81), as batches. Everything
runs in libpico_b200.so (csrc/forest.cu); nothing here computes distances or walks a tree.
"""
import ctypes as C

import numpy as np

from . import _lib
from .kd_tree import _NEIGHBOR, _as_points, _ptr

__all__ = ["KdForest"]


class KdForest:
    """``KdForest(pts, max_leaf_size, forest_size)``. ``rotations`` (forest_size x sdim unit vectors) makes the
    forest reproducible; by default they are drawn at random like the reference's (rkd_tree_hh_data.hpp:14-29)."""

    def __init__(self, pts, max_leaf_size=10, forest_size=4, *, rotations=None, device=0):
        view, row_major = _as_points(pts, what="pts")
        if not row_major:
            view = np.ascontiguousarray(view)
        self._pts = pts
        self._view = view
        self._dtype = view.dtype
        self._scalar = _lib.F32 if view.dtype == np.float32 else _lib.F64
        self._device = int(device)
        self._h = C.c_void_p()
        n, sdim = view.shape
        if int(max_leaf_size) <= 0:
            raise ValueError("max_leaf_size must be > 0")
        if int(forest_size) <= 0:
            raise ValueError("forest_size must be > 0")
        rot = None
        if rotations is not None:
            rot = np.ascontiguousarray(rotations, dtype=self._dtype)
            if rot.shape != (int(forest_size), sdim):
                raise ValueError("rotations must be forest_size vectors of the space's dimension")
        _lib.check(_lib.lib().pico_b200_forest_create(_ptr(view), n, sdim, sdim, self._scalar, int(max_leaf_size),
                                                      _ptr(rot), int(forest_size), self._device, C.byref(self._h)))
        self.last_stats = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().pico_b200_forest_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    @property
    def sdim(self):
        return int(self._view.shape[1])

    @property
    def npts(self):
        return int(self._view.shape[0])

    @property
    def dtype_neighbor(self):
        return _NEIGHBOR[self._dtype]

    def info(self):
        inf = _lib.ForestInfo()
        _lib.check(_lib.lib().pico_b200_forest_info_get(self._h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in inf._fields_}

    @property
    def rotations(self):
        out = np.empty((self.info()["n_trees"], self.sdim), dtype=self._dtype)
        _lib.check(_lib.lib().pico_b200_forest_rotations(self._h, _ptr(out)))
        return out

    def export_tree(self, i):
        """(nodes, indices, root_box, outer_bounds) of tree `i`, copied back from the device."""
        L = _lib.lib()
        t = C.c_void_p()
        _lib.check(L.pico_b200_forest_tree(self._h, int(i), C.byref(t)))
        inf = _lib.TreeInfo()
        _lib.check(L.pico_b200_tree_info_get(t, C.byref(inf)))
        if self._scalar == _lib.F32:
            node_dt = np.dtype([("a", "<u4"), ("b", "<u4"), ("right", "<u4"), ("split_dim", "<u4")])
        else:
            node_dt = np.dtype([("a", "<u8"), ("b", "<u8"), ("right", "<u4"), ("split_dim", "<u4"), ("pad", "<u8")])
        nodes = np.empty(inf.n_nodes, dtype=node_dt)
        indices = np.empty(self.npts, dtype=np.int32)
        box = np.empty((2, self.sdim), dtype=self._dtype)
        outer = np.empty((inf.n_nodes, 2), dtype=self._dtype)
        _lib.check(L.pico_b200_tree_export(t, _ptr(nodes), _ptr(indices), _ptr(box)))
        _lib.check(L.pico_b200_tree_export_outer_bounds(t, _ptr(outer)))
        return nodes, indices, box, outer

    def search_knn(self, pts, k, max_leaves_visited, nns=None):
        """The k approximate nearest neighbours of every row of `pts`, at most `max_leaves_visited` leaves per
        tree: (npts, k) records {index, distance}, ascending. kd_forest::search_nearest with a search_knn visitor."""
        q, row_major = _as_points(pts, self.sdim, self._dtype, "pts")
        if not row_major:
            q = np.ascontiguousarray(q)
        k = int(k)
        if k <= 0:
            raise ValueError("k must be > 0")
        if int(max_leaves_visited) < 0:
            raise ValueError("max_leaves_visited must be >= 0")
        k = min(k, self.npts)
        n = q.shape[0]
        if nns is None or nns.size != n * k or nns.dtype != self.dtype_neighbor or not nns.flags.c_contiguous:
            nns = np.empty((n, k), dtype=self.dtype_neighbor)
        stats = _lib.SearchStats()
        _lib.check(_lib.lib().pico_b200_forest_knn(self._h, _ptr(q), n, self.sdim, k, int(max_leaves_visited),
                                                   _ptr(nns), 0, C.byref(stats)))
        self.last_stats = stats
        return nns.reshape(n, k)

    def search_nn(self, pts, max_leaves_visited, nns=None):
        """kd_forest::search_nn (kd_forest.hpp:83-88) for every row of `pts`."""
        return self.search_knn(pts, 1, max_leaves_visited, nns)

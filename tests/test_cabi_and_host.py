"""CPU-side checks: the C-ABI library loads and exports every symbol include/pico_b200.h
declares; without a GPU the compute entry points fail loudly (no CPU fallback); the Python
host mirror validates its inputs like the reference binding does."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pico_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pico_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pico_tree_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pico_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert L.pico_b200_abi_version() == 1


def _has_gpu():
    from pico_tree_b200 import _lib
    c = C.c_int(0)
    _lib.lib().pico_b200_device_count(C.byref(c))
    return c.value > 0


def test_no_cpu_fallback_without_device():
    if _has_gpu():
        pytest.skip("a GPU is visible")
    import pico_tree_b200 as pt
    from pico_tree_b200._lib import PicoB200Error
    a = np.array([[2, 1], [4, 3], [8, 7]], np.float32)
    with pytest.raises(PicoB200Error) as ei:
        pt.KdTree(a, pt.Metric.L2Squared, 10)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_input_validation_matches_reference_binding():
    # test/pyco_tree/kd_tree_test.py:20-23,39-42: ndim != 2 and non-contiguous arrays raise ValueError
    # before any device work (py_array_map.hpp:40-55)
    import pico_tree_b200 as pt
    with pytest.raises(ValueError, match="ndim"):
        pt.KdTree(np.zeros((3, 1, 2), np.float32), pt.Metric.L2Squared, 10)
    big = np.zeros((6, 4), np.float32)
    with pytest.raises(ValueError, match="contiguous"):
        pt.KdTree(big[::2, ::2], pt.Metric.L2Squared, 10)
    with pytest.raises(ValueError, match="dtype"):
        pt.KdTree(np.zeros((3, 2), np.int32), pt.Metric.L2Squared, 10)
    with pytest.raises(ValueError):
        pt.KdTree(np.zeros((3, 2), np.float32), pt.Metric.L2Squared, 0)


def test_darray_semantics():
    # kd_tree_test.py:151-219: len, indexing, negative index, slices, truthiness, dtype
    import pico_tree_b200 as pt
    d = pt.DArray(np.int32)
    assert not d and len(d) == 0 and d.dtype == np.int32
    d._assign(np.array([0, 1, 1, 4, 5], np.uint64), np.arange(5, dtype=np.int32))
    assert d and len(d) == 4
    assert [len(x) for x in d] == [1, 0, 3, 1]
    s = d[0:4:2]
    assert len(s) == 2 and [len(x) for x in s] == [1, 3] and len(s[-1]) == 3
    with pytest.raises(IndexError):
        d[4]
    # buffer re-use when the new result fits (kd_tree_test.py:107-118)
    addr = d[2].ctypes.data
    d._assign(np.array([0, 1, 1, 4, 5], np.uint64), np.arange(5, dtype=np.int32) + 10)
    assert d[2].ctypes.data == addr and d[2].tolist() == [11, 12, 13]


def test_pkd_header_errors(tmp_path):
    # _pyco_tree/kd_tree.hpp:552-585: bad signature / version / metric string -> RuntimeError
    import pico_tree_b200 as pt
    a = np.zeros((3, 2), np.float32)
    with pytest.raises(RuntimeError, match="unable to open file"):
        pt.load_kd_tree(a, str(tmp_path / "missing.pkd"))
    p = tmp_path / "bad.pkd"
    p.write_bytes(b"NOPE" + b"\0" * 32)
    with pytest.raises(RuntimeError, match="signature"):
        pt.load_kd_tree(a, str(p))
    p.write_bytes(b"\x89PKD" + (7).to_bytes(4, "little") + b"\0" * 32)
    with pytest.raises(RuntimeError, match="version"):
        pt.load_kd_tree(a, str(p))
    p.write_bytes(b"\x89PKD" + (1).to_bytes(4, "little") + (3).to_bytes(8, "little") + b"XYZ")
    with pytest.raises(RuntimeError, match="metric"):
        pt.load_kd_tree(a, str(p))


def test_datasets_are_deterministic():
    from pico_tree_b200 import datasets as D
    a = D.lidar_shape(10000, seed=3)
    b = D.lidar_shape(10000, seed=3)
    assert a.dtype == np.float32 and a.shape == (10000, 3) and np.array_equal(a, b)
    assert not np.array_equal(a, D.lidar_shape(10000, seed=4))
    s = D.sift_shape(100)
    assert s.shape == (100, 128) and s.min() >= 0 and s.max() <= 255 and np.all(s == np.floor(s))


def test_hostmem_helpers_without_gpu():
    """pico_tree_b200/hostmem.py: cpulist parsing, and binding is a described no-op where the GPU's NUMA node
    cannot be told (no GPU here)."""
    import os

    from pico_tree_b200 import hostmem
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = hostmem.bind_to_gpu_node(0)
    assert info["bound"] is False and info["numa_node"] is None
    assert os.sched_getaffinity(0) == before

"""The C++ face of the drop-in boundary (include/pico_tree_b200/kd_tree.hpp).

CPU (`-m "not gpu"`): the header and the C++ test-suite compile and link against libpico_b200.so;
where /root/reference exists, the reference's OWN example programs and its UNMODIFIED pybind11 binding
compile against a shadow include tree in which only pico_tree/kd_tree.hpp is replaced by ours.

GPU (`-m gpu`): tests/_bin/kd_tree_test (the reference's KdTreeTest suite restated, tests/cpp/) runs on
the device; the reference's binding built against our engine (tests/_bin/_pyco_tree.so) answers the
cases of the reference's test/pyco_tree/kd_tree_test.py.
"""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_bin")
REF = "/root/reference"


def _make(*targets):
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s", *targets], check=True)


def test_cpp_suite_builds():
    _make(os.path.join(BIN, "kd_tree_test"))
    assert os.access(os.path.join(BIN, "kd_tree_test"), os.X_OK)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources only exist in the dev container")
def test_reference_examples_compile_against_dropin(tmp_path):
    """examples/kd_tree/*.cpp of the reference, unmodified, against our kd_tree.hpp + the reference's own
    trait headers. kd_tree_custom_metric.cpp must stop at the static_assert that names the device metrics."""
    shadow = tmp_path / "pico_tree"
    shadow.mkdir()
    for f in glob.glob(os.path.join(REF, "src/pico_tree/pico_tree/*")):
        if os.path.basename(f) != "kd_tree.hpp":
            os.symlink(f, shadow / os.path.basename(f))
    (shadow / "kd_tree.hpp").write_text("#pragma once\n#define PICO_TREE_B200_USE_REFERENCE_TRAITS 1\n"
                                        "#include <pico_tree_b200/kd_tree.hpp>\n")
    base = ["g++", "-std=c++17", "-fsyntax-only", f"-I{tmp_path}", f"-I{ROOT}/include",
            f"-I{REF}/examples/pico_toolshed"]
    sources = sorted(glob.glob(os.path.join(REF, "examples/kd_tree/*.cpp")))
    assert len(sources) >= 9
    for src in sources:
        r = subprocess.run(base + [src], capture_output=True, text=True)
        if src.endswith("kd_tree_custom_metric.cpp"):
            assert r.returncode != 0 and "METRIC_HAS_NO_DEVICE_IMPLEMENTATION" in r.stderr
        else:
            assert r.returncode == 0, f"{src}:\n{r.stderr[-2000:]}"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources only exist in the dev container")
def test_reference_binding_builds_against_dropin():
    _make("ref")
    assert os.path.exists(os.path.join(BIN, "_pyco_tree.so"))


# ---------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cpp_suite_on_device():
    exe = os.path.join(BIN, "kd_tree_test")
    assert os.path.exists(exe), "tests/_bin/kd_tree_test missing: run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " 0 failed" in r.stdout


@pytest.fixture(scope="module")
def ref_binding():
    path = os.path.join(BIN, "_pyco_tree.so")
    if not os.path.exists(path):
        pytest.skip("tests/_bin/_pyco_tree.so was not built (needs /root/reference at build time)")
    sys.path.insert(0, BIN)
    import _pyco_tree
    return _pyco_tree


@pytest.mark.gpu
def test_reference_binding_three_point_cases(ref_binding):
    """test/pyco_tree/kd_tree_test.py:53-69,90-118,151-192 through the reference's own binding code."""
    m = ref_binding
    a = np.array([[2, 1], [4, 3], [8, 7]], dtype=np.float32)
    t = m.KdTree(a, m.Metric.L2Squared, 10)
    assert (t.sdim, t.npts) == (2, 3) and t.metric(-2.0) == 4
    nns = t.search_knn(a, 2)
    assert nns.shape == (3, 2)
    assert [int(x[0][0]) for x in nns] == [0, 1, 2] and all(float(x[0][1]) == 0 for x in nns)
    nns = t.search_knn(a, 2, 1.0)
    assert [int(x[0][0]) for x in nns] == [0, 1, 2]
    rad = t.search_radius(a, t.metric(2.5))
    assert len(rad) == 3 and [len(n) for n in rad] == [1, 1, 1] and [int(n[0][0]) for n in rad] == [0, 1, 2]
    boxes = np.array([[0, 0], [3, 3], [2, 2], [3, 3], [0, 0], [9, 9], [6, 6], [9, 9]], dtype=np.float32)
    res = t.search_box(boxes)
    assert [len(n) for n in res] == [1, 0, 3, 1]
    t1 = m.KdTree(a, m.Metric.L1, 10)
    assert t1.metric(-2.0) == 2


@pytest.mark.gpu
def test_reference_binding_agrees_with_host_mirror(ref_binding, tmp_path):
    """Same engine below both hosts: the reference's binding (OpenMP threads issuing single-query calls
    concurrently) and pico_tree_b200.KdTree (one batch call) must return identical arrays."""
    import pico_tree_b200 as pt
    m = ref_binding
    rng = np.random.default_rng(5)
    pts = rng.random((20000, 3), dtype=np.float32)
    q = rng.random((2000, 3), dtype=np.float32)
    ours = pt.KdTree(pts, pt.Metric.L2Squared, 10)
    theirs = m.KdTree(pts, m.Metric.L2Squared, 10)
    a, b = ours.search_knn(q, 4), theirs.search_knn(q, 4)
    assert np.array_equal(a["index"], b["index"]) and np.array_equal(a["distance"], b["distance"])
    ra, rb = ours.search_radius(q, 0.002), theirs.search_radius(q, 0.002)
    assert [len(x) for x in ra] == [len(x) for x in rb]
    f = str(tmp_path / "t.pkd")
    m.save_kd_tree(theirs, f)          # written by the reference's binding code ...
    back = pt.load_kd_tree(pts, f)     # ... read by the host mirror
    assert np.array_equal(back.search_knn(q, 4)["index"], a["index"])

"""pico_tree_b200 — B200-native KdTree engine behind PicoTree's Python API.

Mirrors ``pico_tree`` (src/pyco_tree/pico_tree/__init__.py:1-5 of the reference):
``KdTree``, ``Metric``, ``DArray``, ``load_kd_tree``, ``save_kd_tree``. All
search/build work runs in libpico_b200.so (hand-written CUDA, sm_100a).
"""
from .kd_forest import KdForest  # noqa: F401
from .kd_tree import DArray, KdTree, Metric, load_kd_tree, save_kd_tree  # noqa: F401

__all__ = ["KdTree", "KdForest", "Metric", "DArray", "load_kd_tree", "save_kd_tree"]

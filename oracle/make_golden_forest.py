"""Generates tests/golden/forest/*.npz from the UNMODIFIED reference kd_forest
(oracle/_ref/libpico_ref_forest.so, examples/pico_understory/pico_understory/kd_forest.hpp).

    python oracle/make_golden_forest.py

kd_forest draws its Householder vectors from std::random_device, so every fixture stores the vectors its
forest drew next to the points, the queries and the results; tests/test_oracle_golden.py replays them through
the C restatement (po_forest_*). SURVEY.md §8 f4: oracle and fixtures first, the device path follows.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from oracle.make_golden import cloud  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "forest")

CASES = [
    # name, kind, n, sdim, dtype, max_leaf_size, forest_size
    ("forest_uniform3_f32", "uniform", 4000, 3, np.float32, 10, 4),
    ("forest_clustered3_f32", "clustered", 4000, 3, np.float32, 6, 3),
    ("forest_uniform16_f32", "uniform", 3000, 16, np.float32, 8, 6),
    ("forest_uniform8_f64", "uniform", 3000, 8, np.float64, 10, 4),
]
SEARCHES = [(1, 1), (1, 4), (1, 32), (1, 1 << 30), (4, 8), (8, 64)]  # (k, max_leaves_visited)


def main():
    assert O.ref_forest_available(), "build oracle/_ref first (make -C oracle)"
    os.makedirs(OUT, exist_ok=True)
    for name, kind, n, sdim, dtype, leaf, trees in CASES:
        pts = cloud(kind, n, sdim, seed=len(name) * 7 + n, dtype=dtype)
        q = np.ascontiguousarray(np.random.default_rng(n + sdim).random((600, sdim)).astype(dtype))
        ref = O.RefForest(pts, leaf, trees)
        out = {"pts": pts, "q": q, "rotations": ref.rotations, "max_leaf_size": leaf,
               "searches": np.array(SEARCHES, dtype=np.int64)}
        for k, ml in SEARCHES:
            r = ref.search_knn(q, k, ml)
            out[f"index_k{k}_m{ml}"] = r["index"]
            out[f"distance_k{k}_m{ml}"] = r["distance"]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "rotations", out["rotations"].shape)


if __name__ == "__main__":
    main()

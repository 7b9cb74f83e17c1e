"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libpico_ref.so).

Run in the dev container (needs /root/reference to build _ref):
    python oracle/make_golden.py
The fixtures pin the C oracle (tests/test_oracle_golden.py) and, through it, the CUDA
path on the GPU box where the reference sources do not exist.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cloud(kind, n, sdim, seed, dtype):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        p = rng.random((n, sdim))
    elif kind == "clustered":  # planes + duplicates: exercises both slide directions
        p = rng.random((n, sdim))
        p[: n // 3, sdim - 1] = 0.25 + rng.normal(0, 1e-3, n // 3)
        p[n // 3: n // 2, 0] = 0.5
        p[::11] = p[5]
    elif kind == "grid":  # many exact distance ties
        side = int(round(n ** (1.0 / sdim))) + 1
        g = np.stack(np.meshgrid(*[np.arange(side)] * sdim, indexing="ij"), -1).reshape(-1, sdim)[:n]
        p = g.astype(np.float64) / side
        rng.shuffle(p, axis=0)
    return np.ascontiguousarray(p.astype(dtype))


CASES = [
    # name, kind, n, sdim, dtype, metric, rule, stop, stop_value
    ("uniform3_f32", "uniform", 4000, 3, np.float32, "l2_squared", "sliding_midpoint", "max_leaf_size", 10),
    ("uniform2_f32_l1", "uniform", 3000, 2, np.float32, "l1", "sliding_midpoint", "max_leaf_size", 4),
    ("uniform3_f32_lpinf", "uniform", 3000, 3, np.float32, "lpinf", "sliding_midpoint", "max_leaf_size", 1),
    ("clustered3_f32", "clustered", 5000, 3, np.float32, "l2_squared", "sliding_midpoint", "max_leaf_size", 10),
    ("clustered2_f32_depth", "clustered", 3000, 2, np.float32, "l2_squared", "sliding_midpoint", "max_leaf_depth", 7),
    ("uniform3_f32_midpoint", "uniform", 3000, 3, np.float32, "l2_squared", "midpoint", "max_leaf_size", 6),
    ("uniform3_f32_median", "uniform", 3000, 3, np.float32, "l2_squared", "median", "max_leaf_size", 8),
    ("grid3_f32", "grid", 4096, 3, np.float32, "l2_squared", "sliding_midpoint", "max_leaf_size", 10),
    ("uniform8_f32", "uniform", 3000, 8, np.float32, "l2_squared", "sliding_midpoint", "max_leaf_size", 10),
    ("uniform3_f64", "uniform", 3000, 3, np.float64, "l2_squared", "sliding_midpoint", "max_leaf_size", 10),
    ("clustered5_f64_l1", "clustered", 3000, 5, np.float64, "l1", "sliding_midpoint", "max_leaf_size", 5),
    # topological spaces (search_nearest_topological, metric_box_map): S1 = [0, 1) and R2 x S1
    ("so2_f32", "uniform", 3000, 1, np.float32, "so2", "sliding_midpoint", "max_leaf_size", 10),
    ("se2_f32", "uniform", 4000, 3, np.float32, "se2_squared", "sliding_midpoint", "max_leaf_size", 10),
    ("se2_f64", "uniform", 3000, 3, np.float64, "se2_squared", "sliding_midpoint", "max_leaf_size", 6),
]


def main(only=None):
    """`python oracle/make_golden.py [name ...]` regenerates the named fixtures (all by default)."""
    assert O.ref_available(), "build oracle/_ref first (make -C oracle)"
    os.makedirs(OUT, exist_ok=True)
    for name, kind, n, sdim, dtype, metric, rule, stop, sv in CASES:
        if only and name not in only:
            continue
        pts = cloud(kind, n, sdim, 100 + len(name), dtype)
        rng = np.random.default_rng(7)
        topo = metric in O.TOPOLOGICAL
        extra = {}
        if topo:
            # coordinates on the circle stay inside [0, 1); boxes on S1 dimensions partly wrap (min > max)
            q = np.ascontiguousarray(np.concatenate([rng.random((300, sdim)), pts[:100]]).astype(dtype))
            r = O.RefTree(pts, sv, metric=metric, rule=rule, stop=stop)
            _, idx, box, nodes, outer = r.structure()
            extra["node_outer"] = outer
            rad = 0.03 if metric == "so2" else 0.004
            mins = (q - dtype(0.06)).astype(dtype)
            maxs = (q + dtype(0.045)).astype(dtype)
            s1 = slice(0, None) if metric == "so2" else slice(2, None)
            mins[:, s1] = np.mod(mins[:, s1], 1)
            maxs[:, s1] = np.mod(maxs[:, s1], 1)
        else:
            q = np.ascontiguousarray(np.concatenate([rng.random((300, sdim)) * 1.2 - 0.1, pts[:100]]).astype(dtype))
            r = O.RefTree(pts, sv, metric=metric, rule=rule, stop=stop)
            _, idx, box, nodes = r.structure()
            rad = 0.02 if metric == "l2_squared" else 0.12
            mins = np.minimum(q, q[::-1]) - dtype(0.02)
            maxs = np.maximum(q, q[::-1]) * dtype(0.5) + mins * dtype(0.5) + dtype(0.05)
        ro, rn = r.search_radius(q, rad)
        so, sn = r.search_radius(q, rad, e=1.5, sort=True)
        bo, bi = r.search_box(mins, maxs)
        k = min(7, n)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            pts=pts, q=q, metric=metric, rule=rule, stop=stop, stop_value=sv, radius=rad,
            indices=idx, root_box=box,
            node_left_max=nodes["left_max"], node_right_min=nodes["right_min"], node_split_dim=nodes["split_dim"],
            node_begin=nodes["begin"], node_end=nodes["end"], node_left=nodes["left"], node_right=nodes["right"],
            nn_index=r.search_knn(q, 1)["index"], nn_distance=r.search_knn(q, 1)["distance"],
            knn_index=r.search_knn(q, k)["index"], knn_distance=r.search_knn(q, k)["distance"],
            aknn_index=r.search_knn(q, k, e=1.5)["index"], aknn_distance=r.search_knn(q, k, e=1.5)["distance"],
            radius_offsets=ro, radius_index=rn["index"], radius_distance=rn["distance"],
            aradius_offsets=so, aradius_distance=sn["distance"],
            box_min=mins, box_max=maxs, box_offsets=bo, box_index=bi,
            saved_stream=np.frombuffer(r.saved(), dtype=np.uint8), **extra)
        print(name, len(nodes), "nodes", int(ro[-1]), "radius hits", int(bo[-1]), "box hits")


if __name__ == "__main__":
    main(set(sys.argv[1:]))

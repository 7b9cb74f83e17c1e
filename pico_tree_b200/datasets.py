"""Deterministic synthetic point clouds for tests and bench.py.

The reference's benchmark data (Bremen "Gaussian Point" LiDAR scans, docs/benchmark.md:11-13)
is not available offline; ``lidar_shape`` is the stand-in named by BASELINE.json: a spinning
multi-ring scanner at several poses inside a room, points kept in scan order like the
reference's ``scans*.bin`` files (examples/pico_toolshed/pico_toolshed/format/format_bin.hpp:9-32
raw float[3] records).
"""
import numpy as np

N_TREE = 7_733_372     # README.md:14-16 of the reference
N_QUERY = 7_200_863


def uniform(n, sdim=3, seed=1, dtype=np.float32):
    """cfg1: i.i.d. uniform [0, 1) coordinates."""
    rng = np.random.default_rng(seed)
    return rng.random((n, sdim), dtype=np.float32).astype(dtype, copy=False)


def lidar_shape(n, seed=1, pose_shift=0.0, rings=64, poses=9, room=(50.0, 50.0, 8.0), max_range=60.0,
                sigma=0.01, shuffle=False, dtype=np.float32):
    """cfg2: `poses` scanner positions inside a `room` (metres); per pose a `rings`-ring scanner
    sweeps the azimuth monotonically (rings interleaved, i.e. scan order is preserved in
    memory); rays hit floor / ceiling / walls, are capped at `max_range` and get `sigma` of
    Gaussian noise per coordinate."""
    rng = np.random.default_rng(seed)
    per_pose = [n // poses + (1 if i < n % poses else 0) for i in range(poses)]
    out = np.empty((n, 3), dtype=np.float64)
    lo = np.zeros(3)
    hi = np.asarray(room, dtype=np.float64)
    row = 0
    for p, cnt in enumerate(per_pose):
        # poses on a 3x3 grid with jitter, sensor ~1.8 m above the floor
        gx, gy = (p % 3 + 0.5) / 3.0, (p // 3 % 3 + 0.5) / 3.0
        origin = np.array([gx * room[0], gy * room[1], 1.8]) + rng.normal(0.0, 0.5, 3) * np.array([1, 1, 0.05])
        origin[:2] += pose_shift
        origin = np.clip(origin, lo + 0.5, hi - 0.5)
        i = np.arange(cnt)
        ring = i % rings
        steps = max((cnt + rings - 1) // rings, 1)
        az = 2.0 * np.pi * (i // rings) / steps
        el = np.deg2rad(-24.8 + (ring + 0.5) * (26.8 / rings))  # HDL-64-like vertical field of view
        d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            t_lo = (lo - origin) / d
            t_hi = (hi - origin) / d
        t = np.where(d > 0, t_hi, t_lo)
        t = np.where(d == 0, np.inf, t)
        t = np.minimum(t.min(axis=1), max_range)
        pts = origin + d * t[:, None]
        pts += rng.normal(0.0, sigma, pts.shape)
        out[row:row + cnt] = pts
        row += cnt
    if shuffle:
        rng.shuffle(out, axis=0)
    return np.ascontiguousarray(out.astype(dtype))


def sift_shape(n, sdim=128, seed=1, dtype=np.float32):
    """cfg4: non-negative, integer valued, skewed descriptors (floor(gamma(0.6, 45)) clipped to 255)."""
    rng = np.random.default_rng(seed)
    x = np.floor(rng.gamma(0.6, 45.0, size=(n, sdim)))
    return np.clip(x, 0, 255).astype(dtype)


def bench_clouds(n_tree=N_TREE, n_query=N_QUERY, seed=1):
    """cfg2 tree / query clouds: different seeds and slightly shifted poses."""
    return lidar_shape(n_tree, seed=seed), lidar_shape(n_query, seed=seed + 1, pose_shift=0.35)

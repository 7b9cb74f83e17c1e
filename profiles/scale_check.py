"""Beyond the BASELINE sizes: 50M uniform points, 20M queries, knn=1 and knn=8 — build, search, and index / distance
parity of a 2M-query sample against the reference (the whole tree structure is compared too)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pico_tree_b200 as pt  # noqa: E402
from oracle import oracle as O  # noqa: E402

n, nq = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000, 20_000_000
rng = np.random.default_rng(1)
pts = rng.random((n, 3), dtype=np.float32)
q = rng.random((nq, 3), dtype=np.float32)
t0 = time.perf_counter()
tree = pt.KdTree(pts, pt.Metric.L2Squared, 10)
info = tree.info()
print(f"build: {time.perf_counter() - t0:.2f} s wall, {info['build_ms']:.1f} ms device, {info['n_nodes']} nodes, "
      f"height {info['height']}, {info['device_bytes'] / 1e9:.2f} GB on device", flush=True)
t0 = time.perf_counter()
ref = O.RefTree(pts, 10)
print(f"reference build: {time.perf_counter() - t0:.2f} s", flush=True)
_, idx, box, nodes = ref.structure()
mine_nodes, mine_idx, mine_box = tree.export()
print("index permutation equal:", bool(np.array_equal(mine_idx, idx)), " root box equal:", bool(np.array_equal(mine_box, box)),
      " node count equal:", len(mine_nodes) == len(nodes), flush=True)
for k in (1, 8):
    t0 = time.perf_counter()
    got = tree.search_knn(q, k)
    dt = time.perf_counter() - t0
    ns = 2_000_000
    want = ref.search_knn(q[:ns], k, threads=O.max_threads())
    print(f"knn={k}: {dt * 1e3:.1f} ms end to end (pageable numpy) = {nq / dt / 1e6:.0f} Mq/s; sample of {ns}: distances bit-equal "
          f"{bool(np.array_equal(got['distance'][:ns], want['distance']))}, index mismatches "
          f"{int(np.count_nonzero(got['index'][:ns] != want['index']))}", flush=True)

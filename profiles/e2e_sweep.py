"""Sweep of the host-buffer pipeline of pico_b200_knn (chunk size x stream count) on cfg2, knn=1:
end-to-end Mq/s with pinned host buffers. Each point runs in its own process because the hooks
(PICO_B200_HOST_CHUNK / PICO_B200_HOST_STREAMS) are read once."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ONE = r"""
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
import pico_tree_b200 as pt
from pico_tree_b200 import datasets as D
tree_pts, q = D.bench_clouds()
tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
qp = torch.from_numpy(q).pin_memory().numpy()
out = torch.empty((len(q), 1, 2), dtype=torch.int32).pin_memory().numpy().view(tree.dtype_neighbor).reshape(len(q), 1)
for _ in range(3): tree.search_knn(qp, 1, out)
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10): tree.search_knn(qp, 1, out)
    best = min(best, (time.perf_counter() - t0) / 10)
print("%%.3f ms  %%.1f Mq/s" %% (best * 1e3, len(q) / best / 1e6))
""" % ROOT

for chunk in (524288, 1048576, 1572864, 2097152, 3670016):
    for streams in (2, 3, 4, 6):
        env = dict(os.environ, PICO_B200_HOST_CHUNK=str(chunk), PICO_B200_HOST_STREAMS=str(streams))
        r = subprocess.run([sys.executable, "-c", ONE], capture_output=True, text=True, env=env)
        print("chunk %8d streams %d : %s" % (chunk, streams, (r.stdout.strip() or r.stderr[-300:])), flush=True)

"""128-D path (sift-shape 1M x 128, knn=10): kernel time against the number of queries in the batch —
a single query is latency-bound; throughput saturates with the number of resident warps.
PICO_B200_TILE_ROWS (rows of the staged leaf tile) is swept by the caller."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pico_tree_b200 as pt  # noqa: E402
from pico_tree_b200 import datasets as D  # noqa: E402

counts = [int(x) for x in sys.argv[1:]] or [32, 1184, 4736, 9472]
pts = D.sift_shape(1_000_000, seed=1)
tree = pt.KdTree(pts, pt.Metric.L2Squared, 10)
print(tree.info())
for nq in counts:
    q = D.sift_shape(nq, seed=2)
    tree.search_knn(q, 10, reorder=False)
    ms = tree.last_stats.kernel_ms
    print("tile_rows", os.environ.get("PICO_B200_TILE_ROWS", "default"), nq, "queries:", round(ms, 1), "ms  ->",
          round(nq / ms * 1e3), "q/s", flush=True)

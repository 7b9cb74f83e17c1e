"""Host-buffer pipeline of pico_b200_knn on cfg2, knn=1, against the PCIe floor of the box: copy-ahead depth
(PICO_B200_HOST_AHEAD), stream count, chunk size, head / tail chunk shaping and Morton bits; resident ms/step
beside the end-to-end time of every point, and per-chunk timelines (PICO_B200_TIMELINE). Each point runs in its
own process because the hooks are read once; the clouds are generated once and shared through /tmp."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/tmp/pico_b200_bench_clouds.npz"

ONE = r"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, %r)
import pico_tree_b200 as pt
z = np.load(%r)
tree_pts, q = z["tree"], z["q"]
tree = pt.KdTree(tree_pts, pt.Metric.L2Squared, 10)
qp = torch.from_numpy(q).pin_memory().numpy()
out = torch.empty((len(q), 1, 2), dtype=torch.int32).pin_memory().numpy().view(tree.dtype_neighbor).reshape(len(q), 1)
for _ in range(3): tree.search_knn(qp, 1, out)
if os.environ.get("PICO_B200_TIMELINE"): sys.exit(0)
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10): tree.search_knn(qp, 1, out)
    best = min(best, (time.perf_counter() - t0) / 10)
qd = torch.from_numpy(q).cuda()
od = torch.empty((len(q), 1, 2), dtype=torch.int32, device="cuda")
for _ in range(3): tree.search_knn_device(qd, 1, nns=od)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): tree.search_knn_device(qd, 1, nns=od)
e1.record(); torch.cuda.synchronize()
assert np.array_equal(od.cpu().numpy()[:, 0, 0], out["index"][:, 0])
print("e2e %%.3f ms %%.0f Mq/s | resident %%.3f ms" %% (best * 1e3, len(q) / best / 1e6, e0.elapsed_time(e1) / 50))
""" % (ROOT, CACHE)

FLOOR = r"""
import sys, time, torch
sys.path.insert(0, %r)
import numpy as np
from bench import pcie_floor_ms
z = np.load(%r)
q = torch.from_numpy(z["q"]).pin_memory()
out = torch.empty((len(z["q"]), 1, 2), dtype=torch.int32).pin_memory()
dev = torch.device("cuda", 0)
d = torch.empty(q.shape, dtype=q.dtype, device=dev)
def t(fn, n=10):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3
do = torch.empty(out.shape, dtype=out.dtype, device=dev)
h2d = t(lambda: d.copy_(q, non_blocking=True))
d2h = t(lambda: out.copy_(do, non_blocking=True))
print("H2D %%d MB: %%.3f ms (%%.1f GB/s)   D2H %%d MB: %%.3f ms (%%.1f GB/s)   both at once: %%.3f ms" %% (
    q.nbytes >> 20, h2d, q.nbytes / h2d / 1e6, out.nbytes >> 20, d2h, out.nbytes / d2h / 1e6, pcie_floor_ms(q, out, dev)))
""" % (ROOT, CACHE)


def run(code, **env):
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env=dict(os.environ, **{"PICO_B200_" + k: str(v) for k, v in env.items()}))
    return (r.stdout.strip() + ("\n" + r.stderr.strip()[-2600:] if r.stderr.strip() else "")).strip()


if __name__ == "__main__":
    from pico_tree_b200 import datasets as D
    tree_pts, q = D.bench_clouds()
    np.savez(CACHE, tree=tree_pts, q=q)
    print(run(FLOOR), flush=True)
    timelines = ({"HOST_STREAMS": 8}, {"HOST_STREAMS": 8, "HOST_PRIO": 0})
    if len(sys.argv) > 1:   # python profiles/host_pipeline_sweep.py KEY=VAL,KEY=VAL ...   ("-" = library defaults)
        # (an argument that starts with "T:" asks for the per-chunk timeline of that point instead of its timing)
        parse = lambda p: dict(kv.split("=") for kv in p.split(",") if kv and kv != "-")
        points = [parse(p) for p in sys.argv[1:] if not p.startswith("T:")]
        timelines = [parse(p[2:]) for p in sys.argv[1:] if p.startswith("T:")]
    else:
        points = [
            {},
            {"HOST_PRIO": 0},
            {"HOST_TAPER": 131072},
            {"HOST_TAPER": 262144},
            {"HOST_CHUNK": 786432},
            {"HOST_CHUNK": 524288},
            {"HOST_CHUNK": 524288, "HOST_STREAMS": 8},
            {"HOST_CHUNK": 1572864, "HOST_TAPER": 262144},
            {"HOST_STREAMS": 4},
            {"HOST_STREAMS": 8},
            {"HOST_HEAD": 131072},
        ]
    for env in points:
        print("%-95s %s" % (" ".join("%s=%s" % kv for kv in env.items()) or "(library defaults)", run(ONE, **env)),
              flush=True)
    for env in timelines:
        print("timeline %s\n%s" % (env, run(ONE, TIMELINE=1, **env)[-2600:]), flush=True)

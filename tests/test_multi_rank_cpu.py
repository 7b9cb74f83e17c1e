"""N > 1 host logic on CPU: world_size 2 over gloo (the GPU path uses the same code with NCCL).
Query sharding covers every query once, the tree-image broadcast delivers identical bytes to
every rank, ragged shard results are rebased correctly, and bench.py's reference arm keeps
quiet on ranks != 0."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, hashlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from pico_tree_b200 import distributed as pd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
payload = None
if rank == 0:
    rng = np.random.default_rng(5)
    payload = torch.from_numpy(rng.integers(0, 256, 1_000_003, dtype=np.uint8))
buf = pd.broadcast_bytes(payload, 0)
digest = hashlib.sha256(buf.numpy().tobytes()).hexdigest()
b, e = pd.shard_range(7_200_863, rank, world)
counts = torch.tensor([e - b], dtype=torch.int64)
dist.all_reduce(counts)
# max-over-ranks timing reduction used by bench.py
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
open(os.path.join(%r, "result_%%d.txt" %% rank), "w").write(
    " ".join(map(str, ["RESULT", rank, digest, b, e, int(counts.item()), float(t.item())])))
dist.barrier()
dist.destroy_process_group()
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [(tmp_path / ("result_%d.txt" % r)).read_text().split() for r in range(2)]
    assert len(rows) == 2
    rows.sort(key=lambda r: int(r[1]))
    assert rows[0][2] == rows[1][2]                       # same bytes everywhere
    assert int(rows[0][3]) == 0 and int(rows[0][4]) == int(rows[1][3]) and int(rows[1][4]) == 7_200_863
    assert int(rows[0][5]) == 7_200_863                   # every query exactly once
    assert float(rows[0][6]) == 2.0 and float(rows[1][6]) == 2.0


def test_shard_range_edge_cases():
    from pico_tree_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 9, 7_200_863):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(b <= e for b, e in spans)


def test_rebase_ragged():
    from pico_tree_b200.distributed import rebase_ragged
    o, f = rebase_ragged([np.array([0, 2, 2, 5]), np.array([0, 1]), np.array([0, 0, 3])],
                         [np.arange(5), np.arange(1) + 10, np.arange(3) + 20])
    assert o.tolist() == [0, 2, 2, 5, 6, 6, 9]
    assert f.tolist() == [0, 1, 2, 3, 4, 10, 20, 21, 22]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--n-tree", "1000", "--n-query", "1000"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""

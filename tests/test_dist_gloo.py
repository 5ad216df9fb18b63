"""N > 1 host logic on CPU: world_size-2 gloo run of the stream sharding + count all-reduce that
bench.py / BatchedTracker.all_reduce_counts perform over NCCL on the GPU box."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DD_ROOT"])
from deepdish_b200.sharding import shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
S, C = 37, 3
lo, hi = shard_range(S, rank, world)
# per-stream counters of the whole job (same on every rank), this rank owns streams [lo, hi)
g = torch.Generator().manual_seed(5)
counts = torch.randint(0, 9, (S, C, 4), generator=g, dtype=torch.int64)
local = counts[lo:hi].sum(0)
dist.all_reduce(local, op=dist.ReduceOp.SUM)
assert torch.equal(local, counts.sum(0)), (rank, local, counts.sum(0))
# the asynchronous, double-buffered form bench.py uses per tick (BatchedTracker.all_reduce_counts(async_op=True))
bufs, works = [torch.zeros_like(local) for _ in range(2)], [None, None]
for k in range(5):
    i = k & 1
    if works[i] is not None:
        works[i].wait()
    bufs[i].copy_(counts[lo:hi].sum(0) * (k + 1))
    works[i] = dist.all_reduce(bufs[i], op=dist.ReduceOp.SUM, async_op=True)
for w in works:
    w.wait()
assert torch.equal(bufs[0], counts.sum(0) * 5) and torch.equal(bufs[1], counts.sum(0) * 4)
spans = [None] * world
dist.all_gather_object(spans, (lo, hi))
assert spans[0][0] == 0 and spans[-1][1] == S and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
dist.barrier()
if rank == 0:
    print("gloo-ok", spans)
dist.destroy_process_group()
'''


def test_world_size_2_gloo_count_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DD_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "gloo-ok" in out.stdout

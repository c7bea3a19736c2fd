"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding and the all-gather reassembly
used by the agent-sharded forward (disconet_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from disconet_b200 import parallel


def test_shard_rows_partitions():
    for n in (0, 1, 5, 10, 40, 41):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == parallel.row_counts(n, world)
    with pytest.raises(ValueError):
        parallel.shard_rows(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the "global" activation every rank would hold after the gather: rows are agent-major images
        g = torch.Generator().manual_seed(0)
        full = torch.randn(2, n_total, 3, 4, 8, generator=g).to(torch.bfloat16)
        r0, r1 = parallel.shard_rows(n_total, world, rank)
        out = torch.zeros_like(full)
        parallel.all_gather_rows(full[:, r0:r1].contiguous(), out)
        ok = torch.equal(out, full)
        # wrong local row count must be rejected, not silently mis-assembled
        try:
            parallel.all_gather_rows(full[:, :0].contiguous() if r1 - r0 else full[:, :1].contiguous(), out)
            rejected = False
        except ValueError:
            rejected = True
        # flat gradient all-reduce of the scene-sharded training path: mean over ranks, in place
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        parallel.allreduce_mean_(flat, None)
        ok = ok and torch.allclose(flat, torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world))
        q.put((rank, ok, rejected))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 5])   # even split (5+5) and ragged split (3+2)
def test_all_gather_rows_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "gathered activation differs from the global one"
    assert all(r[2] for r in res), "mismatched local row count was not rejected"

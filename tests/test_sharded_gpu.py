"""Agent-sharded forward on 2 GPUs (NCCL) == single-GPU forward, bit for bit (needs >= 2 devices)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    from disconet_b200 import DiscoNet, parallel, synth
    from test_oracle_cpu import _Cfg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        A, B = 5, 2
        m = DiscoNet(_Cfg(), kd_flag=0, num_agent=A)
        m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=51))
        m = m.to(dev).eval()
        bev = synth.synth_bev(A * B, seed=52)
        na_list = [5, 4]
        for b, n in enumerate(na_list):
            for a in range(n, A):
                bev[a * B + b] = 0
        na = torch.tensor([[n] * A for n in na_list])
        T = synth.synth_poses(B, A, num_agent=na_list, seed=53)
        r0, r1 = parallel.shard_rows(A * B, world, rank)
        with torch.no_grad():
            full, _ = m(bev.to(dev), T, na, batch_size=B)
            ok = True
            for _ in range(3):     # 1st call eager, 2nd captures the two CUDA graphs either side of the all-gather, 3rd replays
                loc = m.forward_sharded(bev[r0:r1].to(dev), T, na, batch_size=B)
                torch.cuda.synchronize()
                ok = ok and torch.equal(loc["cls"], full["cls"][r0:r1]) and torch.equal(loc["loc"], full["loc"][r0:r1])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_forward_sharded_matches_single_gpu(cuda_dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res

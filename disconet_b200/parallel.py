"""Multi-GPU partitioning of the DiscoNet path (one process per GPU, torch.distributed for the plumbing).

Two partitions exist:
  * by scene  -- scenes are independent: every rank runs whole scenes, no data-path collective
                 (this is what `bench.py --gpus N` measures, weak scaling);
  * by agent  -- the path's natural partition (north_star): encoder, decoder and heads are per-image,
                 only the fusion needs every agent's collaboration-layer map.  The agent-major image rows
                 n = a*B + b are split contiguously across ranks; each rank encodes its rows, ONE
                 all-gather of x_3 (bf16 hi+lo, 256x32x32 per row) makes all maps visible, each rank
                 fuses and decodes only its own ego rows.  `DiscoNet.forward_sharded` implements it.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `n_rows` for `rank` (first n_rows % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def row_counts(n_rows: int, world: int) -> List[int]:
    return [shard_rows(n_rows, world, r)[1] - shard_rows(n_rows, world, r)[0] for r in range(world)]


def all_gather_rows(local: torch.Tensor, out: torch.Tensor, group=None) -> torch.Tensor:
    """local [parts, n_local, ...] -> out [parts, n_total, ...], ranks concatenated along dim 1 in rank order.

    Equal row counts (the normal case: A*B divisible by the world size): one `all_gather_into_tensor` per precision part
    straight into its final place -- out[p] is the rank-order concatenation of the ranks' local[p] -- with no staging
    copies.  Ragged counts: padded list all_gather + slice copies (works on gloo and nccl)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts, n_total = out.shape[0], out.shape[1]
    counts = row_counts(n_total, world)
    if local.shape[1] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[1]} rows, expected {counts[rank]}")
    nmax = max(counts)
    if nmax == 0:
        return out
    if min(counts) == nmax and local.is_contiguous() and out.is_contiguous():
        for p in range(parts):
            dist.all_gather_into_tensor(out[p], local[p], group=group)
        return out
    pad = local
    if local.shape[1] != nmax:
        pad = torch.zeros((parts, nmax) + tuple(local.shape[2:]), dtype=local.dtype, device=local.device)
        pad[:, :local.shape[1]] = local
    pad = pad.contiguous()
    pieces = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(pieces, pad, group=group)
    off = 0
    for r, c in enumerate(counts):
        out[:, off:off + c] = pieces[r][:, :c]
        off += c
    return out


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over the ranks of one flat buffer (the training runner's gradient buffer: every live parameter
    gradient of a step in ONE ~32 MB message instead of DistributedDataParallel's per-bucket copies and hooks)."""
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if flat.is_cuda:
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:   # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


def enable_grad_allreduce(model, group=None) -> None:
    """Scene-sharded data-parallel training (BASELINE config 3) without DistributedDataParallel: every rank trains on its
    own scenes and `TrainRunner.backward` averages its flat gradient buffer over `group` before handing the gradients to
    autograd.  Call once after construction; parameters must start identical on all ranks (same checkpoint / seed).
    BatchNorm statistics stay per rank, exactly as the reference's nn.DataParallel replicas keep them."""
    model._grad_group = group if group is not None else dist.group.WORLD
    for r in getattr(model, "_runners", {}).values():
        r.grad_group = model._grad_group

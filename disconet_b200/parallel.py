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

    One collective when every rank holds the same number of rows (all_gather_into_tensor over a
    [world, parts, n_local, ...] staging view); padded list all_gather otherwise (works on gloo and nccl).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts, n_total = out.shape[0], out.shape[1]
    counts = row_counts(n_total, world)
    if local.shape[1] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[1]} rows, expected {counts[rank]}")
    nmax = max(counts)
    if nmax == 0:
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

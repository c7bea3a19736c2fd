"""LiDAR -> voxel indices -> dense BEV on the GPU (reference: utils/data_util.py:625-717 `voxelize_occupy`
and the dataset scatter datasets/V2XSimDet.py:293-302), bit-exact with the numpy reference."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import PREC_BF16X3, check, load
from .ops import _require_cuda, _stream_ptr, alloc_act


def _grid(extents: np.ndarray, voxel_size) -> Tuple[np.ndarray, np.ndarray]:
    """min voxel coordinate and grid dims exactly as data_util.py:692-702 computes them (float64)."""
    ext = np.asarray(extents, dtype=np.float64)
    vs = np.asarray(voxel_size, dtype=np.float64)
    if ext.shape != (3, 2):
        raise ValueError("Extents are the wrong shape {}".format(ext.shape))
    mn = np.floor(ext.T[0] / vs)
    mx = np.ceil(ext.T[1] / vs) - 1
    dims = ((mx - mn) + 1).astype(np.int32)
    return mn, dims


def voxelize_occupy(pts: torch.Tensor, voxel_size, extents=None, return_indices: bool = False):
    """GPU `voxelize_occupy`.  pts: CUDA float32 [P, 3..4].  Returns the dense occupancy grid
    (float32 [X, Y, Z]) and, if `return_indices`, the lexicographically sorted unique voxel indices
    ([M, 3] int32) -- the same values, order and dtypes-after-save as the reference.

    `extents` is required (the reference's extents=None branch sizes the grid from the data; every
    call site on the DiscoNet path passes extents: create_data_det.py:326-334,393-420)."""
    if pts.dim() != 2 or pts.shape[1] < 3 or pts.shape[1] > 4:
        raise ValueError("Points have the wrong shape: {}".format(tuple(pts.shape)))
    if extents is None:
        raise NotImplementedError("voxelize_occupy on the GPU needs explicit extents")
    _require_cuda(pts)
    if pts.dtype != torch.float32:
        raise ValueError("points must be float32 (the reference's point clouds are float32)")
    pts = pts.contiguous()
    _, dims = _grid(extents, voxel_size)
    ext = (C.c_double * 6)(*np.asarray(extents, dtype=np.float64).reshape(-1).tolist())
    vs = (C.c_double * 3)(*[float(v) for v in voxel_size])
    cd = (C.c_int * 3)(*[int(d) for d in dims])
    n_bits = int(dims[0]) * int(dims[1]) * int(dims[2])
    dev = pts.device
    bitmap = torch.empty(((n_bits + 31) // 32,), dtype=torch.int32, device=dev)
    cap = min(n_bits, max(int(pts.shape[0]), 1))
    idx = torch.empty((cap, 3), dtype=torch.int32, device=dev)
    n_vox = torch.zeros((1,), dtype=torch.int32, device=dev)
    dense = torch.empty(tuple(int(d) for d in dims), dtype=torch.float32, device=dev)
    check(load().disco_voxelize_occupy(pts.data_ptr(), int(pts.shape[0]), int(pts.shape[1]), ext, vs, cd,
                                       bitmap.data_ptr(), idx.data_ptr(), n_vox.data_ptr(), dense.data_ptr(),
                                       _stream_ptr(dev)), "voxelize_occupy")
    if return_indices:
        m = int(n_vox.item())
        return dense, idx[:m]
    return dense


def voxelize_occupy_batched(pts: torch.Tensor, n_points: torch.Tensor, voxel_size, extents, max_voxels: Optional[int] = None):
    """`voxelize_occupy` (data_util.py:625-717) for a batch of sweeps in three launches.  pts: CUDA float32 [S, P_max, 3..4]
    (rows >= n_points[s] ignored), n_points int32 [S].  Returns (indices [S, M_max, 3] int32 -- per sweep the sorted unique voxel
    indices, padded with -1 -- and n_voxels [S] int32), the input format of `DiscoNet.forward_voxels` / `bev_scatter_batched`."""
    if pts.dim() != 3 or pts.shape[2] < 3 or pts.shape[2] > 4 or pts.dtype != torch.float32:
        raise ValueError("points must be float32 [S, P_max, 3..4] (got {} {})".format(tuple(pts.shape), pts.dtype))
    _require_cuda(pts, n_points)
    pts, n_points = pts.contiguous(), n_points.to(torch.int32).contiguous()
    _, dims = _grid(extents, voxel_size)
    S, p_max = int(pts.shape[0]), int(pts.shape[1])
    n_bits = int(dims[0]) * int(dims[1]) * int(dims[2])
    n_words = (n_bits + 31) // 32
    m_max = int(max_voxels) if max_voxels else min(n_bits, max(p_max, 1))
    dev = pts.device
    bitmap = torch.empty((S, n_words), dtype=torch.int32, device=dev)
    block_count = torch.empty((S, (n_words + 1023) // 1024), dtype=torch.int32, device=dev)
    idx = torch.empty((S, m_max, 3), dtype=torch.int32, device=dev)
    n_vox = torch.empty((S,), dtype=torch.int32, device=dev)
    ext = (C.c_double * 6)(*np.asarray(extents, dtype=np.float64).reshape(-1).tolist())
    vs = (C.c_double * 3)(*[float(v) for v in voxel_size])
    cd = (C.c_int * 3)(*[int(d) for d in dims])
    check(load().disco_voxelize_occupy_batched(pts.data_ptr(), n_points.data_ptr(), S, p_max, int(pts.shape[2]), ext, vs, cd,
                                               bitmap.data_ptr(), block_count.data_ptr(), idx.data_ptr(), m_max, n_vox.data_ptr(),
                                               _stream_ptr(dev)), "voxelize_occupy_batched")
    return idx, n_vox


def bev_scatter(voxel_indices: torch.Tensor, dims, *, packed: bool = False, precision: int = PREC_BF16X3):
    """Dataset scatter: indices [M,3] -> dense BEV float32 [Y, X, Z] = np.rot90(vox, 3) with vox[idx]=1
    (V2XSimDet.py:293-302).  With `packed=True` also returns the 16-channel NHWC activation the encoder
    consumes, skipping the fp32 round trip."""
    _require_cuda(voxel_indices)
    if voxel_indices.dtype != torch.int32:
        voxel_indices = voxel_indices.to(torch.int32)
    voxel_indices = voxel_indices.contiguous()
    dx, dy, dz = (int(d) for d in dims)
    dev = voxel_indices.device
    cd = (C.c_int * 3)(dx, dy, dz)
    bev = torch.empty((dy, dx, dz), dtype=torch.float32, device=dev)
    act = alloc_act(1, dy, dx, 16, precision, dev) if packed else None
    if act is not None and act.shape[0] == 2:
        act[1].zero_()
    check(load().disco_bev_scatter(voxel_indices.data_ptr(), int(voxel_indices.shape[0]), cd, bev.data_ptr(),
                                   act.data_ptr() if act is not None else None, 16, precision,
                                   _stream_ptr(dev)), "bev_scatter")
    return (bev, act) if packed else bev


def bev_scatter_batched(voxel_indices: torch.Tensor, counts: torch.Tensor, dims, act: torch.Tensor, precision: int = PREC_BF16X3,
                        lo_nonzero: Optional[torch.Tensor] = None):
    """Dataset scatter of a whole batch (V2XSimDet.py:293-302 per agent) straight into the encoder's input activation:
    voxel_indices [N, M_max, 3] int32 (x, y, z), counts [N] int32 -> act [parts, N, Y, X, 16] (zeroed, then act[a, y, X-1-x, z] = 1)."""
    _require_cuda(voxel_indices, counts, act)
    if voxel_indices.dtype != torch.int32 or counts.dtype != torch.int32:
        raise ValueError("voxel indices / counts must be int32")
    voxel_indices, counts = voxel_indices.contiguous(), counts.contiguous()
    dx, dy, dz = (int(d) for d in dims)
    n, m_max = voxel_indices.shape[0], voxel_indices.shape[1]
    if tuple(act.shape[1:]) != (n, dy, dx, 16):
        raise ValueError(f"activation buffer {tuple(act.shape)} does not match {n} agents of {dy}x{dx}x16")
    cd = (C.c_int * 3)(dx, dy, dz)
    lo_off = act.stride(0) if act.shape[0] == 2 else 0
    check(load().disco_bev_scatter_batched(voxel_indices.data_ptr(), counts.data_ptr(), n, m_max, cd, act.data_ptr(), lo_off, 16,
                                           precision, lo_nonzero.data_ptr() if lo_nonzero is not None else None,
                                           _stream_ptr(act.device)), "bev_scatter_batched")
    return act


def bev_to_voxel_indices(bev: torch.Tensor):
    """Inverse of the dataset scatter for host-side synthetic data: dense BEV [N, 1, Y, X, Z] (0/1) -> (indices [N, M_max, 3]
    int32 padded with -1, counts [N] int32) with (x, y, z) such that bev[a, 0, y, X-1-x, z] = 1 (V2XSimDet.py:293-302)."""
    b = bev[:, 0] if bev.dim() == 5 else bev
    n, Y, X, Z = b.shape
    nz = torch.nonzero(b > 0)                      # [K, 4] = (a, r, c, z) sorted by a
    counts = torch.bincount(nz[:, 0], minlength=n).to(torch.int32)
    m_max = int(counts.max()) if nz.numel() else 0
    idx = torch.full((n, max(m_max, 1), 3), -1, dtype=torch.int32)
    start = torch.cumsum(counts, 0) - counts
    pos = torch.arange(nz.shape[0]) - start[nz[:, 0]].long()
    idx[nz[:, 0], pos, 0] = (X - 1 - nz[:, 2]).to(torch.int32)
    idx[nz[:, 0], pos, 1] = nz[:, 1].to(torch.int32)
    idx[nz[:, 0], pos, 2] = nz[:, 3].to(torch.int32)
    return idx, counts

"""CPU oracle (numpy) for the two integer rows of the hot path (TEST INFRASTRUCTURE, never shipped):

  voxelize_occupy : R/utils/data_util.py:625-717   (points -> sorted unique voxel indices + dense grid)
  bev_scatter     : R/datasets/V2XSimDet.py:293-302 (indices -> dense bool grid -> np.rot90(.,3) -> float32)

Pinned against the live reference function by oracle/make_golden.py (tests/golden/voxel_*.npz).
"""
import numpy as np


def voxelize_occupy(pts, voxel_size, extents):
    """Returns (leaf_layout float32 [X,Y,Z], voxel_indices int [M,3]) exactly as the reference does."""
    pts = np.asarray(pts)
    ext = np.asarray(extents, dtype=np.float64)
    vs = np.asarray(voxel_size, dtype=np.float64)
    # strict bounds (data_util.py:657-664); f32 points are compared against f64 extents
    keep = ((ext[0, 0] < pts[:, 0]) & (pts[:, 0] < ext[0, 1]) & (ext[1, 0] < pts[:, 1]) & (pts[:, 1] < ext[1, 1])
            & (ext[2, 0] < pts[:, 2]) & (pts[:, 2] < ext[2, 1]))
    p = pts[keep]
    # float32 / float64 -> float64 division, floor, int32 (data_util.py:668)
    disc = np.floor(p[:, :3].astype(np.float64) / vs).astype(np.int32)
    # lexsort x, then y, then z + unique == sort rows lexicographically and drop duplicates (:671-690)
    if len(disc):
        disc = np.unique(disc, axis=0)
    mn = np.floor(ext.T[0] / vs)
    mx = np.ceil(ext.T[1] / vs) - 1
    dims = ((mx - mn) + 1).astype(np.int32)
    idx = (disc - mn).astype(int)
    grid = np.zeros(dims.astype(int), dtype=np.float32)
    grid[idx[:, 0], idx[:, 1], idx[:, 2]] = 1.0
    return grid, idx


def bev_scatter(indices, dims):
    """indices [M,3] -> float32 [Y, X, Z] (V2XSimDet.py:293-302 with num_past_pcs == 1, seq dim dropped)."""
    vox = np.zeros(tuple(dims), dtype=bool)
    vox[indices[:, 0], indices[:, 1], indices[:, 2]] = 1
    return np.rot90(vox, 3).astype(np.float32)


def synth_points(agent: int, n_points: int = 40000, rsu: bool = False):
    """Seeded synthetic LiDAR sweep: x,y ~ U(-40,40), z ~ U(-3.5,2.5), intensity ~ U(0,1)."""
    rng = np.random.default_rng(1000 + agent)
    xy = rng.uniform(-40, 40, (n_points, 2))
    z = rng.uniform(-8.5, -2.5, (n_points, 1)) if rsu else rng.uniform(-3.5, 2.5, (n_points, 1))
    inten = rng.uniform(0, 1, (n_points, 1))
    return np.concatenate([xy, z, inten], 1).astype(np.float32)


VOXEL_SIZE = (0.25, 0.25, 0.4)                                        # Config.py:76
EXTENTS = np.array([[-32.0, 32.0], [-32.0, 32.0], [-3.0, 2.0]])       # Config.py:78-84 (vehicle)
EXTENTS_RSU = np.array([[-32.0, 32.0], [-32.0, 32.0], [-8.0, -3.0]])  # cross-road / RSU

"""Seeded synthetic parameters / inputs for the DiscoNet path, shared by bench.py, the
tests and the oracle.  numpy Generator streams are stable across platforms/versions (unlike
torch.manual_seed), so the GPU box regenerates exactly the tensors the golden fixtures were made from.
Pure data generation -- no model compute happens here."""
import torch

def synth_state_dict(template: dict, seed: int = 0) -> dict:
    """Deterministic values for every entry of a DiscoNet/FaFNet state_dict (shapes from `template`).

    conv weights ~ U(+-sqrt(6/fan_in)) (He-uniform, keeps activations O(1) through 20+ layers),
    conv biases ~ U(+-1/sqrt(fan_in)); BN weight ~ U(.5,1.5), bias ~ N(0,.1),
    running_mean ~ N(0,.1), running_var ~ U(.5,1.5) so BN folding is genuinely exercised.
    """
    import numpy as np
    rng = np.random.default_rng(seed)
    out = {}
    for k in sorted(template.keys()):
        shape = tuple(template[k].shape)
        if k.endswith("num_batches_tracked"):
            v = np.zeros(shape, dtype=np.int64)
        elif k.endswith("running_mean"):
            v = rng.normal(0, 0.1, shape).astype(np.float32)
        elif k.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif len(shape) == 1 and (k.rsplit(".", 1)[0] + ".running_mean") in template:   # BatchNorm affine
            v = (rng.uniform(0.5, 1.5, shape) if k.endswith("weight") else rng.normal(0, 0.1, shape)).astype(np.float32)
        else:
            if len(shape) > 1:
                fan_in = int(np.prod(shape[1:]))
            else:  # conv bias: fan_in of its weight
                wshape = tuple(template[k[:-4] + "weight"].shape)
                fan_in = int(np.prod(wshape[1:]))
            bound = np.sqrt(6.0 / fan_in) if len(shape) > 1 else 1.0 / np.sqrt(fan_in)
            v = rng.uniform(-bound, bound, shape).astype(np.float32)
        out[k] = torch.from_numpy(v)
    return out


def synth_poses(B, A, num_agent=None, seed=7):
    """trans_matrices [B,A,A,4,4] float64 with T[b,x,y] = inv(P_x) @ P_y; absent agents -> zeros."""
    import numpy as np
    T = np.zeros((B, A, A, 4, 4), dtype=np.float64)
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        n = A if num_agent is None else int(num_agent[b])
        P = []
        for a in range(A):
            x, y = rng.uniform(-20, 20, 2)
            yaw = rng.uniform(-np.pi, np.pi)
            M = np.eye(4)
            M[0, 0], M[0, 1], M[1, 0], M[1, 1] = np.cos(yaw), -np.sin(yaw), np.sin(yaw), np.cos(yaw)
            M[0, 3], M[1, 3] = x, y
            P.append(M)
        for x_ in range(n):
            for y_ in range(n):
                T[b, x_, y_] = np.linalg.inv(P[x_]) @ P[y_]
    return torch.from_numpy(T)


def synth_bev(N, H=256, W=256, Z=13, occupancy=0.03, seed=0):
    """Dense occupancy input [N,1,H,W,Z] float32 of 0/1 (numpy-seeded)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    return torch.from_numpy((rng.random((N, 1, H, W, Z)) < occupancy).astype(np.float32))


# binary detection config of the reference (configs/Config.py:154-163): (w, l, yaw) of the 6 anchors per BEV cell
ANCHOR_SIZE = ((2.0, 4.0, 0.0), (2.0, 4.0, 1.5707963267948966), (2.0, 4.0, -0.7853981633974483),
               (3.0, 12.0, 0.0), (3.0, 12.0, 1.5707963267948966), (3.0, 12.0, -0.7853981633974483))


def synth_anchors(H=256, W=256, extent=32.0, voxel=0.25, anchor_size=ANCHOR_SIZE):
    """Anchor map [H, W, A, 6] float32 = (x, y, w, h, sin, cos) laid out like obj_util.init_anchors_no_check (:611-633):
    cell (i, j) is centred at (j*voxel - extent + voxel/2, i*voxel - extent + voxel/2)."""
    import numpy as np
    a = np.asarray(anchor_size, dtype=np.float64)
    m = np.zeros((H, W, len(a), 6))
    m[..., 2:4] = a[:, :2]
    m[..., 4] = np.sin(a[:, 2])
    m[..., 5] = np.cos(a[:, 2])
    m[..., 0] = (np.arange(W) * voxel - extent + voxel / 2.0)[None, :, None]
    m[..., 1] = (np.arange(H) * voxel - extent + voxel / 2.0)[:, None, None]
    return torch.from_numpy(m.astype(np.float32))

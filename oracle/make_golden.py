"""Generate tests/golden/* from the LIVE reference (build container only; needs /root/reference).

    python -m oracle.make_golden

The reference's own tests pin nothing on this path, so these fixtures -- outputs of the
unmodified reference classes/functions on numpy-seeded inputs -- are the pin for the oracle restatement
(tests/test_oracle_cpu.py) and therefore for the CUDA path.  Outputs are strided subsamples + statistics
to keep the fixtures small; inputs are regenerated from seeds (oracle.disconet_oracle.synth_*).
"""
import json
import os

import numpy as np
import torch

from oracle import disconet_oracle as O
from oracle import ref_import, voxel_oracle as V

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
STRIDES = {"cls": 997, "loc": 2999, "x_8": 1999, "x_7": 997, "x_6": 499, "x_5": 251, "fused": 127}

# name -> kwargs ; every case is regenerated from seeds by tests via `golden_case_inputs`
DISCO_CASES = {
    "disco_a2_b1": dict(A=2, B=1, num_agent=[2], kd_flag=1, only_v2i=False, compress_level=0, seed=11),
    "disco_a3_b2_absent": dict(A=3, B=2, num_agent=[3, 2], kd_flag=0, only_v2i=False, compress_level=0, seed=12),
    "disco_a3_b1_v2i_comp": dict(A=3, B=1, num_agent=[3], kd_flag=1, only_v2i=True, compress_level=2, seed=13),
    "disco_a2_b1_layer2": dict(A=2, B=1, num_agent=[2], kd_flag=1, only_v2i=False, compress_level=0, seed=14, layer=2),
    # the headline shape (BASELINE configs[1]: 5 agents) from the live reference: one scene, and two scenes with an absent agent
    "disco_a5_b1": dict(A=5, B=1, num_agent=[5], kd_flag=0, only_v2i=False, compress_level=0, seed=15),
    "disco_a5_b2": dict(A=5, B=2, num_agent=[5, 4], kd_flag=1, only_v2i=False, compress_level=0, seed=16),
}


# training-mode cases (a12): forward with batch-statistics BN + backward of `probe_loss` through the live reference
TRAIN_CASES = {
    "train_a2_b1": dict(A=2, B=1, num_agent=[2], kd_flag=1, only_v2i=False, compress_level=0, seed=31),
    "train_a3_b1_absent": dict(A=3, B=1, num_agent=[2], kd_flag=1, only_v2i=False, compress_level=0, seed=32),
}
TRAIN_OUT_KEYS = ("cls", "loc", "x_8", "x_7", "x_6", "x_5", "fused")


# BEV segmentation DiscoNet (f1 / BASELINE config 5), eval forward through the live reference
SEG_CASES = {
    "seg_a2_b1": dict(A=2, B=1, num_agent=[2], only_v2i=False, seed=51),
    "seg_a4_b1_absent_v2i": dict(A=4, B=1, num_agent=[3], only_v2i=True, seed=52),
    "seg_a2_b1_comp2": dict(A=2, B=1, num_agent=[2], only_v2i=False, seed=54, compress_level=2),   # 512 -> 128 -> 512 bottleneck
}
SEG_TRAIN_CASE = dict(A=2, B=1, num_agent=[2], only_v2i=False, seed=53)
SEG_KEYS = ("logits", "x9", "x8", "x7", "x6", "x5", "feat")
SEG_STRIDES = {"logits": 499, "x9": 1999, "x8": 1999, "x7": 997, "x6": 499, "x5": 251, "feat": 251}


def seg_case_inputs(case: dict, template_sd: dict):
    A, B = case["A"], case["B"]
    sd = O.synth_state_dict(template_sd, seed=case["seed"])
    bev = O.synth_bev(A * B, seed=case["seed"] + 100)[:, 0].permute(0, 3, 1, 2).contiguous()   # [N, 13, H, W]
    na = torch.tensor([[n] * A for n in case["num_agent"]])
    for b, n in enumerate(case["num_agent"]):
        for a in range(n, A):
            bev[a * B + b] = 0
    T = O.synth_poses(B, A, num_agent=case["num_agent"], seed=case["seed"] + 200)
    return sd, bev, T, na


def grad_digest(named: dict, stride: int = 499):
    """{name: tensor} -> flat strided subsample + per-tensor [l2 norm, max|.|] table (sorted by name)."""
    subs, table = [], []
    for k in sorted(named):
        f = named[k].detach().reshape(-1).double()
        subs.append(f[::stride].float().numpy())
        table.append([f.norm().item(), f.abs().max().item() if f.numel() else 0.0])
    return np.concatenate(subs), np.array(table)


def golden_case_inputs(case: dict, template_sd: dict):
    A, B = case["A"], case["B"]
    sd = O.synth_state_dict(template_sd, seed=case["seed"])
    bev = O.synth_bev(A * B, seed=case["seed"] + 100)
    na = torch.tensor([[n] * A for n in case["num_agent"]])
    for b, n in enumerate(case["num_agent"]):
        for a in range(n, A):  # absent agents: zero BEV (V2XSimDet.py:210-255) and zero matrices
            bev[a * B + b] = 0
    T = O.synth_poses(B, A, num_agent=case["num_agent"], seed=case["seed"] + 200)
    return sd, bev, T, na


def _sub(t: torch.Tensor, stride: int):
    f = t.detach().reshape(-1).double()
    return f[::stride].float().numpy(), np.array([f.sum().item(), f.abs().sum().item(), f.abs().max().item()])


def main(only=None):
    """`only`: iterable of DISCO_CASES names -> (re)generate just those eval fixtures (python -m oracle.make_golden name ...)."""
    os.makedirs(OUT, exist_ok=True)
    RDisco, RFaF, RTeach, Config = ref_import.reference_classes()
    cfg = Config("train", binary=True, only_det=True)
    keys = {}
    kpath = os.path.join(OUT, "state_dict_keys.json")
    if only and os.path.exists(kpath):
        with open(kpath) as f:
            keys = json.load(f)
    for name, case in DISCO_CASES.items():
        if only and name not in only:
            continue
        m = RDisco(cfg, layer=case.get("layer", 3), kd_flag=case["kd_flag"], num_agent=case["A"], compress_level=case["compress_level"],
                   only_v2i=case["only_v2i"]).eval()
        keys[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        sd, bev, T, na = golden_case_inputs(case, m.state_dict())
        m.load_state_dict(sd)
        with torch.no_grad():
            out = m(bev, T, na, batch_size=case["B"])
        rec = {}
        tensors = {"cls": out[0]["cls"], "loc": out[0]["loc"]}
        if case["kd_flag"] == 1:
            tensors.update(x_8=out[1], x_7=out[2], x_6=out[3], x_5=out[4], fused=out[5])
        else:
            wl = out[1]
            rec["n_weight_entries"] = np.array([len(wl)] + [len(e) for e in wl])
            rec["weights_cat"] = torch.cat([torch.stack(e).reshape(-1) for e in wl]).numpy()[::37]
        for k, t in tensors.items():
            rec[k + "_sub"], rec[k + "_stats"] = _sub(t, STRIDES[k])
            rec[k + "_shape"] = np.array(t.shape)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, {k: v.shape for k, v in rec.items()})

    for name, case in (TRAIN_CASES.items() if not only else ()):
        m = RDisco(cfg, layer=3, kd_flag=1, num_agent=case["A"], compress_level=0, only_v2i=case["only_v2i"]).train()
        keys[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        sd, bev, T, na = golden_case_inputs(case, m.state_dict())
        m.load_state_dict(sd)
        out = m(bev, T, na, batch_size=case["B"])
        tensors = dict(zip(TRAIN_OUT_KEYS, (out[0]["cls"], out[0]["loc"]) + tuple(out[1:])))
        loss, _ = O.probe_loss(tensors, seed=case["seed"] + 300)
        loss.backward()
        rec = {"loss": np.array([loss.item()])}
        for k, t in tensors.items():
            rec[k + "_sub"], rec[k + "_stats"] = _sub(t, STRIDES[k])
            rec[k + "_shape"] = np.array(t.shape)
        grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}
        rec["grad_sub"], rec["grad_table"] = grad_digest(grads)
        rec["grad_names"] = np.array(sorted(grads))
        rec["grad_none"] = np.array([k for k, p in m.named_parameters() if p.grad is None])
        bufs = {k: v.float() for k, v in m.named_buffers()}
        rec["buf_sub"], rec["buf_table"] = grad_digest(bufs, stride=7)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "loss", loss.item(), {k: v.shape for k, v in rec.items()})

    from coperception.models.seg.DiscoNet import DiscoNet as RSeg
    for name, case in SEG_CASES.items():
        if only and name not in only:
            continue
        m = RSeg(13, 8, num_agent=case["A"], kd_flag=True, only_v2i=case["only_v2i"], compress_level=case.get("compress_level", 0)).eval()
        keys[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        sd, bev, T, na = seg_case_inputs(case, m.state_dict())
        m.load_state_dict(sd)
        with torch.no_grad():
            out = m(bev, T, na)
        rec = {}
        for k, t in zip(SEG_KEYS, out):
            rec[k + "_sub"], rec[k + "_stats"] = _sub(t, SEG_STRIDES[k])
            rec[k + "_shape"] = np.array(t.shape)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, {k: v.shape for k, v in rec.items() if k.endswith("_shape")})

    if only:
        with open(kpath, "w") as f:
            json.dump(keys, f)
        return
    # seg DiscoNet in train() mode: outputs + parameter-gradient digest + updated BatchNorm buffers
    name, case = "seg_train_a2_b1", SEG_TRAIN_CASE
    m = RSeg(13, 8, num_agent=case["A"], kd_flag=True, only_v2i=case["only_v2i"]).train()
    keys[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    sd, bev, T, na = seg_case_inputs(case, m.state_dict())
    m.load_state_dict(sd)
    out = m(bev, T, na)
    tensors = dict(zip(SEG_KEYS, out))
    loss, _ = O.probe_loss(tensors, seed=case["seed"] + 300)
    loss.backward()
    rec = {"loss": np.array([loss.item()])}
    for k, t in tensors.items():
        rec[k + "_sub"], rec[k + "_stats"] = _sub(t, SEG_STRIDES[k])
        rec[k + "_shape"] = np.array(t.shape)
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}
    rec["grad_sub"], rec["grad_table"] = grad_digest(grads)
    rec["grad_names"] = np.array(sorted(grads))
    rec["grad_none"] = np.array([k for k, p in m.named_parameters() if p.grad is None])
    rec["buf_sub"], rec["buf_table"] = grad_digest({k: v.float() for k, v in m.named_buffers()}, stride=7)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "loss", loss.item())

    # FaFNet lower-bound plumbing config (BASELINE config 1): 2 agents, 128x128x13
    m = RFaF(cfg, kd_flag=0, num_agent=2).eval()
    keys["fafnet_a2_128"] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    sd = O.synth_state_dict(m.state_dict(), seed=21)
    m.load_state_dict(sd)
    bev = O.synth_bev(2, H=128, W=128, seed=121)
    with torch.no_grad():
        res = m(bev)
    rec = {}
    for k in ("cls", "loc"):
        rec[k + "_sub"], rec[k + "_stats"] = _sub(res[k], 97)
        rec[k + "_shape"] = np.array(res[k].shape)
    np.savez_compressed(os.path.join(OUT, "fafnet_a2_128.npz"), **rec)

    t = RTeach(cfg)
    keys["teacher"] = [[k, list(v.shape)] for k, v in t.state_dict().items()]
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f)

    # voxelize_occupy: the reference function itself on seeded sweeps (vehicle + RSU extents, ragged + empty)
    ref_import.install_bypass(mock_heavy=True)
    from coperception.utils.data_util import voxelize_occupy as ref_vox
    rec = {}
    for tag, pts, ext in [("veh", V.synth_points(1, 40000), V.EXTENTS), ("rsu", V.synth_points(0, 30000, rsu=True), V.EXTENTS_RSU),
                          ("tiny", V.synth_points(3, 7), V.EXTENTS), ("xyz_only", V.synth_points(2, 5000)[:, :3], V.EXTENTS)]:
        grid, idx = ref_vox(pts, V.VOXEL_SIZE, ext, return_indices=True)
        rec[tag + "_idx"] = idx.astype(np.int32)
        rec[tag + "_grid_sum"] = np.array([grid.sum(), grid.shape[0], grid.shape[1], grid.shape[2]])
        # dataset scatter exactly as V2XSimDet.py:293-302 writes it
        cur = np.zeros(grid.shape, dtype=bool)
        cur[idx[:, 0], idx[:, 1], idx[:, 2]] = 1
        bevd = np.rot90(cur, 3).astype(np.float32)
        rec[tag + "_bev_nz"] = np.argwhere(bevd > 0).astype(np.int16)
    # boundary torture: points exactly on / next to extents and voxel edges
    edge = np.array([[-32.0, 0, 0, 0], [32.0, 0, 0, 0], [31.999998, 31.999998, 1.9999999, 0], [-31.999998, -31.999998, -2.9999998, 0],
                     [0.25, 0.5, 0.4, 0], [0.24999999, 0.49999997, 0.39999998, 0], [0, 0, 0.8, 0], [0, 0, 1.2, 0], [0, 0, 1.6, 0],
                     [0, 0, -0.4, 0], [0, 0, -1.2, 0], [0, 0, -2.8, 0], [0, 0, 2.0, 0], [0, 0, -3.0, 0]], dtype=np.float32)
    rng = np.random.default_rng(5)
    zs = (np.arange(-7, 5)[:, None] * 0.4 + rng.uniform(-2e-6, 2e-6, (12, 50))).reshape(-1)
    edge2 = np.stack([rng.uniform(-32, 32, zs.size), rng.uniform(-32, 32, zs.size), zs, np.zeros_like(zs)], 1).astype(np.float32)
    pts = np.concatenate([edge, edge2])
    grid, idx = ref_vox(pts, V.VOXEL_SIZE, V.EXTENTS, return_indices=True)
    rec["edge_pts"] = pts
    rec["edge_idx"] = idx.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "voxel.npz"), **rec)
    print("voxel", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    import sys
    main(only=sys.argv[1:] or None)

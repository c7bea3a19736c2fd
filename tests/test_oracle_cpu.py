"""CPU suite: the oracle restatement against the golden fixtures generated from the live reference,
the host-side logic (BN folding, operand packing, state_dict layout) and the C-ABI surface."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import disconet_oracle as O
from oracle import voxel_oracle as V
from oracle.make_golden import DISCO_CASES, STRIDES, golden_case_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class _Cfg:
    """The fields of coperception Config the model constructors read (Config.py:70-177 defaults)."""
    motion_state = False
    only_det = True
    pred_len = 1
    box_code_size = 6
    category_num = 2
    use_map = False
    use_vis = False
    binary = True
    anchor_size = np.zeros((6, 3))
    map_dims = [256, 256, 13]


def _template(case_name):
    with open(os.path.join(GOLD, "state_dict_keys.json")) as f:
        keys = json.load(f)[case_name]
    return {k: torch.empty(shape) for k, shape in keys}


def _check_sub(name, t, rec, key, tol=2e-5, stride=None):
    f = t.detach().reshape(-1).double()
    sub = f[::(stride or STRIDES[key])].float().numpy()
    ref = rec[key + "_sub"]
    assert sub.shape == ref.shape, (name, key)
    scale = rec[key + "_stats"][2]
    assert np.abs(sub - ref).max() <= tol * scale, f"{name}.{key}: {np.abs(sub - ref).max() / scale:.2e}"
    st = rec[key + "_stats"]
    assert abs(f.abs().sum().item() - st[1]) <= 1e-4 * st[1]
    assert list(t.shape) == list(rec[key + "_shape"])


@pytest.mark.parametrize("name", list(DISCO_CASES))
def test_oracle_matches_reference_golden(name):
    case = DISCO_CASES[name]
    rec = np.load(os.path.join(GOLD, name + ".npz"))
    sd, bev, T, na = golden_case_inputs(case, _template(name))
    out = O.disconet_forward(sd, bev, T, na, case["B"], agent_num=case["A"], only_v2i=case["only_v2i"],
                             return_all=True, layer=case.get("layer", 3))
    _check_sub(name, out["cls"], rec, "cls")
    _check_sub(name, out["loc"], rec, "loc")
    if case["kd_flag"] == 1:
        for k in ("x_8", "x_7", "x_6", "x_5", "fused"):
            _check_sub(name, out[k], rec, k)
    else:
        wl = [e for per_b in out["weights"] for e in per_b]
        assert [len(wl)] + [len(e) for e in wl] == rec["n_weight_entries"].tolist()
        cat = torch.cat([torch.stack(e).reshape(-1) for e in wl]).numpy()[::37]
        assert np.abs(cat - rec["weights_cat"]).max() < 1e-5


def test_oracle_fafnet_golden():
    rec = np.load(os.path.join(GOLD, "fafnet_a2_128.npz"))
    sd = O.synth_state_dict(_template("fafnet_a2_128"), seed=21)
    out = O.fafnet_forward(sd, O.synth_bev(2, H=128, W=128, seed=121))
    for k in ("cls", "loc"):
        _check_sub("fafnet", out[k], rec, k, stride=97)


def test_voxel_oracle_bit_exact_vs_reference_golden():
    rec = np.load(os.path.join(GOLD, "voxel.npz"))
    cases = [("veh", V.synth_points(1, 40000), V.EXTENTS), ("rsu", V.synth_points(0, 30000, rsu=True), V.EXTENTS_RSU),
             ("tiny", V.synth_points(3, 7), V.EXTENTS), ("xyz_only", V.synth_points(2, 5000)[:, :3], V.EXTENTS),
             ("edge", rec["edge_pts"], V.EXTENTS)]
    for tag, pts, ext in cases:
        grid, idx = V.voxelize_occupy(pts, V.VOXEL_SIZE, ext)
        assert np.array_equal(idx.astype(np.int32), rec[tag + "_idx"]), tag
        if tag + "_grid_sum" in rec:
            assert grid.sum() == rec[tag + "_grid_sum"][0] and list(grid.shape) == rec[tag + "_grid_sum"][1:].tolist()
            bev = V.bev_scatter(idx, grid.shape)
            assert np.array_equal(np.argwhere(bev > 0).astype(np.int16), rec[tag + "_bev_nz"]), tag
    # empty cloud and all-out-of-range cloud
    g, i = V.voxelize_occupy(np.zeros((0, 4), np.float32), V.VOXEL_SIZE, V.EXTENTS)
    assert g.sum() == 0 and i.shape == (0, 3)
    g, i = V.voxelize_occupy(np.full((5, 4), 100.0, np.float32), V.VOXEL_SIZE, V.EXTENTS)
    assert g.sum() == 0 and i.shape == (0, 3)


# ------------------------------------------------------------------------------------------------------
# host logic of the product package (no GPU needed)
# ------------------------------------------------------------------------------------------------------
def test_state_dict_layout_matches_reference():
    from disconet_b200 import DiscoNet, FaFNet, TeacherNet
    with open(os.path.join(GOLD, "state_dict_keys.json")) as f:
        keys = json.load(f)
    cfg = _Cfg()
    for name, case in DISCO_CASES.items():
        m = DiscoNet(cfg, layer=case.get("layer", 3), kd_flag=case["kd_flag"], num_agent=case["A"],
                     compress_level=case["compress_level"], only_v2i=case["only_v2i"])
        got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert got == keys[name], name
    assert [[k, list(v.shape)] for k, v in FaFNet(cfg, kd_flag=0, num_agent=2).state_dict().items()] == keys["fafnet_a2_128"]
    assert [[k, list(v.shape)] for k, v in TeacherNet(cfg).state_dict().items()] == keys["teacher"]
    # DataParallel-style "module." prefixed checkpoints load through nn.DataParallel (test_codet.py:164,194)
    m = DiscoNet(cfg, kd_flag=0, num_agent=2)
    sd = {"module." + k: v for k, v in m.state_dict().items()}
    torch.nn.DataParallel(m).load_state_dict(sd)


def test_fold_bn_and_pack_roundtrip():
    from disconet_b200._lib import PREC_BF16X3, PREC_FP16
    from disconet_b200.plan import fold_bn, pack_conv
    g = torch.Generator().manual_seed(3)
    w = torch.randn(48, 40, 3, 3, generator=g)
    b = torch.randn(48, generator=g)
    bn = [torch.rand(48, generator=g) + 0.5, torch.randn(48, generator=g), torch.randn(48, generator=g),
          torch.rand(48, generator=g) + 0.5]
    wf, bf = fold_bn(w, b, *bn)
    x = torch.randn(2, 40, 9, 9, generator=g)
    ref = torch.nn.functional.batch_norm(torch.nn.functional.conv2d(x, w, b, padding=1), bn[2], bn[3], bn[0], bn[1],
                                         False, 0.0, 1e-5)
    got = torch.nn.functional.conv2d(x, wf, bf, padding=1)
    assert (got - ref).abs().max() < 1e-4
    wpad = torch.zeros(48, 48, 3, 3)
    wpad[:, :40] = wf
    for prec in (PREC_FP16, PREC_BF16X3):
        plan = pack_conv(wpad, bf, src_channels=[32, 16], precision=prec, block_n=32, keep_ref=True)
        assert plan.c_blk == 16 and plan.block_n == 32 and plan.bias.numel() == 64
        parts = 2 if prec == PREC_BF16X3 else 1
        dt = torch.bfloat16 if prec == PREC_BF16X3 else torch.float16
        if plan.stacked:
            wp = plan.wpack.view(dt).view(2, 3, 9, 2, parts, 32, 8).float().sum(4)
        else:
            wp = plan.wpack.view(dt).view(2, 3, 9, parts, 2, 32, 8).float().sum(3)   # [nt, cb, tap, chunk, n, 8]
        dec = wp.permute(0, 4, 1, 3, 5, 2).reshape(64, 48, 9)                      # [n, c_in, tap]
        tol = (2.0 ** -16 if prec == PREC_BF16X3 else 2.0 ** -10) * wpad.abs().max().item()
        assert (dec[:48] - wpad.reshape(48, 48, 9)).abs().max() < tol
        assert dec[48:].abs().max() == 0
        assert torch.equal(plan.wref, wpad.reshape(48, 48, 9).permute(0, 2, 1))


def test_agent_weight_list_order():
    from disconet_b200.det import AgentWeightList
    B, A, h, w = 2, 3, 4, 4
    wt = torch.arange(B * A * A, dtype=torch.float32).view(B, A, A, 1, 1).expand(B, A, A, h, w).contiguous()
    wl = AgentWeightList(wt, torch.tensor([3, 2], dtype=torch.int32), only_v2i=False)
    assert len(wl) == 5
    # scene 0, ego 1 -> neighbours [1, 0, 2]
    assert [int(t[0, 0]) for t in wl[1]] == [0 * 9 + 1 * 3 + 1, 0 * 9 + 1 * 3 + 0, 0 * 9 + 1 * 3 + 2]
    # scene 1 has 2 agents: ego 0 -> [0, 1]
    assert [int(t[0, 0]) for t in wl[3]] == [9 + 0, 9 + 1]
    wl = AgentWeightList(wt, torch.tensor([3, 3], dtype=torch.int32), only_v2i=True)
    assert [int(t[0, 0]) for t in wl[1]] == [4, 3]      # ego 1 only hears the RSU (agent 0)
    assert [int(t[0, 0]) for t in wl[0]] == [0, 1, 2]   # the RSU hears everyone


def test_no_cpu_fallback_in_train_and_eval_mode():
    from disconet_b200 import DiscoNet
    m = DiscoNet(_Cfg(), kd_flag=0, num_agent=2)
    bev = torch.zeros(2, 1, 32, 32, 13)
    with pytest.raises(ValueError, match="CUDA"):
        m.train()(bev, torch.zeros(1, 2, 2, 4, 4), torch.ones(1, 2, dtype=torch.int64), batch_size=1)
    with pytest.raises(ValueError, match="CUDA"):
        m.eval()(bev, torch.zeros(1, 2, 2, 4, 4), torch.ones(1, 2, dtype=torch.int64), batch_size=1)


def test_c_abi_exports_every_declared_symbol():
    from disconet_b200 import _lib
    with open(os.path.join(ROOT, "include", "disco_b200.h")) as f:
        declared = set(re.findall(r"^(?:int|long long)\s+(disco_\w+)\s*\(", f.read(), flags=re.M))
    assert declared and declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.disco_version() >= 100
    # struct layouts agree with the header (field order + count)
    with open(os.path.join(ROOT, "include", "disco_b200.h")) as f:
        hdr = f.read()
    for struct, cls in (("disco_conv_desc", _lib.ConvDesc), ("disco_fusion_desc", _lib.FusionDesc),
                        ("disco_grad_src", _lib.GradSrc), ("disco_bn_desc", _lib.BnDesc),
                        ("disco_wgrad_desc", _lib.WgradDesc), ("disco_pack_desc", _lib.PackDesc), ("disco_pwf_train_desc", _lib.PwfTrainDesc)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            names.append(re.findall(r"(\w+)(?:\[\d+\])?$", first.strip())[0])
            names += [re.findall(r"(\w+)", r)[0] for r in rest]
        assert names == [f[0] for f in cls._fields_], (struct, names)


# ---- training mode (a12): oracle restatement (batch-stat BN + torch.autograd) vs the live-reference golden ----
def oracle_train_step(case, sd):
    """Oracle forward in training mode + backward of probe_loss.  Returns (outputs, loss, grads, new buffers)."""
    from oracle.make_golden import TRAIN_OUT_KEYS
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))
              else v.clone()) for k, v in sd.items()}
    _, bev, T, na = golden_case_inputs(case, sd)
    with O.training(sd) as ctx:
        out = O.disconet_forward_graph(sd, bev, T, na, case["B"], agent_num=case["A"], only_v2i=case["only_v2i"],
                                       return_all=True)
    tensors = {k: out[k] for k in TRAIN_OUT_KEYS}
    loss, cot = O.probe_loss(tensors, seed=case["seed"] + 300)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items() if v.requires_grad}
    return tensors, loss, grads, ctx.buffers, cot


@pytest.mark.parametrize("name", ["train_a2_b1", "train_a3_b1_absent"])
def test_oracle_training_matches_reference_golden(name):
    from oracle.make_golden import TRAIN_CASES, grad_digest
    case = TRAIN_CASES[name]
    rec = np.load(os.path.join(GOLD, name + ".npz"))
    sd, *_ = golden_case_inputs(case, _template(name))
    tensors, loss, grads, bufs, _ = oracle_train_step(case, sd)
    assert abs(loss.item() - rec["loss"][0]) <= 1e-3 * max(1.0, abs(rec["loss"][0]))   # sum of ~1e7 signed terms
    for k, t in tensors.items():
        _check_sub(name, t, rec, k, tol=2e-4)   # reference BN kernels: 3e-4 between 1 and 8 threads
    assert sorted(grads) == rec["grad_names"].tolist()
    sub, table = grad_digest(grads)
    ref_table = rec["grad_table"]
    # Gradient tolerance.  Train-mode gradients of this ReLU/BatchNorm stack are ill-conditioned in fp32: a forward
    # perturbation of relative size f flips a fraction ~f of the ReLU gates and moves cancellation-dominated sums
    # by ~sqrt(f).  Measured in this container: the reference against ITSELF at 1 vs 8 CPU threads moves its
    # outputs by 3e-4 and its gradients by 2-8 % (rel-max per tensor); this fp32 oracle against its own float64
    # run: outputs 1e-5, gradients up to 4e-2.  So the pin is: per-tensor l2 norms within 2 %, cosine similarity of
    # the sampled gradient vector >= 0.9995.  (Conv biases in front of a BatchNorm have a mathematically zero
    # gradient -- pure rounding noise on both sides -- and are skipped.)  Exact backward arithmetic is pinned
    # per kernel in tests/test_train_gpu.py against torch.autograd on identical inputs.
    names = sorted(grads)
    for i, k in enumerate(names):
        if k.endswith(".bias") and not (".bn" in k or "bn_" in k or "box_prediction.1" in k or "conv2." in k
                                        or "box_prediction.3" in k or "conv1_4" in k):
            assert table[i, 0] <= 1e-4 * ref_table[:, 0].max(), k     # BN-shadowed conv bias: zero up to noise
            continue
        assert abs(table[i, 0] - ref_table[i, 0]) <= 2e-2 * ref_table[i, 0], (k, table[i, 0], ref_table[i, 0])
    cos = float(np.dot(sub, rec["grad_sub"]) / (np.linalg.norm(sub) * np.linalg.norm(rec["grad_sub"])))
    assert cos >= 0.9995, cos
    bsub, _ = grad_digest({k: v.float() for k, v in bufs.items()}, stride=7)
    assert np.abs(bsub - rec["buf_sub"]).max() <= 2e-4 * np.abs(rec["buf_sub"]).max()


# ---- BEV segmentation DiscoNet (f1 / BASELINE config 5): oracle vs the live-reference golden ----
@pytest.mark.parametrize("name", ["seg_a2_b1", "seg_a4_b1_absent_v2i", "seg_a2_b1_comp2"])
def test_seg_oracle_matches_reference_golden(name):
    from oracle import seg_oracle as S
    from oracle.make_golden import SEG_CASES, SEG_KEYS, SEG_STRIDES, seg_case_inputs
    case = SEG_CASES[name]
    rec = np.load(os.path.join(GOLD, name + ".npz"))
    sd, bev, T, na = seg_case_inputs(case, _template(name))
    out = S.seg_disconet_forward(sd, bev, T, na, agent_num=case["A"], only_v2i=case["only_v2i"], return_all=True)
    for k in SEG_KEYS:
        _check_sub(name, out[k], rec, k, stride=SEG_STRIDES[k])


def test_seg_state_dict_layout_matches_reference():
    from disconet_b200.seg import SegDiscoNet
    with open(os.path.join(GOLD, "state_dict_keys.json")) as f:
        keys = json.load(f)["seg_a2_b1"]
    m = SegDiscoNet(13, 8, num_agent=2)
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == keys
    with pytest.raises(ValueError, match="CUDA"):
        m.eval()(torch.zeros(2, 13, 32, 32), torch.zeros(1, 2, 2, 4, 4), torch.ones(1, 2, dtype=torch.int64))


def test_training_layer_tables_match_the_reference_parameters():
    """Host logic of the training driver: every conv / BatchNorm the layer tables name exists in the reference
    state_dict with the channel counts the tables claim (Backbone.py:11-47), and the set of parameters the autograd node
    differentiates is exactly the set that receives a gradient in the reference (golden `grad_none`)."""
    from disconet_b200.det import runner_param_names
    from disconet_b200.train import backbone_layers
    shapes = dict((k, tuple(s)) for k, s in json.load(open(os.path.join(GOLD, "state_dict_keys.json")))["train_a2_b1"])
    E, D = backbone_layers("u_encoder.", "decoder.", "x3f", 3)
    for L in E + D:
        w = shapes[L.conv + ".weight"]
        assert w[0] == L.c_out and w[1] == L.c_in_real and w[2] * w[3] == L.taps, L.name
        assert shapes[L.bn + ".weight"] == (L.c_out,) and (L.bn + ".running_var") in shapes, L.name
        assert sum(L.c_in) >= L.c_in_real and all(c % 16 == 0 for c in L.c_in), L.name

    class R:
        enc, dec, head_layers, pwf_prefix = E, D, [1], "pixel_weighted_fusion."
    live = runner_param_names(R)
    rec = np.load(os.path.join(GOLD, "train_a2_b1.npz"))
    params = set(rec["grad_names"].tolist())
    assert live == params - set(rec["grad_none"].tolist())


def test_patcher_swaps_the_reference_classes_in_place():
    """The drop-in mechanism against the REAL package layout (build container only: /root/reference is absent on the
    GPU box): after `patch_coperception()`, `from coperception.models.det import *` -- what tools/det/train_codet.py:12
    and test_codet.py:14 do -- yields our classes, constructible with the reference's own Config object."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    ref_import.install_bypass(mock_heavy=True)
    import importlib
    det_pkg = importlib.import_module("coperception.models.det")
    ref_disco = det_pkg.DiscoNet
    originals = {n: getattr(det_pkg, n) for n in ("DiscoNet", "FaFNet", "TeacherNet")}
    from disconet_b200 import patch
    import disconet_b200
    patch.patch_coperception()
    try:
        ns = {}
        exec("from coperception.models.det import *", ns)
        assert ns["DiscoNet"] is disconet_b200.DiscoNet and ns["FaFNet"] is disconet_b200.FaFNet
        assert ns["TeacherNet"] is disconet_b200.TeacherNet
        from coperception.configs.Config import Config
        cfg = Config("train", binary=True, only_det=True)
        m = ns["DiscoNet"](cfg, layer=3, kd_flag=1, num_agent=5, compress_level=0, only_v2i=False)   # train_codet.py:115-123
        r = ref_disco(cfg, layer=3, kd_flag=1, num_agent=5, compress_level=0, only_v2i=False)
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v.shape)) for k, v in r.state_dict().items()]
        assert [k for k, _ in m.named_parameters()] == [k for k, _ in r.named_parameters()]      # Adam state reload order
        m.load_state_dict(r.state_dict())
        # checkpoints written through nn.DataParallel carry the `module.` prefix (test_codet.py:190-196)
        torch.nn.DataParallel(m).load_state_dict({"module." + k: v for k, v in r.state_dict().items()})
        seg_pkg = importlib.import_module("coperception.models.seg")
        assert seg_pkg.DiscoNet is disconet_b200.seg.SegDiscoNet
    finally:
        for name, cls in originals.items():
            setattr(det_pkg, name, cls)
            setattr(importlib.import_module(f"coperception.models.det.{name}"), name, cls)
        seg_mod = importlib.import_module("coperception.models.seg.DiscoNet")
        importlib.reload(seg_mod)
        importlib.import_module("coperception.models.seg").DiscoNet = seg_mod.DiscoNet


def test_post_oracle_matches_reference():
    """f3: the numpy restatement of score / decode / corners against the reference's own functions (build container)."""
    from oracle import post_oracle as P, ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    ref_import.install_bypass(mock_heavy=True)
    from coperception.utils.detection_util import bev_box_decode_torch
    from coperception.utils.obj_util import center_to_corner_box2d
    rng = np.random.default_rng(3)
    enc = (rng.standard_normal((500, 6)) * 0.3).astype(np.float32)
    anc = np.concatenate([rng.uniform(-30, 30, (500, 2)), rng.uniform(1, 5, (500, 2)), rng.uniform(-1, 1, (500, 2))], 1).astype(np.float32)
    dec_ref = bev_box_decode_torch(torch.from_numpy(enc), torch.from_numpy(anc)).numpy()
    dec = P.decode_boxes(enc, anc)
    assert np.abs(dec - dec_ref).max() <= 1e-5 * np.abs(dec_ref).max()
    cor_ref = center_to_corner_box2d(dec_ref[:, :2], dec_ref[:, 2:4], dec_ref[:, 4:])
    assert np.abs(P.box_corners(dec_ref) - cor_ref).max() <= 1e-5 * np.abs(cor_ref).max()


def oracle_seg_train_step(case, sd):
    from oracle import seg_oracle as S
    from oracle.make_golden import SEG_KEYS, seg_case_inputs
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    _, bev, T, na = seg_case_inputs(case, sd)
    with O.training(sd) as ctx:
        out = S.seg_disconet_forward_graph(sd, bev, T, na, agent_num=case["A"], only_v2i=case["only_v2i"], return_all=True)
    tensors = {k: out[k] for k in SEG_KEYS}
    loss, _ = O.probe_loss(tensors, seed=case["seed"] + 300)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items() if v.requires_grad}
    return tensors, loss, grads, ctx.buffers


def test_seg_oracle_training_matches_reference_golden():
    """seg DiscoNet in train() mode: same pin as the detection model (outputs, gradient norms / cosine, BN buffers)."""
    from oracle.make_golden import SEG_KEYS, SEG_STRIDES, SEG_TRAIN_CASE, grad_digest
    name = "seg_train_a2_b1"
    rec = np.load(os.path.join(GOLD, name + ".npz"))
    sd, *_ = __import__("oracle.make_golden", fromlist=["x"]).seg_case_inputs(SEG_TRAIN_CASE, _template(name))
    tensors, loss, grads, bufs = oracle_seg_train_step(SEG_TRAIN_CASE, sd)
    for k in SEG_KEYS:
        _check_sub(name, tensors[k], rec, k, tol=2e-4, stride=SEG_STRIDES[k])
    assert sorted(grads) == rec["grad_names"].tolist()
    sub, table = grad_digest(grads)
    ref_table = rec["grad_table"]
    for i, k in enumerate(sorted(grads)):
        if k.endswith(".bias") and ".conv" in k or k.endswith((".0.bias", ".3.bias")) or (k.endswith(".bias") and "conv1_" in k and "conv1_4" not in k):
            continue                       # conv biases (BatchNorm-shadowed ones are rounding noise on both sides)
        if ref_table[i, 0] > 1e-6:
            assert abs(table[i, 0] - ref_table[i, 0]) <= 3e-2 * ref_table[i, 0], (k, table[i, 0], ref_table[i, 0])
    cos = float(np.dot(sub, rec["grad_sub"]) / (np.linalg.norm(sub) * np.linalg.norm(rec["grad_sub"])))
    assert cos >= 0.999, cos
    bsub, _ = grad_digest({k: v.float() for k, v in bufs.items()}, stride=7)
    assert np.abs(bsub - rec["buf_sub"]).max() <= 5e-4 * np.abs(rec["buf_sub"]).max()


def test_fused_subpix_weight_stream_reproduces_the_conv():
    """plan.pack_conv_subpix_fused: walk the packed slot stream exactly the way conv_tc.cu MODE 4 addresses it (two issuers py, the
    px-merged B operands of 4 * block_n rows, unit order inside the chunk-major groups) and check that the four class accumulators
    add up to conv(cat(nearest_up2(a), b)) (Backbone.py:214-216,233-235)."""
    import torch.nn.functional as F
    from disconet_b200.plan import pack_conv_subpix_fused
    g = torch.Generator().manual_seed(11)
    c0, c1, c_out, H, W = 32, 16, 24, 8, 12                # c_out padded to block_n = 32
    bf = lambda t: t.to(torch.bfloat16).float()            # activations exact in bf16: the A_lo pass contributes nothing
    a = bf(torch.randn(1, c0, H // 2, W // 2, generator=g))
    b = bf(torch.randn(1, c1, H, W, generator=g))
    w = torch.randn(c_out, c0 + c1, 3, 3, generator=g) / 10
    bias = torch.randn(c_out, generator=g)
    plan = pack_conv_subpix_fused(w, bias, src_channels=[c0, c1])
    n = plan.block_n
    assert n == 32 and plan.stacked and plan.fused_subpix and plan.c_blk == 16
    unit = 2 * 2 * n * 8                                   # elements of one unit [chunk 2][part 2][n][8]
    stream = plan.wpack.view(torch.bfloat16).float().view(-1, 9 * unit)          # slots of nine units
    ncb0, ncb1 = c0 // 16, c1 // 16
    assert stream.shape[0] == 2 * ncb0 + ncb1

    def operand(group, units_in_group, u0, rows):
        """B operand that starts at unit u0 of a chunk-major group and spans `rows` 8-channel rows -> [rows, 16 channels]."""
        gr = group.view(2, units_in_group * 2 * n, 8)      # [chunk][(unit, part, n) rows][8 ch]
        return torch.cat((gr[0, u0 * 2 * n: u0 * 2 * n + rows], gr[1, u0 * 2 * n: u0 * 2 * n + rows]), dim=1)

    ap = F.pad(a, (1, 1, 1, 1))                            # low-res window origin (a0 - 1, b0 - 1)
    bp = F.pad(b, (1, 2, 1, 2))                            # full-res window origin (2 a0 - 1, 2 b0 - 1), 34 x 18 per tile
    hl, wl = H // 2, W // 2
    acc = torch.zeros(2, 2, hl, wl, 2 * n)                 # [py][px][a][b][hh | hl columns]
    slot = 0
    for cb in range(ncb0):
        A = ap[0, cb * 16:(cb + 1) * 16]                   # [16, hl + 2, wl + 2]
        for ty in range(2):
            s = stream[slot]; slot += 1
            assert s[8 * unit:].abs().max() == 0           # pad unit
            for py in range(2):
                grp = s[py * 4 * unit:(py + 1) * 4 * unit]
                win = lambda c: A[:, py + ty: py + ty + hl, c: c + wl].permute(1, 2, 0)      # A rows of window column c
                acc[py, 0] += win(0) @ operand(grp, 4, 0, 2 * n).T                            # class px = 0, tx = 0
                acc[py, 1] += win(2) @ operand(grp, 4, 3, 2 * n).T                            # class px = 1, tx = 1
                both = win(1) @ operand(grp, 4, 1, 4 * n).T                                   # px-merged: (px 0, tx 1) | (px 1, tx 0)
                acc[py, 0] += both[..., :2 * n]; acc[py, 1] += both[..., 2 * n:]
    for cb in range(ncb1):
        Bw = bp[0, cb * 16:(cb + 1) * 16]
        s = stream[slot]; slot += 1
        for py in range(2):
            for kh in range(3):
                grp = s[kh * 3 * unit:(kh + 1) * 3 * unit]
                win = lambda c: Bw[:, py + kh: py + kh + 2 * hl: 2, c: c + 2 * wl: 2].permute(1, 2, 0)   # window column c = px + kw
                acc[py, 0] += win(0) @ operand(grp, 3, 2, 2 * n).T                            # px 0, kw 0
                acc[py, 1] += win(3) @ operand(grp, 3, 0, 2 * n).T                            # px 1, kw 2
                m1 = win(1) @ operand(grp, 3, 1, 4 * n).T                                     # [W(kh,1); W(kh,0)]
                m2 = win(2) @ operand(grp, 3, 0, 4 * n).T                                     # [W(kh,2); W(kh,1)]
                acc[py, 0] += m1[..., :2 * n] + m2[..., :2 * n]; acc[py, 1] += m1[..., 2 * n:] + m2[..., 2 * n:]
    out = torch.zeros(H, W, n)
    for py in range(2):
        for px in range(2):
            out[py::2, px::2] = acc[py, px][..., :n] + acc[py, px][..., n:] + plan.bias       # epilogue: hh + hl columns + bias
    ref = F.conv2d(torch.cat((F.interpolate(a, scale_factor=2), b), 1), w, bias, padding=1)[0].permute(1, 2, 0)
    assert (out[..., :c_out] - ref).abs().max() < 2e-4 * ref.abs().max()
    assert out[..., c_out:].abs().max() == 0

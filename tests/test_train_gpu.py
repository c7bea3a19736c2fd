"""GPU parity of the TRAINING path: batch-statistics BatchNorm, weight / data gradients on the
tensor cores, the DiscoGraph fusion block in train mode and the whole training step behind the drop-in class.

Two levels, because train-mode gradients of a deep ReLU/BatchNorm stack are ill-conditioned (see the measured
repeatability of the reference itself in tests/test_oracle_cpu.py::test_oracle_training_matches_reference_golden):
  * per kernel, on IDENTICAL inputs, against torch (fp64 on the CPU): tolerance 1e-3 rel-max or tighter;
  * end to end against the oracle's autograd and the live-reference goldens: outputs <= 1e-3 rel-max, gradients
    by per-tensor norm (5 %) and cosine similarity (>= 0.999) -- the level at which two fp32 runs of the reference
    agree with each other.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import act_value, rel_l2, rel_max, to_act
from oracle import disconet_oracle as O
from oracle.make_golden import TRAIN_CASES, TRAIN_OUT_KEYS, golden_case_inputs, grad_digest
from test_oracle_cpu import GOLD, _Cfg, _template, oracle_train_step

pytestmark = pytest.mark.gpu
P = 1  # PREC_BF16X3


def _lib():
    from disconet_b200 import _lib
    return _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _lo(t):
    return t.stride(0)


# ------------------------------------------------------------------------------------------------------------
def test_bn_train_forward_backward_matches_torch(cuda_dev):
    L = _lib()
    lib = L.load()
    dev = cuda_dev
    rng = np.random.default_rng(0)
    for (n, h, w, c) in [(2, 16, 24, 32), (1, 8, 8, 512), (3, 32, 16, 64)]:
        z = torch.from_numpy(rng.standard_normal((n, c, h, w)).astype(np.float32) * 1.5 + 0.3)
        gam = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
        bet = torch.from_numpy(rng.normal(0, 0.3, c).astype(np.float32))
        rm0 = torch.from_numpy(rng.normal(0, 0.1, c).astype(np.float32))
        rv0 = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
        g_a = torch.from_numpy(rng.standard_normal((n, c + 8, h, w)).astype(np.float32))          # slice [4, 4+c)
        g_b = torch.from_numpy(rng.standard_normal((n, c, 2 * h, 2 * w)).astype(np.float32))      # pooled source
        # torch reference in float64
        zd = z.double().requires_grad_(True)
        gd, bd = gam.double().requires_grad_(True), bet.double().requires_grad_(True)
        rm, rv = rm0.double().clone(), rv0.double().clone()
        y = F.relu(F.batch_norm(zd, rm, rv, gd, bd, training=True, momentum=0.1, eps=1e-5))
        gy = g_a[:, 4:4 + c].double() + F.avg_pool2d(g_b.double(), 2) * 4
        y.backward(gy)
        # ours
        z_d = z.permute(0, 2, 3, 1).contiguous().to(dev)
        out = torch.empty((2, n, h, w, c), dtype=torch.bfloat16, device=dev)
        dz = torch.empty_like(out)
        sums = torch.zeros(1024, dtype=torch.float64, device=dev)
        stats = torch.zeros(2 * c, device=dev)
        gam_d, bet_d, rm_d, rv_d = gam.to(dev), bet.to(dev), rm0.to(dev), rv0.to(dev)
        nbt = torch.zeros((), dtype=torch.int64, device=dev)
        ga_d = g_a.permute(0, 2, 3, 1).contiguous().to(dev)
        gb_d = g_b.permute(0, 2, 3, 1).contiguous().to(dev)
        dgam, dbet = torch.empty(c, device=dev), torch.empty(c, device=dev)
        d = L.BnDesc()
        d.z, d.n, d.h, d.w, d.c = z_d.data_ptr(), n, h, w, c
        d.gamma, d.beta = gam_d.data_ptr(), bet_d.data_ptr()
        d.running_mean, d.running_var, d.num_batches_tracked = rm_d.data_ptr(), rv_d.data_ptr(), nbt.data_ptr()
        d.momentum, d.eps = 0.1, 1e-5
        d.sums, d.stats = sums.data_ptr(), stats.data_ptr()
        d.out_hi, d.out_lo_off, d.relu = out.data_ptr(), _lo(out), 1
        d.n_g = 2
        d.g[0].ptr, d.g[0].c_total, d.g[0].c_off, d.g[0].pool = ga_d.data_ptr(), c + 8, 4, 0
        d.g[1].ptr, d.g[1].c_total, d.g[1].c_off, d.g[1].pool = gb_d.data_ptr(), c, 0, 1
        d.dz_hi, d.dz_lo_off = dz.data_ptr(), _lo(dz)
        d.dgamma, d.dbeta = dgam.data_ptr(), dbet.data_ptr()
        L.check(lib.disco_bn_train_forward(C.byref(d), _stream(dev)), "bn_fwd")
        L.check(lib.disco_bn_train_backward(C.byref(d), _stream(dev)), "bn_bwd")
        torch.cuda.synchronize()
        y_ours = act_value(out).cpu().permute(0, 3, 1, 2)
        dz_ours = act_value(dz).cpu().permute(0, 3, 1, 2)
        assert rel_max(y_ours, y.detach()) < 3e-5
        assert rel_max(rm_d.cpu(), rm) < 1e-5 and rel_max(rv_d.cpu(), rv) < 1e-5 and int(nbt) == 1
        assert rel_max(dz_ours, zd.grad) < 1e-4, rel_max(dz_ours, zd.grad)
        assert rel_max(dgam.cpu(), gd.grad) < 1e-4 and rel_max(dbet.cpu(), bd.grad) < 1e-4


def test_device_weight_pack_is_bit_identical_to_host_pack(cuda_dev):
    """disco_pack_weights (device) == plan.pack_conv (host tensor ops) for forward and data-gradient images."""
    from disconet_b200 import _lib as L
    from disconet_b200.plan import pack_conv
    from disconet_b200.train import TrainRunner
    dev = cuda_dev
    lib = L.load()
    rng = np.random.default_rng(3)
    for (co, ci, k, srcs, stride) in [(32, 13, 3, [16], 1), (64, 32, 3, [32], 2), (256, 768, 3, [512, 256], 1), (48, 64, 1, [64], 1),
                                      (512, 512, 3, [512], 1), (64, 64, 1, [64], 1)]:
        w = torch.from_numpy(rng.standard_normal((co, ci, k, k)).astype(np.float32)).to(dev)
        b = torch.from_numpy(rng.standard_normal(co).astype(np.float32)).to(dev)
        wp = torch.zeros(co, sum(srcs), k, k, device=dev)
        wp[:, :ci] = w
        ref = pack_conv(wp, b, src_channels=srcs, stride=stride, relu=False, precision=P)
        got = pack_conv(torch.zeros_like(wp), torch.zeros_like(b), src_channels=srcs, stride=stride, relu=False, precision=P)
        L.check(lib.disco_pack_weights(C.byref(TrainRunner._pack_desc(got, w, b)), _stream(dev)), "pack")
        torch.cuda.synchronize()
        assert torch.equal(got.wpack, ref.wpack) and torch.equal(got.bias, ref.bias), (co, ci, k)
        if ci == sum(srcs):   # data-gradient image of each source slice
            c0 = 0
            for cs in srcs:
                wt = TrainRunner._dgrad_weight(w, c0, cs)
                ref = pack_conv(wt, torch.zeros(cs, device=dev), src_channels=[co], relu=False, precision=P)
                got = pack_conv(torch.zeros_like(wt), torch.zeros(cs, device=dev), src_channels=[co], relu=False, precision=P)
                L.check(lib.disco_pack_weights(C.byref(TrainRunner._pack_desc(got, w, None, transpose=True, c0=c0, n_real=cs)), _stream(dev)), "pack")
                torch.cuda.synchronize()
                assert torch.equal(got.wpack, ref.wpack), ("dgrad", co, ci, k, c0)
                c0 += cs


# ------------------------------------------------------------------------------------------------------------
WGRAD_CASES = [
    # (n, h_in, w_in, src channels, ups, c_out, stride, taps, c_in_real)
    (2, 32, 32, [32], [0], 32, 1, 9, 32),
    (1, 32, 48, [16], [0], 32, 1, 9, 13),
    (2, 32, 32, [64, 32], [1, 0], 32, 1, 9, 96),
    (1, 16, 16, [512, 256], [1, 0], 256, 1, 9, 768),
    (2, 64, 64, [32], [0], 64, 2, 9, 32),
    (1, 32, 32, [256], [0], 512, 2, 9, 256),
    (2, 32, 32, [64], [0], 64, 1, 1, 64),
    (1, 64, 64, [64], [0], 48, 1, 1, 64),
    (2, 32, 32, [256], [0], 256, 1, 1, 256),
    (1, 20, 28, [48], [0], 16, 1, 9, 48),     # ragged tile edges
]


def _wgrad_inputs(case, dev, seed):
    n, hi, wi, cs, ups, co, stride, taps, cir = case
    rng = np.random.default_rng(seed)
    ho, wo = (hi - 1) // stride + 1, (wi - 1) // stride + 1
    srcs = [torch.from_numpy(rng.standard_normal((n, c, hi >> u, wi >> u)).astype(np.float32)) for c, u in zip(cs, ups)]
    if cir < sum(cs):
        srcs[0][:, cir:] = 0
    dz = torch.from_numpy(rng.standard_normal((n, co, ho, wo)).astype(np.float32))
    return srcs, dz, (ho, wo)


@pytest.mark.parametrize("ci", range(len(WGRAD_CASES)))
def test_wgrad_tensor_core_vs_validator_and_torch(ci, cuda_dev):
    L = _lib()
    lib = L.load()
    dev = cuda_dev
    case = WGRAD_CASES[ci]
    n, hi, wi, cs, ups, co, stride, taps, cir = case
    srcs, dz, (ho, wo) = _wgrad_inputs(case, dev, 100 + ci)
    acts = [to_act(s, P).to(dev) for s in srcs]
    dz_act = to_act(dz, P).to(dev)
    wg = L.WgradDesc()
    for i, a in enumerate(acts):
        wg.src[i], wg.src_lo_off[i], wg.src_c[i], wg.src_up[i] = a.data_ptr(), _lo(a), cs[i], ups[i]
    wg.n, wg.h_in, wg.w_in, wg.h_out, wg.w_out, wg.stride, wg.taps = n, hi, wi, ho, wo, stride, taps
    wg.dz_hi, wg.dz_lo_off, wg.c_out = dz_act.data_ptr(), _lo(dz_act), co
    wg.c_in_real, wg.passes = cir, 3
    dw = torch.zeros((co, cir, taps), device=dev)
    dw_ref = torch.zeros((co, cir, taps), device=dev)
    wg.dw = dw_ref.data_ptr()
    L.check(lib.disco_conv_wgrad_reference(C.byref(wg), _stream(dev)), "wgrad_ref")
    wg.dw = dw.data_ptr()
    splits = lib.disco_conv_wgrad_splits(C.byref(wg))
    L.check(splits, "splits")
    partial = torch.empty(splits * co * taps * sum(cs), device=dev)
    wg.partial, wg.splits = partial.data_ptr(), splits
    L.check(lib.disco_conv_wgrad(C.byref(wg), _stream(dev)), "wgrad")
    torch.cuda.synchronize()
    # torch (fp64, CPU): weight gradient of the equivalent conv on the (upsampled, concatenated) input
    x = torch.cat([F.interpolate(s, scale_factor=2) if u else s for s, u in zip(srcs, ups)], 1).double()
    k = 3 if taps == 9 else 1
    w0 = torch.zeros(co, sum(cs), k, k, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w0, stride=stride, padding=k // 2).backward(dz.double())
    want = w0.grad[:, :cir].reshape(co, cir, taps)
    e_ref = rel_max(dw_ref.cpu(), want)
    e_tc = rel_max(dw.cpu(), want)
    print(f"wgrad case {ci}: validator {e_ref:.2e} tensor-core {e_tc:.2e} splits {splits}")
    assert e_ref < 2e-4, e_ref
    assert e_tc < 2e-4, e_tc


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 32, 32, 32, 64), (1, 16, 48, 128, 256), (1, 32, 32, 256, 512)])
def test_dgrad_stride2_zero_stuffed_matches_autograd(shape, cuda_dev):
    """Data gradient of a stride-2 3x3 conv = stride-1 conv of the zero-stuffed gradient with flipped, transposed
    weights (conv kernel `src_up = 2`)."""
    from disconet_b200 import ops
    from disconet_b200.plan import pack_conv
    from disconet_b200.train import TrainRunner
    dev = cuda_dev
    n, hi, wi, cin, cout = shape
    rng = np.random.default_rng(7)
    w = torch.from_numpy((rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32))
    dz = torch.from_numpy(rng.standard_normal((n, cout, hi // 2, wi // 2)).astype(np.float32))
    x = torch.zeros(n, cin, hi, wi, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double(), stride=2, padding=1).backward(dz.double())
    wt = TrainRunner._dgrad_weight(w.to(dev), 0, cin)
    plan = pack_conv(wt, torch.zeros(cin, device=dev), src_channels=[cout], stride=1, relu=False, precision=P, keep_ref=True)
    dz_act = to_act(dz, P).to(dev)
    for reference in (True, False):
        out = torch.zeros((n, hi, wi, cin), device=dev)
        ops.ConvCall(plan, [dz_act], [2], (out,), n=n, h_in=hi, w_in=wi).launch(_stream(dev), reference=reference)
        torch.cuda.synchronize()
        e = rel_max(out.cpu().permute(0, 3, 1, 2), x.grad)
        print("dgrad s2", shape, "validator" if reference else "tensor-core", f"{e:.2e}")
        assert e < 1e-4, e


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 32, 32), (1, 32, 48, 128, 64, 64), (1, 16, 16, 512, 256, 256)])
def test_dgrad_stride1_and_upsample_concat_matches_autograd(shape, cuda_dev):
    """Data gradient of conv(cat(nearest_up2(a), b)) (Backbone.py:176-178 ...): one stride-1 conv of the output gradient per
    source with that source's flipped, transposed weight slice (`TrainRunner._dgrad_weight`); the gradient of the upsampled
    source is the 2x2 sum-pool of its full-resolution slice (what bn_bwd's `pool` gradient sources do).  Against fp64 autograd,
    through the TMA-fed tensor-core kernel and the CUDA-core validator."""
    from disconet_b200 import ops
    from disconet_b200.plan import pack_conv
    from disconet_b200.train import TrainRunner
    dev = cuda_dev
    n, h, w_, c_up, c_skip, cout = shape
    rng = np.random.default_rng(11)
    cin = c_up + c_skip
    wgt = torch.from_numpy((rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32))
    dz = torch.from_numpy(rng.standard_normal((n, cout, h, w_)).astype(np.float32))
    a = torch.zeros(n, c_up, h // 2, w_ // 2, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(n, c_skip, h, w_, dtype=torch.float64, requires_grad=True)
    F.conv2d(torch.cat((F.interpolate(a, scale_factor=2), b), 1), wgt.double(), padding=1).backward(dz.double())
    dz_act = to_act(dz, P).to(dev)
    for (c0, cs, want, pool) in ((0, c_up, a.grad, True), (c_up, c_skip, b.grad, False)):
        wt = TrainRunner._dgrad_weight(wgt.to(dev), c0, cs)
        plan = pack_conv(wt, torch.zeros(cs, device=dev), src_channels=[cout], stride=1, relu=False, precision=P, keep_ref=True)
        for reference in (True, False):
            out = torch.zeros((n, h, w_, cs), device=dev)
            ops.ConvCall(plan, [dz_act], [0], (out,), n=n, h_in=h, w_in=w_).launch(_stream(dev), reference=reference)
            torch.cuda.synchronize()
            got = out.cpu().permute(0, 3, 1, 2).double()
            if pool:
                got = F.avg_pool2d(got, 2) * 4
            e = rel_max(got, want)
            print("dgrad s1", shape, "upsampled source" if pool else "skip source", "validator" if reference else "tensor-core", f"{e:.2e}")
            assert e < 1e-4, e


# ------------------------------------------------------------------------------------------------------------
def _disco(case, sd, dev, train=True):
    from disconet_b200 import DiscoNet
    m = DiscoNet(_Cfg(), layer=case.get("layer", 3), kd_flag=case["kd_flag"], num_agent=case["A"],
                 compress_level=case["compress_level"], only_v2i=case["only_v2i"])
    m.load_state_dict(sd)
    m = m.to(dev)
    return m.train() if train else m.eval()


@pytest.mark.parametrize("name", ["train_a2_b1", "train_a3_b1_absent"])
def test_fusion_block_train_forward_backward_matches_oracle(name, cuda_dev):
    """The DiscoGraph block alone (PWF with per-pair batch statistics, softmax, weighted sum, warp) fed with the
    SAME collaboration-layer features as the oracle: output, running statistics, gradients wrt the features and
    the PWF parameters."""
    from disconet_b200.train import TrainRunner
    dev = cuda_dev
    case = TRAIN_CASES[name]
    A, B = case["A"], case["B"]
    sd, bev, T, na = golden_case_inputs(case, _template(name))
    rng = np.random.default_rng(case["seed"] + 500)
    x3 = torch.from_numpy(np.maximum(rng.standard_normal((A * B, 256, 32, 32)), 0).astype(np.float32))
    for b, nn_ in enumerate(case["num_agent"]):
        for a in range(nn_, A):
            x3[a * B + b] = 0
    x3_act = to_act(x3, P)
    x3q = act_value(x3_act).permute(0, 3, 1, 2).contiguous()      # exactly what the kernels see
    cot = torch.from_numpy(rng.standard_normal((A * B, 256, 32, 32)).astype(np.float32))
    # ---- oracle ----
    sdo = {k: (v.clone().double().requires_grad_(True) if k.startswith("pixel_weighted_fusion.") and v.is_floating_point()
               and "running" not in k else (v.clone().double() if v.is_floating_point() else v.clone())) for k, v in sd.items()}
    x3o = x3q.double().requires_grad_(True)
    with O.training(sdo) as ctx:
        fused, _ = O.fuse(sdo, x3o, T, na, B, A, case["only_v2i"])
    (fused * cot.double()).sum().backward()
    # ---- ours ----
    m = _disco(case, sd, dev)
    runner = TrainRunner(m._getter(), A * B, 256, 256, dev, "u_encoder.", "decoder.", heads=True,
                         pwf_prefix="pixel_weighted_fusion.", batch_size=B, agents=A, only_v2i=case["only_v2i"],
                         kd_keys=["x8", "x7", "x6", "x5", "x3f"])
    st = _stream(dev)
    runner.run_ops(runner.ops_pre, st)                       # pack the current weights
    runner.act["x3"].copy_(x3_act.to(dev))
    runner.trans.copy_(T)
    runner.na.copy_(na[:, 0])
    runner.run_ops(runner.ops_fusion_fwd, st)
    torch.cuda.synchronize()
    got = act_value(runner.act["x3f"]).cpu().permute(0, 3, 1, 2)
    e = rel_max(got, fused.detach())
    print(name, "fused fwd rel-max", e)
    assert e < 2e-4, e
    for k in ("bn1_1", "bn1_2", "bn1_3"):
        for f in ("running_mean", "running_var"):
            key = f"pixel_weighted_fusion.{k}.{f}"
            assert rel_max(m.state_dict()[key].cpu(), ctx.buffers[key]) < 1e-4, key
        assert int(m.state_dict()[f"pixel_weighted_fusion.{k}.num_batches_tracked"]) == int(ctx.buffers[f"pixel_weighted_fusion.{k}.num_batches_tracked"])
    # backward: the fused map's gradient sources are the external (KD) buffer and conv5_1's data gradient
    runner.ext["x3f"].copy_(cot.permute(0, 2, 3, 1).contiguous())
    runner.st["c5_1"].gbufs[1].zero_()
    runner.run_ops(runner.ops_bwd_fusion, st)
    torch.cuda.synchronize()
    out = {k: v for k, v in runner._collect(runner.G.clone()).items() if k.startswith("pixel_weighted_fusion.")}
    dx3 = (runner.en_gbuf + runner.dfeat).cpu().permute(0, 3, 1, 2)
    e = rel_max(dx3, x3o.grad)
    print(name, "d x3 rel-max", e, "rel-l2", rel_l2(dx3, x3o.grad))
    # A ReLU gate of the PWF tail that sits within rounding distance of zero can fall on the other side in this fp32
    # path than in the fp64 oracle; the whole gradient of THAT pixel then differs (measured: 99.9 % of a row's squared
    # error in one pixel).  So: everything but the 4 worst pixels of a row must agree to 2e-4 rel-l2, and all of it
    # to 1e-2.
    for r in range(A * B):
        ref_r = x3o.grad[r]
        if ref_r.abs().max() == 0:
            assert dx3[r].abs().max() == 0
            continue
        per_px = ((dx3[r].double() - ref_r) ** 2).sum(0).flatten().sort(descending=True).values
        assert (per_px[4:].sum().sqrt() / ref_r.norm()).item() < 2e-4, r
        assert (per_px.sum().sqrt() / ref_r.norm()).item() < 1e-2, r
    for k, g in out.items():
        ref = sdo[k].grad
        if k.endswith("bias") and ("conv1_1" in k or "conv1_2" in k or "conv1_3" in k):
            assert g.abs().max() == 0            # BN-shadowed conv bias: exactly zero here, rounding noise in torch
            continue
        e, e2 = rel_max(g.cpu(), ref), rel_l2(g.cpu(), ref)
        print(name, k, "rel-max", e, "rel-l2", e2)
        assert e < 5e-2 and e2 < 1e-2, (k, e, e2)   # (without a gate flip: 1e-5 .. 5e-4, see the a2_b1 case)


# ------------------------------------------------------------------------------------------------------------
def _ours_train_step(case, sd, dev, seed):
    m = _disco(case, sd, dev)
    _, bev, T, na = golden_case_inputs(case, sd)
    out = m(bev.to(dev), T, na, batch_size=case["B"])
    tensors = dict(zip(TRAIN_OUT_KEYS, (out[0]["cls"], out[0]["loc"]) + tuple(out[1:])))
    loss, _ = O.probe_loss(tensors, seed=seed)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in m.named_parameters()}
    bufs = {k: v.detach().float().cpu() for k, v in m.named_buffers()}
    return {k: v.detach().cpu() for k, v in tensors.items()}, loss.item(), grads, bufs


@pytest.mark.parametrize("name", ["train_a2_b1", "train_a3_b1_absent"])
def test_disconet_training_step_matches_oracle_and_golden(name, cuda_dev):
    case = TRAIN_CASES[name]
    rec = np.load(os.path.join(GOLD, name + ".npz"))
    sd, *_ = golden_case_inputs(case, _template(name))
    ref_t, ref_loss, ref_g, ref_bufs, _ = oracle_train_step(case, sd)
    got_t, loss, grads, bufs = _ours_train_step(case, sd, cuda_dev, case["seed"] + 300)
    for k in TRAIN_OUT_KEYS:
        e = rel_max(got_t[k], ref_t[k].detach())
        print(f"{name} fwd {k}: rel-max {e:.2e}")
        assert e <= 1e-3, (k, e)
    assert abs(loss - ref_loss.item()) <= 2e-3 * abs(ref_loss.item())
    # gradients: the same set of parameters receives one as in the reference (dead parameters get None)
    none_ref = set(rec["grad_none"].tolist())
    assert {k for k, g in grads.items() if g is None} == none_ref
    a, b = [], []
    for k, g in ref_g.items():
        if k in none_ref:
            continue
        ours = grads[k]
        shadowed = k.endswith(".bias") and not (".bn" in k or "bn_" in k or "box_prediction.1" in k or "conv2." in k
                                               or "box_prediction.3" in k or "conv1_4" in k)
        if shadowed:
            assert ours.abs().max() == 0, k
            continue
        n_ref, n_ours = g.norm().item(), ours.norm().item()
        cos = (ours.double().flatten() @ g.double().flatten()).item() / max(n_ref * n_ours, 1e-30)
        print(f"{name} grad {k}: norm ratio {n_ours / n_ref:.4f} cos {cos:.5f} rel-max {rel_max(ours, g):.2e}")
        assert abs(n_ours - n_ref) <= 5e-2 * n_ref, (k, n_ours, n_ref)
        assert cos >= 0.995, (k, cos)
        a.append(ours.flatten()); b.append(g.flatten())
    a, b = torch.cat(a).double(), torch.cat(b).double()
    cos_all = (a @ b / (a.norm() * b.norm())).item()
    print(f"{name} all gradients: cos {cos_all:.6f} rel-l2 {((a - b).norm() / b.norm()).item():.3e}")
    assert cos_all >= 0.999
    # live-reference golden digest
    sub, table = grad_digest({k: (g if g is not None else torch.zeros_like(ref_g.get(k, torch.zeros(1)))) for k, g in grads.items()
                              if k in ref_g})
    cos_g = float(np.dot(sub, rec["grad_sub"]) / (np.linalg.norm(sub) * np.linalg.norm(rec["grad_sub"])))
    assert cos_g >= 0.999, cos_g
    # BatchNorm buffers after the step
    for k, v in ref_bufs.items():
        if k.startswith(("u_encoder.bn5", "u_encoder.bn6", "u_encoder.bn7", "u_encoder.bn8")):
            continue
        e = rel_max(bufs[k], v.float())
        assert e <= 1e-3, (k, e)


def test_fafnet_training_step_matches_oracle(cuda_dev):
    from disconet_b200 import FaFNet
    dev = cuda_dev
    sd = O.synth_state_dict(_template("fafnet_a2_128"), seed=41)
    bev = O.synth_bev(2, H=128, W=128, seed=141)
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    with O.training(sdo):
        ref = O.fafnet_forward_graph(sdo, bev)
    keys = ("cls", "loc", "x_7", "x_5")
    lref, _ = O.probe_loss({k: ref[k] for k in keys}, seed=5)
    lref.backward()
    m = FaFNet(_Cfg(), kd_flag=1, num_agent=2)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    res, x8, x7, x6, x5, x3 = m(bev.to(dev))
    got = {"cls": res["cls"], "loc": res["loc"], "x_7": x7, "x_5": x5}
    for k in keys:
        assert rel_max(got[k].detach().cpu(), ref[k].detach()) <= 1e-3, k
    loss, _ = O.probe_loss(got, seed=5)
    loss.backward()
    torch.cuda.synchronize()
    a, b = [], []
    for k, p in m.named_parameters():
        g = sdo[k].grad
        if g is None:
            assert p.grad is None, k
            continue
        if p.grad.abs().max() == 0:
            continue
        a.append(p.grad.detach().cpu().flatten()); b.append(g.flatten())
    a, b = torch.cat(a).double(), torch.cat(b).double()
    cos = (a @ b / (a.norm() * b.norm())).item()
    print("fafnet train grads cos", cos)
    assert cos >= 0.999


def test_kd_loss_kernel_matches_reference_formula(cuda_dev):
    """f2: one fused launch == the reference's permute/log_softmax/softmax/KLDivLoss(mean) chain (CoDetModule.py:334-382)."""
    from disconet_b200.kd import kd_kl_mean
    dev = cuda_dev
    rng = np.random.default_rng(9)
    for (n, c, h, w) in [(2, 64, 32, 32), (1, 256, 16, 24), (3, 128, 8, 8)]:
        s = torch.from_numpy((rng.standard_normal((n, c, h, w)) * 2).astype(np.float32))
        t = torch.from_numpy((rng.standard_normal((n, c, h, w)) * 3).astype(np.float32))
        sd = s.double().requires_grad_(True)
        ref = torch.nn.KLDivLoss(reduction="mean")(F.log_softmax(sd.permute(0, 2, 3, 1).reshape(-1, c), dim=1),
                                                   F.softmax(t.double().permute(0, 2, 3, 1).reshape(-1, c), dim=1))
        (ref * 7.0).backward()
        sg = s.to(dev).requires_grad_(True)
        got = kd_kl_mean(sg, t.to(dev))
        (got * 7.0).backward()
        torch.cuda.synchronize()
        assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()), (got.item(), ref.item())
        assert rel_max(sg.grad.cpu(), sd.grad) < 1e-5


def test_focal_loss_kernel_matches_reference_formula(cuda_dev):
    """f4: fused focal loss forward / backward == the reference chain (loss.py:213-219,322-394), incl. sum()/N."""
    from disconet_b200.loss import SoftmaxFocalClassificationLoss
    dev = cuda_dev
    rng = np.random.default_rng(13)
    gamma, alpha = 2.0, 0.25
    for (n, m, k) in [(3, 4096, 2), (2, 1000, 5)]:
        z = torch.from_numpy((rng.standard_normal((n, m, k)) * 2).astype(np.float32))
        lab = rng.integers(0, k, (n, m))
        lab[rng.random((n, m)) < 0.9] = 0
        t = torch.nn.functional.one_hot(torch.from_numpy(lab), k).float()
        zd = z.double().requires_grad_(True)
        td = t.double()
        ce = F.cross_entropy(zd.permute(0, 2, 1), td.max(dim=-1)[1], reduction="none").unsqueeze(-1) * td
        p = F.softmax(zd, dim=-1)
        pt = td * p + (1 - td) * (1 - p)
        aw = torch.where(td[..., 0] == 1, torch.tensor(1 - alpha, dtype=torch.float64), torch.tensor(alpha, dtype=torch.float64)).unsqueeze(-1)
        ref = torch.pow(1.0 - pt, gamma) * aw * ce
        (ref.sum() / n).backward()
        zg = z.to(dev).requires_grad_(True)
        got = SoftmaxFocalClassificationLoss(gamma, alpha)(zg, t.to(dev))
        (got.sum() / n).backward()
        torch.cuda.synchronize()
        assert got.shape == ref.shape
        assert rel_max(got.detach().cpu(), ref.detach()) < 1e-5
        assert rel_max(zg.grad.cpu(), zd.grad) < 1e-5
        # dense (non-broadcast) upstream gradient
        w = torch.from_numpy(rng.random((n, m, k)).astype(np.float32))
        zd.grad = None
        (torch.pow(1.0 - (td * F.softmax(zd, -1) + (1 - td) * (1 - F.softmax(zd, -1))), gamma) * aw *
         (F.cross_entropy(zd.permute(0, 2, 1), td.max(dim=-1)[1], reduction="none").unsqueeze(-1) * td) * w.double()).sum().backward()
        zg.grad = None
        (SoftmaxFocalClassificationLoss(gamma, alpha)(zg, t.to(dev)) * w.to(dev)).sum().backward()
        assert rel_max(zg.grad.cpu(), zd.grad) < 1e-5


# ---- more of the path's flags in train() mode, against the oracle's autograd (no extra golden: the oracle itself is pinned
# ---- to the live reference for eval with these flags and for train() on the two cases above) ----
EXTRA_TRAIN_CASES = {
    # two scenes with different agent counts, infrastructure-only links, one ego in communication outage
    "b2_v2i_outage": dict(A=3, B=2, num_agent=[3, 2], kd_flag=1, only_v2i=True, compress_level=0, seed=61, layer=3,
                          outage=[[0, 1, 0], [0, 0, 0]]),
    # collaboration on the 128-channel 64x64 level (--layer 2)
    "layer2": dict(A=2, B=1, num_agent=[2], kd_flag=1, only_v2i=False, compress_level=0, seed=62, layer=2, outage=None),
    # the communication bottleneck 256 -> 64 -> 256 on x_3 (Backbone.py:139-141)
    "compress2": dict(A=2, B=1, num_agent=[2], kd_flag=1, only_v2i=False, compress_level=2, seed=63, layer=3, outage=None),
}


@pytest.mark.parametrize("name", list(EXTRA_TRAIN_CASES))
def test_training_step_flags_match_oracle(name, cuda_dev):
    from disconet_b200 import DiscoNet
    case = EXTRA_TRAIN_CASES[name]
    A, B = case["A"], case["B"]
    m = DiscoNet(_Cfg(), layer=case["layer"], kd_flag=1, num_agent=A, only_v2i=case["only_v2i"], compress_level=case["compress_level"])
    sd, bev, T, na = golden_case_inputs(case, m.state_dict())
    m.load_state_dict(sd)
    # ---- oracle ----
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    with O.training(sdo) as ctx:
        ref = O.disconet_forward_graph(sdo, bev, T, na, B, agent_num=A, layer=case["layer"], only_v2i=case["only_v2i"],
                                       return_all=True, outage=case["outage"])
    ref_t = {k: ref[k] for k in TRAIN_OUT_KEYS}
    lref, _ = O.probe_loss(ref_t, seed=case["seed"] + 300)
    lref.backward()
    # ---- ours ----
    m = m.to(cuda_dev).train()
    if case["outage"] is not None:
        m.p_com_outage = 0.5
        draws = iter([bool(case["outage"][b][i]) for b in range(B) for i in range(case["num_agent"][b])])
        m.outage = lambda: next(draws)            # the reference draws once per (scene, present ego) in this order
    out = m(bev.to(cuda_dev), T, na, batch_size=B)
    got_t = dict(zip(TRAIN_OUT_KEYS, (out[0]["cls"], out[0]["loc"]) + tuple(out[1:])))
    loss, _ = O.probe_loss(got_t, seed=case["seed"] + 300)
    loss.backward()
    torch.cuda.synchronize()
    for k in TRAIN_OUT_KEYS:
        e = rel_max(got_t[k].detach().cpu(), ref_t[k].detach())
        print(f"{name} fwd {k}: rel-max {e:.2e}")
        assert e <= 1e-3, (k, e)
    a, b = [], []
    for k, p in m.named_parameters():
        g = sdo[k].grad
        if g is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        if p.grad.abs().max() == 0:          # BN-shadowed conv bias
            continue
        a.append(p.grad.detach().cpu().flatten()); b.append(g.flatten())
    a, b = torch.cat(a).double(), torch.cat(b).double()
    cos = (a @ b / (a.norm() * b.norm())).item()
    print(f"{name} all gradients: cos {cos:.6f} rel-l2 {((a - b).norm() / b.norm()).item():.3e}")
    assert cos >= 0.999
    for k, v in ctx.buffers.items():
        if k.startswith(("u_encoder.bn5", "u_encoder.bn6", "u_encoder.bn7", "u_encoder.bn8", "decoder.bn_pre", "decoder.bn1", "decoder.bn2",
                         "decoder.bn3", "decoder.bn4", "decoder.conv3d")) or "num_batches" in k:
            continue
        assert rel_max(m.state_dict()[k].float().cpu(), v.float()) <= 1e-3, k


def test_backward_of_a_stale_forward_raises(cuda_dev):
    """One set of saved activations per shape: a second training forward invalidates the first one's backward, loudly."""
    from disconet_b200 import FaFNet
    m = FaFNet(_Cfg(), kd_flag=0, num_agent=2)
    m.load_state_dict(O.synth_state_dict(m.state_dict(), seed=71))
    m = m.to(cuda_dev).train()
    bev = O.synth_bev(2, H=64, W=64, seed=72).to(cuda_dev)
    r1 = m(bev)
    r2 = m(bev)
    with pytest.raises(RuntimeError, match="stale forward"):
        r1["cls"].sum().backward()
    r2["cls"].sum().backward()
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in m.parameters())

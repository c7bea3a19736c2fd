"""GPU parity of the tcgen05 conv kernel (through the C-ABI) against torch fp32 conv2d and the
CUDA-core validator, over every layer shape class of the DiscoNet path."""
import pytest
import torch
import torch.nn.functional as F

from disconet_b200._lib import PREC_BF16X3, PREC_FP16
from disconet_b200.ops import alloc_act, conv_forward
from disconet_b200.plan import pack_conv
from helpers import act_value, rel_max, to_act

pytestmark = pytest.mark.gpu

# (name, n, h, w, sources [(c_real, c_pad, up)], c_out, k, stride, out)   out: 'act' | 'f32' | (split,)
CASES = [
    ("c32_s1", 2, 32, 32, [(32, 32, 0)], 32, 3, 1, "act"),
    ("pre1_13to32", 2, 32, 40, [(13, 16, 0)], 32, 3, 1, "act"),
    ("c32to64_s2", 2, 64, 64, [(32, 32, 0)], 64, 3, 2, "act"),
    ("upcat_64up_32", 2, 32, 32, [(64, 64, 1), (32, 32, 0)], 32, 3, 1, "act"),
    ("upcat_512up_256", 1, 32, 32, [(512, 512, 1), (256, 256, 0)], 256, 3, 1, "act"),
    ("c256_s1_multi_stage", 3, 32, 32, [(256, 256, 0)], 256, 3, 1, "act"),
    ("c256to512_s2_two_ntiles", 2, 32, 32, [(256, 256, 0)], 512, 3, 2, "act"),
    ("c512_16x16", 2, 16, 16, [(512, 512, 0)], 512, 3, 1, "act"),
    ("partial_tiles_24x20", 1, 24, 20, [(64, 64, 0)], 64, 3, 1, "act"),
    ("tiny_8x8_s2", 1, 8, 8, [(64, 64, 0)], 128, 3, 2, "act"),
    ("persistent_many_items_stationary", 8, 64, 64, [(32, 32, 0)], 32, 3, 1, "act"),
    ("persistent_many_items_streamed_msub2", 12, 64, 64, [(128, 128, 0)], 128, 3, 1, "act"),
    ("persistent_s2_many_items", 6, 128, 128, [(32, 32, 0)], 64, 3, 2, "act"),
    ("persistent_upcat_many_items", 6, 64, 64, [(256, 256, 1), (128, 128, 0)], 128, 3, 1, "act"),
    ("pw_many_items_msub2", 4, 128, 128, [(128, 128, 0)], 128, 1, 1, "act"),
    ("pw_64to64", 2, 32, 32, [(64, 64, 0)], 64, 1, 1, "act"),
    ("pw_256to256_f32", 2, 32, 32, [(256, 256, 0)], 256, 1, 1, "f32"),
    ("pw_heads_64to48_split", 1, 32, 24, [(64, 64, 0)], 48, 1, 1, (12,)),
    ("pw_ragged_pixels", 1, 10, 13, [(32, 32, 0)], 32, 1, 1, "act"),
]


def _run_case(case, precision, dev, seed=0):
    name, n, h, w, sources, c_out, k, stride, outk = case
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs, acts, ups = [], [], []
    for (c_real, c_pad, up) in sources:
        hs, ws = (h // 2, w // 2) if up else (h, w)
        x = torch.randn(n, c_real, hs, ws, generator=g).to(dev)
        a = to_act(x, precision, c_pad)
        acts.append(a)
        ups.append(up)
        v = act_value(a)[..., :c_real].permute(0, 3, 1, 2)  # value the kernel actually sees
        xs.append(F.interpolate(v, scale_factor=2) if up else v)
    x_cat = torch.cat(xs, 1)
    c_in_real = x_cat.shape[1]
    wgt = (torch.randn(c_out, c_in_real, k, k, generator=g) / (c_in_real * k * k) ** 0.5).to(dev)
    bias = torch.randn(c_out, generator=g).to(dev) * 0.1
    relu = outk == "act"
    # weights as the kernel sees them (rounded to the operand precision) so the check isolates the kernel
    if precision == PREC_BF16X3:
        hi = wgt.to(torch.bfloat16)
        w_seen = hi.float() + (wgt - hi.float()).to(torch.bfloat16).float()
    else:
        w_seen = wgt.to(torch.float16).float()
    # real channels sit at the start of each padded source: scatter weights accordingly
    w_pad = torch.zeros(c_out, sum(s[1] for s in sources), k, k, device=dev)
    o_r = o_p = 0
    for (c_real, c_pad, up) in sources:
        w_pad[:, o_p:o_p + c_real] = wgt[:, o_r:o_r + c_real]
        o_r += c_real
        o_p += c_pad
    plan = pack_conv(w_pad, bias, src_channels=[s[1] for s in sources], stride=stride, relu=relu,
                     precision=precision, keep_ref=True, name=name)
    ref = F.conv2d(x_cat, w_seen, bias, stride=stride, padding=k // 2)
    if relu:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    ho, wo = ref.shape[1], ref.shape[2]
    results = {}
    for which in ("tc", "ref"):
        if outk == "act":
            out = alloc_act(n, ho, wo, c_out, precision, dev)
            out.fill_(float("nan"))
            conv_forward(plan, acts, ups, out, n=n, h_in=h, w_in=w, reference=(which == "ref"))
            got = act_value(out)
        elif outk == "f32":
            out = torch.full((n, ho, wo, c_out), float("nan"), device=dev)
            conv_forward(plan, acts, ups, (out,), n=n, h_in=h, w_in=w, reference=(which == "ref"))
            got = out
        else:
            sp = outk[0]
            o0 = torch.full((n, ho, wo, sp), float("nan"), device=dev)
            o1 = torch.full((n, ho, wo, c_out - sp), float("nan"), device=dev)
            conv_forward(plan, acts, ups, (o0, o1), n=n, h_in=h, w_in=w, out_split=sp, reference=(which == "ref"))
            got = torch.cat((o0, o1), -1)
        torch.cuda.synchronize()
        results[which] = got
    return results, ref


@pytest.mark.parametrize("precision", [PREC_FP16, PREC_BF16X3], ids=["fp16", "bf16x3"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_matches_torch(case, precision, cuda_dev):
    results, ref = _run_case(case, precision, cuda_dev)
    # output rounding: fp16 act 2^-11, bf16 hi+lo ~2^-17; fp32 accumulate-order noise ~1e-6;
    # bf16x3 drops the lo*lo term (~2^-16 relative per product)
    tol = {PREC_FP16: 1e-3, PREC_BF16X3: 5e-5}[precision]
    for which, got in results.items():
        assert torch.isfinite(got).all(), f"{which}: non-finite / unwritten outputs"
        err = rel_max(got, ref)
        assert err < tol, f"{case[0]} {which}: rel-max error {err:.3e} >= {tol}"


def test_conv_rejects_bad_descriptor(cuda_dev):
    from disconet_b200._lib import DiscoError
    w = torch.randn(32, 24, 3, 3, device=cuda_dev)
    with pytest.raises(ValueError):
        pack_conv(w, torch.zeros(32, device=cuda_dev), src_channels=[24])  # not a multiple of 16
    plan = pack_conv(torch.randn(32, 32, 3, 3, device=cuda_dev), torch.zeros(32, device=cuda_dev), src_channels=[32])
    a = alloc_act(1, 16, 16, 32, PREC_BF16X3, cuda_dev)
    out = alloc_act(1, 16, 16, 32, PREC_BF16X3, cuda_dev)
    with pytest.raises(DiscoError):
        conv_forward(plan, [a], [1], out, n=1, h_in=15, w_in=16)  # odd size with upsample flag
    with pytest.raises(ValueError):
        conv_forward(plan, [a.cpu()], [0], out, n=1, h_in=16, w_in=16)  # CPU tensor: no fallback


def test_chained_1x1_matches_two_convs(cuda_dev):
    """conv3x3(32->64)+ReLU with the 1x1 (64->48, split 12|36, fp32) chained inside the same kernel ==
    the two-kernel composition in fp32 torch (heads: DetModelBase.py:283-351)."""
    from disconet_b200.plan import pack_chain
    dev = cuda_dev
    g = torch.Generator().manual_seed(7)
    n, h, w = 3, 40, 48
    x = torch.randn(n, 32, h, w, generator=g).to(dev)
    w1 = (torch.randn(64, 32, 3, 3, generator=g) / (32 * 9) ** 0.5).to(dev)
    b1 = (torch.randn(64, generator=g) * 0.1).to(dev)
    w2 = (torch.randn(48, 64, generator=g) / 8).to(dev)
    b2 = (torch.randn(48, generator=g) * 0.1).to(dev)
    a = to_act(x, PREC_BF16X3)
    xv = act_value(a).permute(0, 3, 1, 2)
    mid = F.relu(F.conv2d(xv, w1, b1, padding=1))
    ref = F.conv2d(mid, w2.view(48, 64, 1, 1), b2).permute(0, 2, 3, 1)
    plan = pack_conv(w1, b1, src_channels=[32], relu=True, precision=PREC_BF16X3, c_blk=16, name="chain_test")
    assert plan.stacked
    plan.chain = pack_chain(w2, b2, 64)
    o0 = torch.full((n, h, w, 12), float("nan"), device=dev)
    o1 = torch.full((n, h, w, 36), float("nan"), device=dev)
    conv_forward(plan, [a], [0], (o0, o1), n=n, h_in=h, w_in=w, out_split=12)
    torch.cuda.synchronize()
    got = torch.cat((o0, o1), -1)
    assert torch.isfinite(got).all()
    assert rel_max(got, ref) < 1e-4


def test_conv_fuzz_against_validator(cuda_dev):
    """Seeded random layer shapes (ragged sizes, both strides, 1 or 2 sources with/without upsample, 16..512
    channels): the tcgen05 kernel must agree with the CUDA-core validator (same inputs, unpacked weights)."""
    import random
    rnd = random.Random(1234)
    chans = [16, 32, 48, 64, 96, 128, 256]
    for trial in range(28):
        k = rnd.choice([3, 3, 3, 1])
        stride = rnd.choice([1, 1, 2]) if k == 3 else 1
        two = k == 3 and stride == 1 and rnd.random() < 0.4
        n = rnd.randint(1, 3)
        h, w = rnd.choice([8, 16, 24, 40, 56]), rnd.choice([8, 16, 20, 48, 72])
        if two:
            h, w = h + (h & 1), w + (w & 1)
        c0 = rnd.choice(chans)
        srcs = [(c0, c0, 1 if two else 0)] + ([(rnd.choice(chans[:5]), None, 0)] if two else [])
        srcs = [(c, c, u) for c, _, u in srcs]
        c_out = rnd.choice([16, 32, 64, 128, 256, 512] if k == 3 else [16, 48, 64, 256])
        case = (f"fuzz{trial}", n, h, w, srcs, c_out, k, stride, "act")
        prec = rnd.choice([PREC_BF16X3, PREC_BF16X3, PREC_FP16])
        results, ref = _run_case(case, prec, cuda_dev, seed=100 + trial)
        tol = {PREC_FP16: 1e-3, PREC_BF16X3: 5e-5}[prec]
        for which, got in results.items():
            assert torch.isfinite(got).all(), (case, which)
            err = rel_max(got, ref)
            assert err < tol, f"{case} {which}: {err:.3e}"


# ---- output-parity ("sub-pixel") decomposition of the upsample-concat convs (conv.h `subpix`, plan.pack_conv_subpix) ----------
SUBPIX_CASES = [
    # (name, n, h, w, c_up, c_skip, c_out)   -- conv8_1 / conv7_1 / conv6_1 / conv5_1 shape classes + ragged / tiny grids
    ("c8_1_like", 2, 64, 64, 64, 32, 32),
    ("c7_1_like", 2, 32, 32, 128, 64, 64),
    ("c6_1_like_many_items", 5, 64, 64, 256, 128, 128),
    ("c5_1_like", 2, 32, 32, 512, 256, 256),
    ("ragged_class_grid_24x20", 1, 24, 20, 64, 32, 32),
    ("tiny_8x8", 3, 8, 8, 128, 64, 64),
]


@pytest.mark.parametrize("case", SUBPIX_CASES, ids=[c[0] for c in SUBPIX_CASES])
def test_subpix_classes_match_torch_conv(cuda_dev, case):
    """Four class launches == F.conv2d(cat(interpolate(a, 2), b)) (Backbone.py:176-178,195-197,214-216,233-235)."""
    from disconet_b200.ops import ConvCall
    from disconet_b200.plan import pack_conv_subpix
    name, n, h, w, c_up, c_skip, c_out = case
    dev = cuda_dev
    g = torch.Generator(device="cpu").manual_seed(5)
    a = to_act(torch.randn(n, c_up, h // 2, w // 2, generator=g).to(dev), PREC_BF16X3)
    b = to_act(torch.randn(n, c_skip, h, w, generator=g).to(dev), PREC_BF16X3)
    c_in = c_up + c_skip
    wgt = (torch.randn(c_out, c_in, 3, 3, generator=g) / (c_in * 9) ** 0.5).to(dev)
    bias = (torch.randn(c_out, generator=g) * 0.1).to(dev)
    x_cat = torch.cat((F.interpolate(act_value(a).permute(0, 3, 1, 2), scale_factor=2), act_value(b).permute(0, 3, 1, 2)), 1)
    ref = F.relu(F.conv2d(x_cat.double(), wgt.double(), bias.double(), padding=1)).permute(0, 2, 3, 1).float()
    out = alloc_act(n, h, w, c_out, PREC_BF16X3, dev)
    out.fill_(float("nan"))
    stream = torch.cuda.current_stream(dev).cuda_stream
    for py in (0, 1):
        for px in (0, 1):
            plan = pack_conv_subpix(wgt, bias, src_channels=[c_up, c_skip], py=py, px=px, relu=True, name=name)
            ConvCall(plan, [a, b], [1, 0], out, n=n, h_in=h, w_in=w).launch(stream)
    torch.cuda.synchronize()
    got = act_value(out)
    assert torch.isfinite(got).all(), "a class launch left output pixels unwritten"
    err = rel_max(got, ref)
    print(name, "sub-pixel rel-max error vs fp64 torch conv", err)
    assert err < 3e-5          # bf16x3 operands (~16 mantissa bits), fp32 accumulation


FUSED_SUBPIX_CASES = [
    # (name, n, h, w, c_up, c_skip, c_out)
    ("c8_1_like", 2, 64, 64, 64, 32, 32),
    ("c7_1_like_single_tmem_buffer", 2, 64, 32, 128, 64, 64),
    ("many_items_per_cta", 5, 256, 256, 64, 32, 32),
    ("ragged_low_res_grid_12x10", 1, 24, 20, 64, 32, 32),
    ("c_out_16", 3, 32, 48, 32, 16, 16),
    ("tiny_4x4_low_res", 3, 8, 8, 128, 64, 64),
]


@pytest.mark.parametrize("case", FUSED_SUBPIX_CASES, ids=[c[0] for c in FUSED_SUBPIX_CASES])
def test_fused_subpix_matches_torch_conv(cuda_dev, case):
    """One launch, four class accumulators per low-res tile (conv.h subpix == 2) == F.conv2d(cat(interpolate(a, 2), b))
    (Backbone.py:214-216,233-235: conv7_1 / conv8_1)."""
    from disconet_b200.ops import ConvCall
    from disconet_b200.plan import pack_conv_subpix_fused
    name, n, h, w, c_up, c_skip, c_out = case
    dev = cuda_dev
    g = torch.Generator(device="cpu").manual_seed(6)
    a = to_act(torch.randn(n, c_up, h // 2, w // 2, generator=g).to(dev), PREC_BF16X3)
    b = to_act(torch.randn(n, c_skip, h, w, generator=g).to(dev), PREC_BF16X3)
    c_in = c_up + c_skip
    wgt = (torch.randn(c_out, c_in, 3, 3, generator=g) / (c_in * 9) ** 0.5).to(dev)
    bias = (torch.randn(c_out, generator=g) * 0.1).to(dev)
    x_cat = torch.cat((F.interpolate(act_value(a).permute(0, 3, 1, 2), scale_factor=2), act_value(b).permute(0, 3, 1, 2)), 1)
    ref = F.relu(F.conv2d(x_cat.double(), wgt.double(), bias.double(), padding=1)).permute(0, 2, 3, 1).float()
    out = alloc_act(n, h, w, c_out, PREC_BF16X3, dev)
    out.fill_(float("nan"))
    plan = pack_conv_subpix_fused(wgt, bias, src_channels=[c_up, c_skip], relu=True, name=name)
    call = ConvCall(plan, [a, b], [1, 0], out, n=n, h_in=h, w_in=w)
    for _ in range(2):      # second launch: same result from warm tensor maps / L2
        call.launch(torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    got = act_value(out)
    assert torch.isfinite(got).all(), "output pixels left unwritten"
    err = rel_max(got, ref)
    print(name, "fused sub-pixel rel-max error vs fp64 torch conv", err)
    assert err < 3e-5          # bf16x3 operands (~16 mantissa bits), fp32 accumulation

"""GPU parity of the BEV-segmentation DiscoNet (BASELINE config 5): U-Net on the tcgen05 conv kernel,
MaxPool / bilinear-upsample streaming kernels, the fusion kernel at C = 512 -- against the CPU oracle on the same seeded
inputs and against the live-reference goldens.  Tolerance: max|ours - ref| / max|ref| <= 1e-3 per returned tensor."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import act_value, rel_l2, rel_max, to_act
from oracle import seg_oracle as S
from oracle.make_golden import SEG_CASES, SEG_KEYS, SEG_STRIDES, seg_case_inputs
from test_oracle_cpu import GOLD, _template

pytestmark = pytest.mark.gpu
P = 1


def test_maxpool_and_bilinear_upsample_kernels_match_torch(cuda_dev):
    from disconet_b200 import _lib as L
    lib = L.load()
    dev = cuda_dev
    st = torch.cuda.current_stream(dev).cuda_stream
    rng = np.random.default_rng(1)
    for (n, c, h, w) in [(2, 64, 32, 48), (1, 512, 16, 16), (3, 16, 8, 24)]:
        x = torch.from_numpy(rng.standard_normal((n, c, h, w)).astype(np.float32))
        a = to_act(x, P).to(dev)
        xq = act_value(a).cpu().permute(0, 3, 1, 2)
        pooled = torch.empty((2, n, h // 2, w // 2, c), dtype=torch.bfloat16, device=dev)
        up = torch.empty((2, n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=dev)
        L.check(lib.disco_maxpool2(a.data_ptr(), a.stride(0), pooled.data_ptr(), pooled.stride(0), P, n, h, w, c, st), "pool")
        L.check(lib.disco_upsample_bilinear2x(a.data_ptr(), a.stride(0), up.data_ptr(), up.stride(0), P, n, h, w, c, st), "up")
        torch.cuda.synchronize()
        assert torch.equal(act_value(pooled).cpu().permute(0, 3, 1, 2), F.max_pool2d(xq, 2))       # exact
        ref = F.interpolate(xq.double(), scale_factor=2, mode="bilinear", align_corners=True)
        assert rel_max(act_value(up).cpu().permute(0, 3, 1, 2), ref) < 2e-5
        src = torch.from_numpy(rng.standard_normal((n, h, w, 8)).astype(np.float32)).to(dev)
        out = torch.empty((n, 5, h, w), device=dev)
        L.check(lib.disco_nhwc_to_nchw(src.data_ptr(), n, h, w, 8, 5, out.data_ptr(), st), "nhwc_to_nchw")
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), src.cpu().permute(0, 3, 1, 2)[:, :5])


@pytest.mark.parametrize("name", list(SEG_CASES))
def test_seg_disconet_matches_oracle_and_golden(name, cuda_dev):
    from disconet_b200.seg import SegDiscoNet
    case = SEG_CASES[name]
    sd, x, T, na = seg_case_inputs(case, _template(name))
    ref = S.seg_disconet_forward(sd, x, T, na, agent_num=case["A"], only_v2i=case["only_v2i"], return_all=True)
    m = SegDiscoNet(13, 8, num_agent=case["A"], kd_flag=True, only_v2i=case["only_v2i"], compress_level=case.get("compress_level", 0))
    m.load_state_dict(sd)
    m = m.to(cuda_dev).eval()
    with torch.no_grad():
        out = m(x.to(cuda_dev), T, na)
    torch.cuda.synchronize()
    rec = np.load(os.path.join(GOLD, name + ".npz"))
    for k, g in zip(SEG_KEYS, out):
        g = g.float().cpu()
        assert g.shape == ref[k].shape and g.is_contiguous(), k
        e = rel_max(g, ref[k])
        print(f"{name} {k}: rel-max {e:.2e} rel-l2 {rel_l2(g, ref[k]):.2e}")
        assert e <= 1e-3, (k, e)
        sub = g.reshape(-1)[::SEG_STRIDES[k]].numpy()
        assert np.abs(sub - rec[k + "_sub"]).max() <= 1e-3 * rec[k + "_stats"][2], k
    # kd_flag = False returns the logits only
    m.kd_flag = False
    with torch.no_grad():
        lg = m(x.to(cuda_dev), T, na)
    assert torch.equal(lg, out[0])


def test_maxpool_and_upsample_backward_kernels_match_autograd(cuda_dev):
    from disconet_b200 import _lib as L
    lib = L.load()
    dev = cuda_dev
    st = torch.cuda.current_stream(dev).cuda_stream
    rng = np.random.default_rng(2)
    for (n, c, h, w) in [(2, 64, 32, 48), (1, 512, 16, 16), (2, 16, 8, 24)]:
        x = torch.from_numpy(np.maximum(rng.standard_normal((n, c, h, w)), 0).astype(np.float32))   # post-ReLU: many exact ties at 0
        a = to_act(x, P).to(dev)
        xq = act_value(a).cpu().permute(0, 3, 1, 2).double().requires_grad_(True)
        g = torch.from_numpy(rng.standard_normal((n, c, h // 2, w // 2)).astype(np.float32))
        F.max_pool2d(xq, 2).backward(g.double())
        gx = torch.empty((n, h, w, c), device=dev)
        g_d = g.permute(0, 2, 3, 1).contiguous().to(dev)
        L.check(lib.disco_maxpool2_backward(a.data_ptr(), a.stride(0), P, g_d.data_ptr(), gx.data_ptr(), n, h, w, c, st), "pool_bwd")
        torch.cuda.synchronize()
        assert torch.equal(gx.cpu().permute(0, 3, 1, 2).double(), xq.grad)
        # bilinear x2 (align_corners=True) backward
        src = torch.zeros((n, c, h, w), dtype=torch.float64, requires_grad=True)
        gu = torch.from_numpy(rng.standard_normal((n, c, 2 * h, 2 * w)).astype(np.float32))
        F.interpolate(src, scale_factor=2, mode="bilinear", align_corners=True).backward(gu.double())
        gs = torch.empty((n, h, w, c), device=dev)
        gu_d = gu.permute(0, 2, 3, 1).contiguous().to(dev)
        L.check(lib.disco_upsample_bilinear2x_backward(gu_d.data_ptr(), gs.data_ptr(), n, h, w, c, st), "up_bwd")
        torch.cuda.synchronize()
        assert rel_max(gs.cpu().permute(0, 3, 1, 2), src.grad) < 1e-5


def test_seg_training_step_matches_oracle(cuda_dev):
    """seg DiscoNet in train() mode behind the same autograd node: outputs <= 1e-3, gradients by norm / cosine (see
    tests/test_train_gpu.py for why gradients of this ReLU/BatchNorm stack are compared that way), BN buffers."""
    from disconet_b200.seg import SegDiscoNet
    from oracle.make_golden import SEG_TRAIN_CASE
    from test_oracle_cpu import oracle_seg_train_step
    from oracle import disconet_oracle as O
    case = SEG_TRAIN_CASE
    name = "seg_train_a2_b1"
    sd, x, T, na = seg_case_inputs(case, _template(name))
    ref_t, ref_loss, ref_g, ref_bufs = oracle_seg_train_step(case, sd)
    m = SegDiscoNet(13, 8, num_agent=case["A"], kd_flag=True, only_v2i=case["only_v2i"])
    m.load_state_dict(sd)
    m = m.to(cuda_dev).train()
    out = m(x.to(cuda_dev), T, na)
    got = dict(zip(SEG_KEYS, out))
    loss, _ = O.probe_loss(got, seed=case["seed"] + 300)
    loss.backward()
    torch.cuda.synchronize()
    for k in SEG_KEYS:
        e = rel_max(got[k].detach().cpu(), ref_t[k].detach())
        print(f"seg train fwd {k}: rel-max {e:.2e}")
        assert e <= 1e-3, (k, e)
    a, b = [], []
    for k, p in m.named_parameters():
        g = ref_g[k]
        assert p.grad is not None, k
        if p.grad.abs().max() == 0:       # BatchNorm-shadowed conv bias
            continue
        n_ref, n_ours = g.norm().item(), p.grad.norm().item()
        assert abs(n_ours - n_ref) <= 5e-2 * n_ref, (k, n_ours, n_ref)
        a.append(p.grad.detach().cpu().flatten()); b.append(g.flatten())
    a, b = torch.cat(a).double(), torch.cat(b).double()
    cos = (a @ b / (a.norm() * b.norm())).item()
    print(f"seg train all gradients: cos {cos:.6f} rel-l2 {((a - b).norm() / b.norm()).item():.3e}")
    assert cos >= 0.999
    for k, v in ref_bufs.items():
        if "num_batches" in k:
            assert int(m.state_dict()[k]) == int(v), k
        else:
            assert rel_max(m.state_dict()[k].float().cpu(), v.float()) <= 1e-3, k

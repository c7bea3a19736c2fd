"""GPU parity of the BEV-segmentation DiscoNet (BASELINE config 5): U-Net on the tcgen05 conv kernel,
MaxPool / bilinear-upsample streaming kernels, the fusion kernel at C = 512 -- against the CPU oracle on the same seeded
inputs and against the live-reference goldens.  Tolerance: max|ours - ref| / max|ref| <= 1e-3 per returned tensor."""
import ctypes as C
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
    m = SegDiscoNet(13, 8, num_agent=case["A"], kd_flag=True, only_v2i=case["only_v2i"])
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

"""One pass over the round-2 data-path kernels either side of the conv stack, for an ncu capture: batched voxel scatter ->
forward_voxels -> det_candidates / sort / polygon-IoU mask / greedy scan (post.detect), plus the fused corner loss.
    SCENES=16 ncu --set full -k regex:'nms_|det_cand|bev_scatter_batched|corner_loss' ... python tools/prof_post.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Cfg, AGENTS, synth_inputs
from disconet_b200 import DiscoNet, synth, post
from disconet_b200.loss import corner_loss
from disconet_b200.voxel import bev_to_voxel_indices

B = int(os.environ.get("SCENES", "16"))
dev = torch.device("cuda:0")
m = DiscoNet(Cfg(), kd_flag=0, num_agent=AGENTS)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0))
m = m.to(dev).eval()
bev, T, na = synth_inputs(B, 100)
idx, cnt = bev_to_voxel_indices(bev)
idx, cnt, T, na = idx.to(dev), cnt.to(dev), T.to(dev), na.to(dev)
anchors = synth.synth_anchors().to(dev)
with torch.no_grad():
    for _ in range(2):
        res, _ = m.forward_voxels(idx, cnt, T, na, batch_size=B)
        out = post.detect(res["loc"], res["cls"], anchors, device_only=True, max_candidates=2048)
n = AGENTS * B
g = torch.Generator().manual_seed(1)
mask = (torch.rand((n, 256, 256, 6, 1), generator=g) < 1e-3).to(dev)
tgt = (torch.randn((n, 256, 256, 6, 1, 6), generator=g) * 0.1).to(dev)
p = res["loc"].clone().requires_grad_(True)
corner_loss(anchors.unsqueeze(0).expand(n, -1, -1, -1, -1).contiguous(), mask, tgt, p).backward()
torch.cuda.synchronize()
print("candidates/agent", float(out[6].float().mean()), "kept/agent", float(out[4].float().mean()))

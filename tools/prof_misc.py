"""One pass over the HBM-bound kernels of the path for an `ncu --set full` capture: voxelize / scatter, one training step
(BatchNorm, PWF, fusion backward, packing kernels), the fused losses, the seg data-movement kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import Cfg, AGENTS, synth_inputs
from disconet_b200 import DiscoNet, synth, voxelize_occupy, bev_scatter
from disconet_b200.kd import kd_kl_mean
from disconet_b200.loss import SoftmaxFocalClassificationLoss
from disconet_b200.seg import SegDiscoNet

dev = torch.device("cuda:0")
B = int(os.environ.get("SCENES", "4"))
rng = np.random.default_rng(0)
pts = np.stack([rng.uniform(-40, 40, 40000), rng.uniform(-40, 40, 40000), rng.uniform(-3.5, 2.5, 40000), rng.uniform(0, 1, 40000)], 1).astype(np.float32)
p_d = torch.from_numpy(pts).to(dev)
ext = np.array([[-32.0, 32.0], [-32.0, 32.0], [-3.0, 2.0]])
for _ in range(2):
    grid, idx = voxelize_occupy(p_d, (0.25, 0.25, 0.4), ext, return_indices=True)
    bev = bev_scatter(idx, grid.shape)
m = DiscoNet(Cfg(), kd_flag=1, num_agent=AGENTS)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0))
m = m.to(dev).train()
os.environ["DISCO_B200_TRAIN_GRAPH"] = "0"
bevs, T, na = synth_inputs(B, 100)
bevs, na = bevs.to(dev), na.to(dev)
focal = SoftmaxFocalClassificationLoss()
labels = torch.zeros((AGENTS * B, 256 * 256 * 6, 2), device=dev)
labels[..., 0] = 1
for _ in range(2):
    res, x8, x7, x6, x5, fused = m(bevs, T, na, batch_size=B)
    loss = focal(res["cls"], labels).sum() / (AGENTS * B) + res["loc"].square().mean() + kd_kl_mean(x7, x7.detach() * 0.9) + \
        kd_kl_mean(x6, x6.detach() * 0.9) + kd_kl_mean(x5, x5.detach() * 0.9) + kd_kl_mean(fused, fused.detach() * 0.9)
    loss.backward()
s = SegDiscoNet(13, 8, num_agent=AGENTS, kd_flag=False)
s.load_state_dict(synth.synth_state_dict(s.state_dict(), seed=0))
s = s.to(dev).eval()
with torch.no_grad():
    s(bevs[:, 0].permute(0, 3, 1, 2).contiguous(), T.to(dev), na)
torch.cuda.synchronize()
print("done")

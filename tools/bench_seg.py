"""Throughput of the BEV-segmentation DiscoNet eval forward (BASELINE config 5: 5 agents, 256x256x13), CUDA-event timed.
    SCENES=8 STEPS=20 python tools/bench_seg.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from disconet_b200 import synth
from disconet_b200.seg import SegDiscoNet

A, B, STEPS = 5, int(os.environ.get("SCENES", "8")), int(os.environ.get("STEPS", "20"))
dev = torch.device("cuda:0")
m = SegDiscoNet(13, 8, num_agent=A, kd_flag=False)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0))
m = m.to(dev).eval()
x = synth.synth_bev(A * B, seed=100)[:, 0].permute(0, 3, 1, 2).contiguous().to(dev)
T = synth.synth_poses(B, A, seed=101).to(dev)
na = torch.full((B, A), A, device=dev)
with torch.no_grad():
    for _ in range(3):
        m(x, T, na)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        m(x, T, na)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / STEPS
ws = next(iter(m._ws.values()))
# training step (forward + backward + Adam on a cross-entropy loss, as SegModule.step does), 2 scenes per step
TB = int(os.environ.get("TRAIN_SCENES", "2"))
mt = SegDiscoNet(13, 8, num_agent=A, kd_flag=False)
mt.load_state_dict(synth.synth_state_dict(mt.state_dict(), seed=0))
mt = mt.to(dev).train()
opt = torch.optim.Adam(mt.parameters(), lr=1e-4)
xt, Tt, nat = x[:A * TB].contiguous(), synth.synth_poses(TB, A, seed=101), torch.full((TB, A), A, device=dev)
labels = torch.randint(0, 8, (A * TB, 256, 256), device=dev)


def tstep():
    loss = torch.nn.functional.cross_entropy(mt(xt, Tt, nat), labels)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(5):
    tstep()
torch.cuda.synchronize()
e0.record()
for _ in range(STEPS):
    tstep()
e1.record()
torch.cuda.synchronize()
ms_t = e0.elapsed_time(e1) / STEPS
print(json.dumps({"metric": "scenes/sec, seg DiscoNet eval forward", "value": B / (ms / 1e3), "ms_per_step": ms,
                  "scenes_per_step": B, "agents": A, "conv_gflop_per_scene": ws.flops / 1e9 / B,
                  "algorithmic_tflops": ws.flops / 1e9 / ms, "dtype": "bf16x3", "data": "synthetic",
                  "train": {"metric": "scenes/sec, seg DiscoNet training step (forward + CE loss + backward + Adam)",
                            "value": TB / (ms_t / 1e3), "ms_per_step": ms_t, "scenes_per_step": TB,
                            "algorithmic_tflops": 3 * ws.flops / 1e9 / B * TB / ms_t}}))

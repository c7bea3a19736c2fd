"""Training steps (forward + backward of a probe loss) of the bench workload: CUDA-event timing per phase, or run
under ncu for the per-kernel launch list.   SCENES=1 STEPS=5 python tools/prof_train.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Cfg, AGENTS, synth_inputs
from disconet_b200 import DiscoNet, synth

B = int(os.environ.get("SCENES", "1"))
STEPS = int(os.environ.get("STEPS", "5"))
dev = torch.device("cuda:0")
m = DiscoNet(Cfg(), kd_flag=1, num_agent=AGENTS)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0))
m = m.to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
bev, T, na = synth_inputs(B, 100)
bev, na = bev.to(dev), na.to(dev)


def step():
    t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t[0].record()
    res, x8, x7, x6, x5, fused = m(bev, T, na, batch_size=B)
    t[1].record()
    loss = res["cls"].square().mean() + res["loc"].square().mean() + x7.mean() + x6.mean() + x5.mean() + fused.mean()
    opt.zero_grad()
    loss.backward()
    t[2].record()
    opt.step()
    t[3].record()
    return t, loss


for i in range(STEPS):
    w0 = time.perf_counter()
    t, loss = step()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    print(f"step {i}: fwd {t[0].elapsed_time(t[1]):.2f} ms  loss+bwd {t[1].elapsed_time(t[2]):.2f} ms  adam {t[2].elapsed_time(t[3]):.2f} ms  "
          f"wall {(w1 - w0) * 1e3:.2f} ms  loss {loss.item():.5f}  ({B} scenes x {AGENTS} agents)")
print("done")

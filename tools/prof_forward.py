"""Two eval forwards (warm-up + one to profile) of the bench workload; run under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Cfg, AGENTS, synth_inputs
from disconet_b200 import DiscoNet, synth

B = int(os.environ.get("SCENES", "8"))
prec = os.environ.get("PRECISION", "bf16x3")
dev = torch.device("cuda:0")
m = DiscoNet(Cfg(), kd_flag=0, num_agent=AGENTS, precision=prec)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0))
m = m.to(dev).eval()
bev, T, na = synth_inputs(B, 100)
bev, T, na = bev.to(dev), T.to(dev), na.to(dev)
with torch.no_grad():
    for _ in range(2):
        m(bev, T, na, batch_size=B)
torch.cuda.synchronize()
print("done")

"""Timing experiments on the fused sub-pixel layers (conv_tc.cu MODE 4): which role bounds an item?
DISCO_CONV_DBG bits (results are wrong when set): 1 no weight copies, 2 no A copies, 4 no output stores, 8 no MMAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Cfg, AGENTS, synth_inputs
from disconet_b200 import DiscoNet, synth
dev = torch.device("cuda:0")
B = int(os.environ.get("SCENES", "16"))
m = DiscoNet(Cfg(), kd_flag=0, num_agent=AGENTS)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0)); m = m.to(dev).eval()
bev, T, na = synth_inputs(B, 100)
with torch.no_grad():
    m(bev.to(dev), T, na, batch_size=B)
ws = next(iter(m._ws.values()))
calls = {c.plan.name.split(".")[-1]: c for c in ws.dec_calls}
stream = torch.cuda.current_stream(dev).cuda_stream
def t(call, reps=5):
    call.launch(stream); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): call.launch(stream)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
envs = [e.split("=") for e in os.environ.get("SWEEP", "").split(",") if e]
for name in sys.argv[1:]:
    c = calls[name]
    for dbg in (0, 1, 2, 4, 8, 3, 7, 9, 11, 15):
        os.environ["DISCO_CONV_DBG"] = str(dbg)
        print(f"{name} dbg={dbg:2d}: {t(c):.4f} ms", flush=True)
    os.environ["DISCO_CONV_DBG"] = "0"
    for k, v in envs:
        os.environ[k] = v
        print(f"{name} {k}={v}: {t(c):.4f} ms", flush=True)
        del os.environ[k]

"""Debug: per-role clock64 trace of CTA 0 for one conv layer (env DISCO_CONV_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Cfg, AGENTS, synth_inputs
from disconet_b200 import DiscoNet, synth
dev = torch.device("cuda:0")
B = int(os.environ.get("TRACE_B", "16"))
m = DiscoNet(Cfg(), kd_flag=0, num_agent=AGENTS, precision=os.environ.get("PRECISION", "bf16x3"))
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0)); m = m.to(dev).eval()
bev, T, na = synth_inputs(B, 100)
with torch.no_grad():
    m(bev.to(dev), T, na, batch_size=B)
ws = next(iter(m._ws.values()))
calls = {c.plan.name.split(".")[-1] if "conv3d" not in c.plan.name else c.plan.name: c for c in ws.enc_calls + ws.dec_calls + ws.head_calls}
stream = torch.cuda.current_stream(dev).cuda_stream
for name in sys.argv[1:]:
    c = calls[name]
    tr = torch.zeros(4 * 64 * 4, dtype=torch.int64, device=dev)
    os.environ["DISCO_CONV_TRACE"] = str(tr.data_ptr())
    c.launch(stream); torch.cuda.synchronize()
    del os.environ["DISCO_CONV_TRACE"]
    t = tr.cpu().view(4, 64, 4)
    t0 = t[t > 0].min().item()
    print(f"== {name}: stamps in cycles since first stamp (CTA 0)")
    print("item | epi: wait_start got_acc done first_ld_done | mma: wait_accempty got start_issue committed | prod0(stage): wait_empty got issued landed")
    for i in range(12):
        f = lambda r: " ".join(f"{(x - t0) if x > 0 else -1:7d}" for x in t[r, i].tolist())
        print(f"{i:3d} | {f(0)} | {f(2)} | {f(1)}" + (f" | epi detail (chunk 0 math done, staged, stored; chunk 1 ready): {f(3)}" if (t[3, i] > 0).any() else ""))

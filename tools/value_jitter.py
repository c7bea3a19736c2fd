"""Why does bench.py's device-resident `value` window intermittently run 15-60 % slow?  Times the same 20-step window repeatedly
with (a) no clock sampler, (b) the `nvidia-smi -lms 100` subprocess, (c) an in-process NVML thread."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Cfg, AGENTS, synth_inputs, ClockSampler
from disconet_b200 import DiscoNet, synth
dev = torch.device("cuda:0")
B, K = 16, 20
m = DiscoNet(Cfg(), kd_flag=0, num_agent=AGENTS)
m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0)); m = m.to(dev).eval()
bev, T, na = (x.to(dev) for x in synth_inputs(B, 100))

def window():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0 = torch.cuda.memory_reserved()
    e0.record()
    for _ in range(K):
        res, _ = m(bev, T, na, batch_size=B)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, (torch.cuda.memory_reserved() - r0) >> 20

class NvmlThread:
    def __init__(self):
        import pynvml
        self.n = pynvml; pynvml.nvmlInit(); self.h = pynvml.nvmlDeviceGetHandleByIndex(0); self.rows = []; self.stop_ = False
    def start(self):
        def run():
            while not self.stop_:
                self.rows.append((time.time(), self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM),
                                  self.n.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
                time.sleep(0.02)
        self.t = threading.Thread(target=run, daemon=True); self.t.start()
    def stop(self):
        self.stop_ = True; self.t.join()
        return len(self.rows), sorted({r[1] for r in self.rows})[:3], hex(max(r[2] for r in self.rows))

with torch.no_grad():
    for _ in range(5): m(bev, T, na, batch_size=B)
    torch.cuda.synchronize()
    for mode in ("none", "nvidia-smi", "nvml-thread", "none", "nvidia-smi", "nvml-thread"):
        s = None
        if mode == "nvidia-smi":
            s = ClockSampler(0); s.start(); time.sleep(0.5)
        if mode == "nvml-thread":
            s = NvmlThread(); s.start(); time.sleep(0.1)
        out = [window() for _ in range(6)]
        info = s.stop() if s else None
        print(f"{mode:12s} ms/step:", " ".join(f"{a:.2f}" for a, _ in out), "| reserved growth MiB:", [b for _, b in out], "|", info, flush=True)

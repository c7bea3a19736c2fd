"""Host <-> device pipelining for the eval path: overlaps the pinned-host -> HBM copy of step i+1 and the
logits -> pinned-host copy of step i-1 with the kernels of step i (PCIe gen5 is full duplex and the copy
engines run beside the SMs).  The reference's eval loop (tools/det/test_codet.py:215-280) does
`.to(device)` -> forward -> CPU post-processing strictly in sequence."""
from __future__ import annotations

from typing import Callable, Optional

import torch


class HostPipeline:
    """Double-buffered eval runner around a `disconet_b200.DiscoNet`.

        pipe = HostPipeline(model, batch_size=B)
        for bev_host, trans, num_agent in loader:           # pinned host tensors
            done = pipe.submit(bev_host, trans, num_agent)    # returns the slot whose results are now on the host
        pipe.flush()

    Results of step i land in `pipe.cls_host[i % 2]`, `pipe.loc_host[i % 2]` (pinned) and are valid once
    `pipe.d2h_done[i % 2]` has completed (`submit`/`flush` synchronise on it before reusing the slot).
    """

    def __init__(self, model, batch_size: int, device: Optional[torch.device] = None):
        self.model = model
        self.B = int(batch_size)
        self.device = device or next(model.parameters()).device
        self.compute = torch.cuda.current_stream(self.device)
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.x = [None, None]
        self.cls_host = [None, None]
        self.loc_host = [None, None]
        self.h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.comp_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.d2h_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.step = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @torch.no_grad()
    def submit(self, bev_host: torch.Tensor, trans: torch.Tensor, num_agent: torch.Tensor) -> int:
        s = self.step % 2
        if self.x[s] is None:
            self.x[s] = torch.empty(bev_host.shape, dtype=bev_host.dtype, device=self.device)
        # slot reuse: the forward of step-2 must have consumed x[s]; its results must have left cls/loc_host[s]
        if self.step >= 2:
            self.h2d.wait_event(self.comp_done[s])
            self.d2h_done[s].synchronize()
        with torch.cuda.stream(self.h2d):
            self.x[s].copy_(bev_host, non_blocking=True)
            self.h2d_done[s].record(self.h2d)
        self.compute.wait_event(self.h2d_done[s])
        res, _ = self.model(self.x[s], trans, num_agent, batch_size=self.B)
        self.comp_done[s].record(self.compute)
        cls, loc = res["cls"], res["loc"]
        if self.cls_host[s] is None:
            self.cls_host[s] = torch.empty(cls.shape, dtype=cls.dtype).pin_memory()
            self.loc_host[s] = torch.empty(loc.shape, dtype=loc.dtype).pin_memory()
        self.d2h.wait_event(self.comp_done[s])
        with torch.cuda.stream(self.d2h):
            self.cls_host[s].copy_(cls, non_blocking=True)
            self.loc_host[s].copy_(loc, non_blocking=True)
            self.d2h_done[s].record(self.d2h)
        cls.record_stream(self.d2h)
        loc.record_stream(self.d2h)
        self.h2d_bytes = bev_host.numel() * bev_host.element_size() + trans.numel() * trans.element_size() + \
            num_agent.numel() * num_agent.element_size()
        self.d2h_bytes = cls.numel() * 4 + loc.numel() * 4
        self.step += 1
        return s

    def flush(self) -> None:
        for e in self.d2h_done:
            e.synchronize()
        self.compute.synchronize()

"""Knowledge-distillation loss of the training step on one fused kernel per feature map.

Mirror of `FaFModule.get_kd_loss` (coperception/utils/CoDetModule.py:312-388): the teacher's early-fusion maps
(x_7, x_6, x_5, x_3) against the student's (x_7, x_6, x_5, fused), each term
`nn.KLDivLoss(size_average=True, reduce=True)(log_softmax(student_pixels, 1), softmax(teacher_pixels, 1))` -- the MEAN
over all elements -- summed and scaled by `kd_weight`.  The reference permutes each map to [pixels, C] and runs five
torch ops per map (plus their autograd backward); `kd_kl_mean` computes the value and the gradient wrt the student map
in one kernel launch on the NCHW tensors the model returns.
"""
from __future__ import annotations

import torch

from ._lib import check, load


class _KdKl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher):
        if not (student.is_cuda and teacher.is_cuda):
            raise ValueError("disconet_b200.kd runs on CUDA tensors only (no CPU fallback)")
        if student.shape != teacher.shape or student.dim() != 4:
            raise ValueError(f"student {tuple(student.shape)} and teacher {tuple(teacher.shape)} must be equal NCHW maps")
        s = student.detach().float().contiguous()
        t = teacher.detach().float().contiguous()
        n, c, h, w = s.shape
        numel = s.numel()
        acc = torch.zeros((), dtype=torch.float64, device=s.device)
        grad = torch.empty_like(s) if student.requires_grad else None
        stream = torch.cuda.current_stream(s.device).cuda_stream
        check(load().disco_kd_kl(s.data_ptr(), t.data_ptr(), n, c, h * w, acc.data_ptr(),
                                 grad.data_ptr() if grad is not None else None, 1.0 / numel, stream), "kd_kl")
        ctx.grad = grad
        return (acc / numel).float()

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None


def kd_kl_mean(student: torch.Tensor, teacher: torch.Tensor) -> torch.Tensor:
    """KLDivLoss(mean over elements)(log_softmax(student, C), softmax(teacher, C)) for NCHW maps; differentiable wrt student."""
    return _KdKl.apply(student, teacher)


def get_kd_loss(self, batch_size, data, fused_layer, num_agent, x_5, x_6, x_7):
    """Drop-in for FaFModule.get_kd_loss (CoDetModule.py:312-388); bound onto FaFModule by disconet_b200.patch."""
    if not self.kd_flag:
        return 0
    bev_seq_teacher = data["bev_seq_teacher"]
    kd_weight = data["kd_weight"]
    (x_8_teacher, x_7_teacher, x_6_teacher, x_5_teacher, x_3_teacher, x_2_teacher) = self.teacher(bev_seq_teacher)
    return kd_weight * (kd_kl_mean(x_7, x_7_teacher) + kd_kl_mean(x_6, x_6_teacher) + kd_kl_mean(x_5, x_5_teacher)
                        + kd_kl_mean(fused_layer, x_3_teacher))

"""Classification loss of the training step as one fused kernel each way.

`SoftmaxFocalClassificationLoss` mirrors coperception.utils.loss.SoftmaxFocalClassificationLoss (loss.py:322-394): same
constructor, same call signature `(prediction_tensor, target_tensor, weights=None)`, same [N, anchors, classes] result
(`FaFModule.loss_calculator` sums it and divides by N, CoDetModule.py:121).  The reference evaluates ~15 elementwise
torch ops + CrossEntropyLoss over 393 216 anchors per agent and lets autograd walk them back; here the forward is one
launch and the backward is one launch (the gradient of `torch.sum` arrives as a broadcast scalar and is read as such).
"""
from __future__ import annotations

import torch

from ._lib import check, load


class _Focal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, gamma, alpha):
        if not (logits.is_cuda and target.is_cuda):
            raise ValueError("disconet_b200.loss runs on CUDA tensors only (no CPU fallback)")
        if logits.shape != target.shape or logits.shape[-1] > 8:
            raise ValueError(f"logits {tuple(logits.shape)} / targets {tuple(target.shape)}: equal shapes, <= 8 classes")
        z = logits.detach().float().contiguous()
        t = target.detach().float().contiguous()
        out = torch.empty_like(z)
        k = z.shape[-1]
        stream = torch.cuda.current_stream(z.device).cuda_stream
        check(load().disco_focal_loss(z.data_ptr(), t.data_ptr(), k, z.numel() // k, float(gamma or 0.0),
                                      float(alpha if alpha is not None else 0.0), int(alpha is not None), None, 0,
                                      out.data_ptr(), stream), "focal_loss")
        ctx.save_for_backward(z, t)
        ctx.cfg = (gamma, alpha)
        return out

    @staticmethod
    def backward(ctx, g):
        z, t = ctx.saved_tensors
        gamma, alpha = ctx.cfg
        k = z.shape[-1]
        if all(s == 0 for s in g.stride()):            # backward of torch.sum: one broadcast scalar
            gbuf, stride = g.as_strided((1,), (1,)).float().contiguous(), 0
        else:
            gbuf, stride = g.float().contiguous(), 1
        dz = torch.empty_like(z)
        stream = torch.cuda.current_stream(z.device).cuda_stream
        check(load().disco_focal_loss(z.data_ptr(), t.data_ptr(), k, z.numel() // k, float(gamma or 0.0),
                                      float(alpha if alpha is not None else 0.0), int(alpha is not None), gbuf.data_ptr(), stride,
                                      dz.data_ptr(), stream), "focal_loss_backward")
        return dz, None, None, None


class SoftmaxFocalClassificationLoss:
    """Drop-in for coperception.utils.loss.SoftmaxFocalClassificationLoss (gamma=2.0, alpha=0.25)."""

    def __init__(self, gamma=2.0, alpha=0.25):
        self._alpha = alpha
        self._gamma = gamma

    def __call__(self, prediction_tensor, target_tensor, ignore_nan_targets=False, scope=None, **params):
        if ignore_nan_targets:
            target_tensor = torch.where(torch.isnan(target_tensor), prediction_tensor, target_tensor)
        return self._compute_loss(prediction_tensor, target_tensor, **params)

    def _compute_loss(self, prediction_tensor, target_tensor, weights=None, class_indices=None):
        if class_indices is not None:
            raise NotImplementedError("class_indices is not used on the DiscoNet path")
        loss = _Focal.apply(prediction_tensor, target_tensor, self._gamma, self._alpha)
        return loss * weights if weights is not None else loss


class _CornerLoss(torch.autograd.Function):
    """`FaFModule.corner_loss` (utils/CoDetModule.py:80-105) as one kernel: value and the gradient wrt the regression map."""

    @staticmethod
    def forward(ctx, pred, anchors, mask, targets):
        if not (pred.is_cuda and anchors.is_cuda and mask.is_cuda and targets.is_cuda):
            raise ValueError("disconet_b200.loss runs on CUDA tensors only (no CPU fallback)")
        if pred.shape != targets.shape or pred.shape[-1] != 6 or tuple(mask.shape) != tuple(pred.shape[:-1]):
            raise ValueError(f"pred {tuple(pred.shape)} / targets {tuple(targets.shape)} / mask {tuple(mask.shape)} do not match")
        t_len = pred.shape[-2]
        if anchors.numel() * t_len != pred.numel():
            raise ValueError(f"anchors {tuple(anchors.shape)} do not match pred {tuple(pred.shape)}")
        p = pred.detach().float().contiguous()
        t = targets.detach().float().contiguous()
        a = anchors.detach().float().contiguous()
        m = mask.detach().to(torch.uint8).contiguous()
        n = pred.shape[0]
        acc = torch.empty((), dtype=torch.float64, device=p.device)
        grad = torch.empty_like(p) if pred.requires_grad else None
        check(load().disco_corner_loss(p.data_ptr(), t.data_ptr(), a.data_ptr(), m.data_ptr(), p.numel() // 6, t_len, 1.0 / n,
                                       acc.data_ptr(), grad.data_ptr() if grad is not None else None,
                                       torch.cuda.current_stream(p.device).cuda_stream), "corner_loss")
        ctx.grad = grad
        return (acc / n).float()

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None, None, None


def corner_loss(anchors, reg_loss_mask, reg_targets, pred_result):
    """Same arguments and result as `FaFModule.corner_loss(self, anchors, reg_loss_mask, reg_targets, pred_result)`:
    anchors [N,H,W,A,6], reg_loss_mask [N,H,W,A,T] bool, reg_targets / pred_result [N,H,W,A,T,6] -> scalar
    sum over assigned anchors and corners of ||pred corner - target corner|| / N; differentiable wrt pred_result."""
    return _CornerLoss.apply(pred_result, anchors, reg_loss_mask, reg_targets)


def _corner_loss_method(self, anchors, reg_loss_mask, reg_targets, pred_result):
    return corner_loss(anchors, reg_loss_mask, reg_targets, pred_result)

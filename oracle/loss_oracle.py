"""CPU oracle of the training-step losses -- TEST INFRASTRUCTURE, never imported by the product.

Plain torch (float64 by default) restatements, R = /root/reference/coperception/coperception/utils:
  kd_loss      FaFModule.get_kd_loss                                       R/CoDetModule.py:312-388
  focal_loss   SoftmaxFocalClassificationLoss._compute_loss                R/loss.py:213-219,322-394
(the corner loss lives in oracle/post_oracle.py::corner_loss next to the box decode it shares with the NMS path).
Pinned against the reference objects themselves by tests/golden/losses.npz (oracle/make_golden_losses.py).
"""
import torch


def kl_mean(student: torch.Tensor, teacher: torch.Tensor) -> torch.Tensor:
    """nn.KLDivLoss(size_average=True, reduce=True)(log_softmax(s_pixels, 1), softmax(t_pixels, 1)) on NCHW maps:
    mean over ALL elements of t * (log t - log_softmax(s)) with the channel axis as the softmax axis (:334-350)."""
    s = student.permute(0, 2, 3, 1).reshape(-1, student.shape[1])
    t = teacher.permute(0, 2, 3, 1).reshape(-1, teacher.shape[1])
    logp = torch.log_softmax(s, dim=1)
    q = torch.softmax(t, dim=1)
    return (q * (torch.log(q) - logp)).mean()


def kd_loss(students, teachers, kd_weight):
    """students = (x_7, x_6, x_5, fused), teachers = (x_7, x_6, x_5, x_3) -> kd_weight * sum of the four KL terms (:379-381)."""
    total = 0
    for s, t in zip(students, teachers):
        total = total + kl_mean(s, t)
    return kd_weight * total


def focal_loss(logits: torch.Tensor, target: torch.Tensor, gamma: float = 2.0, alpha: float = 0.25) -> torch.Tensor:
    """[N, anchors, C] logits / one-hot targets -> per-entry focal loss [N, anchors, C] (loss.py:350-394)."""
    logp = torch.log_softmax(logits, dim=-1)
    label = target.max(dim=-1)[1]
    ce = -logp.gather(-1, label.unsqueeze(-1))                      # CrossEntropyLoss(reduction='none') on argmax labels
    per_entry = ce * target
    prob = torch.softmax(logits, dim=-1)
    p_t = target * prob + (1 - target) * (1 - prob)
    mod = torch.pow(1.0 - p_t, gamma) if gamma else 1.0
    aw = 1.0
    if alpha is not None:
        aw = torch.where(target[..., 0] == 1, torch.tensor(1 - alpha, dtype=logits.dtype), torch.tensor(alpha, dtype=logits.dtype)).unsqueeze(-1)
    return mod * aw * per_entry

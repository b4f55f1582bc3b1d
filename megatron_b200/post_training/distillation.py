"""Knowledge distillation (reference ``megatron/post_training`` + ``core/post_training/modelopt`` distillation specs).

``DistillationModel`` runs a frozen teacher next to the student on the same batch; the loss is
``(1 - alpha) * CE(student, labels) + alpha * T^2 * KL(softmax(teacher/T) || softmax(student/T))`` computed over the vocab-parallel logits
(the log-sum-exp terms are all-reduced over the TP group, so no rank ever gathers the full vocabulary), plus an optional MSE between
selected intermediate hidden states (with a linear projector when widths differ)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def vocab_parallel_kl(student_logits: torch.Tensor, teacher_logits: torch.Tensor, temperature: float = 1.0, tp_group=None) -> torch.Tensor:
    """KL(teacher || student) per token for logits sharded along the last dim over ``tp_group``.  [..., v/tp] → [...]."""
    s, t = student_logits.float() / temperature, teacher_logits.float() / temperature
    ws = dist.get_world_size(tp_group) if (tp_group is not None and dist.is_initialized()) else 1

    def lse(x):
        m = x.max(dim=-1, keepdim=True).values
        if ws > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=tp_group)
        e = torch.exp(x - m).sum(dim=-1, keepdim=True)
        if ws > 1:
            e = _AllReduceSum.apply(e, tp_group)
        return m + torch.log(e)

    ls, lt = s - lse(s), t - lse(t)
    kl = (torch.exp(lt) * (lt - ls)).sum(dim=-1)
    if ws > 1:
        kl = _AllReduceSum.apply(kl, tp_group)
    return kl * temperature * temperature


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        x = x.clone()
        dist.all_reduce(x, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


class DistillationModel(torch.nn.Module):
    def __init__(self, student: torch.nn.Module, teacher: torch.nn.Module, alpha: float = 0.5, temperature: float = 1.0, tp_group=None,
                 hidden_loss_weight: float = 0.0, student_width: Optional[int] = None, teacher_width: Optional[int] = None):
        super().__init__()
        self.student, self.teacher = student, teacher
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.teacher.eval()
        self.alpha, self.temperature, self.tp_group, self.hidden_w = alpha, temperature, tp_group, hidden_loss_weight
        self.proj = None
        if hidden_loss_weight > 0 and student_width and teacher_width and student_width != teacher_width:
            self.proj = torch.nn.Linear(student_width, teacher_width, bias=False)

    def forward(self, input_ids, position_ids, attention_mask, labels, loss_mask=None):
        """Returns (loss, {"ce", "kd"}).  ``labels`` [b, s]."""
        s_logits = self.student(input_ids, position_ids, attention_mask)            # [b, s, v/tp]
        with torch.no_grad():
            t_logits = self.teacher(input_ids, position_ids, attention_mask)
        kd = vocab_parallel_kl(s_logits, t_logits, self.temperature, self.tp_group)   # [b, s]
        ce = self.student.compute_language_model_loss(labels, s_logits.transpose(0, 1).contiguous())
        m = torch.ones_like(kd) if loss_mask is None else loss_mask.float()
        denom = m.sum().clamp(min=1)
        ce_m, kd_m = (ce * m).sum() / denom, (kd * m).sum() / denom
        return (1 - self.alpha) * ce_m + self.alpha * kd_m, {"ce": ce_m.detach(), "kd": kd_m.detach()}

"""Distillation criteria selectable by ``quantization.params.distillation_loss``
(reference: src/aux/loss/*.py, gdnsq_quant.py:40-66).  Logits-sized, plain PyTorch."""
import torch.nn.functional as F
from torch import nn


class SymmetricalKL(nn.Module):
    """KL(student||teacher) + KL(teacher||student), batchmean (symm_kl_loss.py:6-14)."""

    def forward(self, input, target):
        x, y = F.log_softmax(input, dim=1), F.log_softmax(target, dim=1)
        return (F.kl_div(x, y, log_target=True, reduction="batchmean")
                + F.kl_div(y, x, log_target=True, reduction="batchmean"))


class KL(nn.Module):
    def forward(self, input, target):
        return F.kl_div(F.log_softmax(input, dim=1), F.log_softmax(target, dim=1), log_target=True)


class SoftCrossEntropy(nn.Module):
    """Cross-entropy against the teacher's soft labels."""

    def forward(self, input, target):
        return -(F.softmax(target, dim=1) * F.log_softmax(input, dim=1)).sum(1).mean()


class SymmetricalSoftCrossEntropy(nn.Module):
    """Soft cross-entropy in both directions (symm_ce_loss.py:8-15)."""

    def forward(self, input, target):
        a = (F.softmax(target, dim=1) * F.log_softmax(input, dim=1)).sum(1).mean()
        b = (F.softmax(input, dim=1) * F.log_softmax(target, dim=1)).sum(1).mean()
        return -(a + b)


class Hellinger(nn.Module):
    """Mean squared difference of the square-rooted probabilities (hellinger.py:6-12)."""

    def forward(self, input, target):
        return F.mse_loss(input.softmax(-1).sqrt(), target.softmax(-1).sqrt())


class JSD(nn.Module):
    """Divergence of student and teacher from their log-domain midpoint, default ('mean') reduction
    of `kl_div` like the reference (jsdloss.py:6-16)."""

    def forward(self, input, target):
        p, q = input.log_softmax(-1), target.log_softmax(-1)
        m = 0.5 * (p + q)
        return F.kl_div(m, p, log_target=True) + F.kl_div(m, q, log_target=True)


def get_distillation_loss(name: str):
    table = {"Symmetrical KL": SymmetricalKL, "KL": KL, "Cross-Entropy": SoftCrossEntropy,
             "Symmetrical Cross-Entropy": SymmetricalSoftCrossEntropy, "Hellinger": Hellinger,
             "JSD": JSD, "L1": nn.L1Loss, "L2": nn.MSELoss}
    if name not in table:
        raise NotImplementedError(
            f"Loss type {name!r} is not available in this build; valid: {sorted(table)}")
    return table[name]()

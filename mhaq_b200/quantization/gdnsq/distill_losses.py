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


def get_distillation_loss(name: str):
    table = {"Symmetrical KL": SymmetricalKL, "KL": KL, "Cross-Entropy": SoftCrossEntropy,
             "L1": nn.L1Loss, "L2": nn.MSELoss}
    if name not in table:
        raise NotImplementedError(
            f"Loss type {name!r} is not available in this build; valid: {sorted(table)}")
    return table[name]()

"""Forward-hook registration helpers — mirror of the reference's
src/quantization/gdnsq/calib/hooks.py:6-22."""
import torch.nn as nn

from ..layers.gdnsq_act import NoisyAct


def forward_hook_register(module: nn.Module, hook):
    for name, child in module.named_children():
        child.register_forward_hook(hook(name))


def pre_forward_hook_register(module: nn.Module, hook):
    for name, child in module.named_children():
        child.register_forward_pre_hook(hook(name))


def register_lightning_activation_forward_hook(module, hook):
    """Attach `hook` to every NoisyAct; returns the handles (hooks.py:16-22)."""
    return [m.register_forward_hook(hook=hook) for _, m in module.named_modules()
            if isinstance(m, NoisyAct)]

"""Drive the LIVE reference (oracle/_ref, via oracle/ref_loader) the way its own pipeline does:
``Quantizer(config)().quantize(lmodel)`` -> ``Trainer.calibrate`` -> patched ``training_step``
(scripts/gdnsq_q_config.py:37-59).  TEST / BENCH INFRASTRUCTURE ONLY.

Lightning is absent from the image, so the orchestration lines of ``Trainer.calibrate``
(training/trainer.py:187-223) and of one ``fit`` iteration (forward -> loss -> backward ->
optimizer step) are restated here; every quantization-related call is the reference's own
function, imported unmodified from the staged tree.

``swapped_layers`` performs INTEGRATION.md §B in memory: the reference's modules keep their own
code but see this repo's ``NoisyAct / NoisyConv2d / NoisyLinear`` instead of theirs.
"""
from __future__ import annotations

import contextlib
import types

import torch
from torch import nn

from . import ref_loader


def make_cfg(ref, yaml_name=None, **over):
    """`config` as GDNSQQuant reads it.  From one of the reference's YAML files (quantization
    section, qscheme coerced to the reference's enum like its pydantic schema does,
    config_schema.py:47) or from keyword overrides."""
    if yaml_name is not None:
        cfg = ref_loader.load_yaml_quantization(yaml_name)
        q = cfg.quantization
    else:
        q = types.SimpleNamespace(name="GDNSQQuant", qscheme=1, act_bit=4, weight_bit=4,
                                  freeze_batchnorm=False, fuse_batchnorm=False, quantize_bias=False,
                                  excluded_layers=[], calibration=None,
                                  params=types.SimpleNamespace(distillation=False,
                                                               distillation_loss="Cross-Entropy",
                                                               distillation_teacher=None, qnmethod="STE"))
        cfg = types.SimpleNamespace(quantization=q)
    for k, v in over.items():
        if hasattr(q.params, k):
            setattr(q.params, k, v)
        else:
            setattr(q, k, v)
    if not isinstance(q.qscheme, ref.QScheme):
        q.qscheme = ref.QScheme(int(q.qscheme)) if not isinstance(q.qscheme, str) else ref.QScheme[q.qscheme]
    return cfg


def build_lmodule(ref, model: nn.Module, num_classes: int, lr: float = 3e-4,
                  criterion=None, optimizer=torch.optim.RAdam):
    """The reference's LVisionCls around `model` (models/compose/vision/vision_cls_module.py)."""
    setup = {"model": model, "criterion": criterion or nn.CrossEntropyLoss(), "optimizer": optimizer,
             "lr": lr, "config": types.SimpleNamespace(model=types.SimpleNamespace(params={"num_classes": num_classes}))}
    return ref.LVisionCls(setup)


def quantize(ref, lmodule, cfg):
    """scripts/gdnsq_q_config.py:44,49 — the plugin entry point, unmodified."""
    return ref.QuantizerFactory(cfg)().quantize(lmodule, in_place=True)


@contextlib.contextmanager
def swapped_layers(ref, NoisyAct, NoisyConv2d, NoisyLinear):
    """replace the three layer classes in every reference module that
    imports them (gdnsq_quant.py:9-11, model_helper.py:6-8, model_stats.py:7-9,
    minmaxobserver.py:3-5, hooks.py:4)."""
    mods = [ref.gdnsq_quant, ref.model_helper, ref.model_stats, ref.minmaxobserver, ref.hooks]
    new = {"NoisyAct": NoisyAct, "NoisyConv2d": NoisyConv2d, "NoisyLinear": NoisyLinear}
    saved = [(m, n, getattr(m, n)) for m in mods for n in new if hasattr(m, n)]
    try:
        for m, n, _ in saved:
            setattr(m, n, new[n])
        yield
    finally:
        for m, n, old in saved:
            setattr(m, n, old)


def calibrate(ref, qmodel, batch, act_bits=10, weight_bits=10, device=None):
    """Trainer.calibrate (training/trainer.py:187-223) on one batch: the reference's own
    apply_quantile_weights_s, MinMaxObserver forward hooks and apply_mean_stats_activations."""
    mm = ref.minmaxobserver
    if weight_bits:
        mm.apply_quantile_weights_s(qmodel.model, wbits=weight_bits)
    if act_bits:
        obs = mm.MinMaxObserver.__new__(mm.MinMaxObserver)   # (__init__ allocates on "cuda" unconditionally)
        handlers = ref.hooks.register_lightning_activation_forward_hook(qmodel.model, obs)
        was = qmodel.training
        qmodel.eval()
        with torch.no_grad():
            qmodel.model(batch)
        qmodel.train(was)
        for h in handlers:
            h.remove()
        mm.apply_mean_stats_activations(qmodel.model, abits=act_bits)
    if device is not None:      # the reference re-creates the activation Parameters on the CPU
        qmodel.to(device)
    return qmodel


def train_steps(qmodel, batch, n_steps, opt=None, on_step=None):
    """n_steps x (patched training_step -> backward -> optimizer step); returns (losses, opt)."""
    opt = opt or qmodel.configure_optimizers()
    qmodel.train()
    if hasattr(qmodel, "wrapped_criterion"):
        qmodel.wrapped_criterion.train()
    losses = []
    for i in range(n_steps):
        loss = qmodel.training_step(batch, i)
        loss.backward()
        if on_step is not None:
            on_step(i, qmodel)
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(loss.detach())
    return losses, opt

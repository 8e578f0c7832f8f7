"""Timing legs that execute the REFERENCE implementation (bench.py's `--impl reference` arm, its
`cpu_baseline` leg and the `reference_eager_gpu` denominators).  TEST / BENCH INFRASTRUCTURE
ONLY: nothing here is on the product path, and nothing here imports `mhaq_b200`.

`kind` says what ran:
  "reference" — the LIVE reference, unmodified, imported from the staged copy `oracle/_ref/`
                (oracle/make_ref.py; class `Quantizer`, gdnsq.py:159-241, and for the QAT legs
                the whole `Quantizer(config)().quantize()` pipeline);
  "port"      — the op-for-op restatement `oracle/fq_oracle.py`, used only when the staged copy
                is missing.
"""
from __future__ import annotations

import logging
import math
import os
import time
import types

import torch

from . import ref_loader


def backend():
    """-> (kind, namespace)."""
    if ref_loader.available():
        return "reference", ref_loader.load_ops()
    from . import fq_oracle
    return "port", fq_oracle


def make_step(kind, ns, x, go, scale, zp, lo, hi, method):
    """One fake-quant forward + backward through the reference operator on x's device:
    Q.dequantize(Q.quantize(x)) exactly as the layer wrappers call it (gdnsq_act.py:50-55,
    gdnsq_conv2d.py:98), gradient w.r.t. the input and the scale."""
    lo_ = -math.inf if lo is None else lo
    hi_ = math.inf if hi is None else hi
    sp = scale.clone().requires_grad_(True)
    if kind == "reference":
        mod = types.SimpleNamespace(training=True)
        Q = ns.Quantizer(mod, sp, zp, lo_, hi_, qnmethod=ns.QNMethod[method])

        def step():
            xs = x.detach().requires_grad_(True)
            sp.grad = None
            y = Q.dequantize(Q.quantize(xs))
            y.backward(go)
            return y, xs.grad, sp.grad
    else:
        def step():
            xs = x.detach().requires_grad_(True)
            sp.grad = None
            y = ns.fake_quant(xs, sp, zp, lo_, hi_, method=method)
            y.backward(go)
            return y, xs.grad, sp.grad
    return step


def cpu_microbench(inputs, method, steps, warmup, log2n, full_log2n):
    """The reference's CPU quantizer path on all host cores, on a bounded sample (N = 2^log2n
    elements of the same layout) of the bench workload."""
    kind, ns = backend()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, go, scale, zp, lo, hi = inputs
    step = make_step(kind, ns, x, go, scale, zp, lo, hi, method)
    n = x.numel()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    what = ("the live reference's Quantizer.quantize/dequantize + QN*.backward (oracle/_ref, unmodified)"
            if kind == "reference" else "oracle port of the reference's ATen op sequence (oracle/_ref not staged)")
    return {"value": round(20 * n / dt / 1e9, 3), "unit": "GB/s", "cores": cores, "kind": kind,
            "ms_per_step": round(dt * 1e3, 2),
            "sample": f"same workload cut to N=2^{log2n} of 2^{full_log2n} elements per step, {warmup} warm-up + "
                      f"{steps} timed steps, {what}, torch {torch.__version__} CPU, {cores} threads"}


def _quiet():
    logging.getLogger().setLevel(logging.WARNING)
    for name in list(logging.root.manager.loggerDict):
        if name.startswith(("src", "lightning")):
            logging.getLogger(name).setLevel(logging.WARNING)


def config0_step(steps=3, warmup=2):
    """BASELINE configs[0]: ResNet-20 CIFAR-10 GDNSQ(STE) W4A4 QAT step on the CPU, synthetic
    32x32 batch of 128 — through the LIVE reference: its in-tree `resnet20_cifar10`, its
    LVisionCls, `Quantizer(config)().quantize()`, `Trainer.calibrate`'s functions, the patched
    `training_step`, RAdam.  (Lightning itself is stubbed: SURVEY.md §8c.)"""
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref not staged"}
    from . import ref_harness as H
    ref = ref_loader.load_full()
    _quiet()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.randn(128, 3, 32, 32)
    t = torch.randint(0, 10, (128,))
    cfg = H.make_cfg(ref, act_bit=4, weight_bit=4, qscheme=1, qnmethod="STE",
                     excluded_layers=["conv1", "linear"])
    lm = H.build_lmodule(ref, ref.resnet_cifar.resnet20_cifar10(num_classes=10), 10)
    q = H.quantize(ref, lm, cfg)
    H.calibrate(ref, q, x[:32])
    opt, ts = None, []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, opt = H.train_steps(q, (x, t), 1, opt)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[warmup:])
    med = ts[len(ts) // 2]
    return {"workload": "configs[0] ResNet-20 CIFAR-10 GDNSQ(STE) W4A4 QAT step, batch 128, CPU, through the live "
                        "reference (in-tree resnet20_cifar10, LVisionCls, GDNSQQuant.quantize, patched training_step, RAdam)",
            "kind": "reference", "s_per_step": round(med, 3), "img_per_s": round(128 / med, 1),
            "threads": torch.get_num_threads(), "protocol": f"{warmup} warm-up + median of {steps}"}


def _time_gpu(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def eager_gpu_microbench(inputs, method, log2n):
    """The reference operator run eagerly on the GPU the bench runs on (its ATen launches)."""
    kind, ns = backend()
    x, go, scale, zp, lo, hi = inputs
    step = make_step(kind, ns, x, go, scale, zp, lo, hi, method)
    for _ in range(3):
        step()
    ms = _time_gpu(step, 10)
    return {"value": round(20 * x.numel() / (ms * 1e-3) / 1e9, 1), "unit": "GB/s", "ms_per_step": round(ms, 3),
            "kind": kind, "what": f"{'live reference Quantizer' if kind == 'reference' else 'oracle port'} run eagerly "
                                  f"on this GPU, N=2^{log2n}"}


def eager_gpu_resnet18_step(batch, channels_last, steps=8, seed=1234):
    """BASELINE configs[3] through the LIVE reference on the GPU: torchvision ResNet-18, the
    reference's LVisionCls + GDNSQQuant (config/gdnsq_config_resnet18_imagenet_ste_w4a4.yaml,
    quantization section), its calibration functions, its distillation training step, RAdam
    lr 3e-4, fp32 + TF32 — the reference's own layers and autograd Functions, eagerly launched."""
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref not staged"}
    import torchvision
    from . import ref_harness as H
    ref = ref_loader.load_full()
    _quiet()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.backends.cudnn.benchmark = True
    torch.set_float32_matmul_precision("high")
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(batch, 3, 224, 224, device=dev, generator=g)
    t = torch.randint(0, 1000, (batch,), device=dev, generator=g)
    cfg = H.make_cfg(ref, "gdnsq_config_resnet18_imagenet_ste_w4a4.yaml")
    model = torchvision.models.resnet18(num_classes=1000).to(dev)
    lm = H.build_lmodule(ref, model, 1000, lr=3e-4).to(dev)
    q = H.quantize(ref, lm, cfg).to(dev)
    H.calibrate(ref, q, x[: min(batch, 64)], device=dev)
    if channels_last:
        q.model.to(memory_format=torch.channels_last)
        q.tmodel.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    opt = q.configure_optimizers()
    q.train(); q.wrapped_criterion.train(); q.tmodel.eval()

    def step():
        loss = q.training_step((x, t), 0)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(3):
        step()
    ms = _time_gpu(step, steps)
    del q, lm, model, opt
    torch.cuda.empty_cache()
    return {"img_per_s": round(batch / (ms * 1e-3), 1), "ms_per_step": round(ms, 2), "kind": "reference",
            "what": "the live reference's whole QAT step (its layers, autograd Functions, ModelHelper, PotentialLoss) "
                    "launched eagerly on this GPU, " + ("channels_last like our leg" if channels_last else
                                                        "row-major NCHW as the reference's Trainer runs it")}

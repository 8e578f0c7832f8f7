#!/usr/bin/env python
"""Which Python call sites launch the small kernels of one eager QAT step (torch.profiler with
stacks): ATen ops that launch kernels, grouped by (op, input shapes, innermost repo / torch.nn frame).

    python tools/launch_sources.py [--model resnet20] [--method STE]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="resnet20")
ap.add_argument("--method", default="STE")
ap.add_argument("--bits", type=int, default=4)
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
from mhaq_b200 import harness  # noqa: E402
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
side, classes = (224, 1000) if a.model == "resnet18" else (32, 100)
x = torch.randn(a.batch, 3, side, side, device=dev).contiguous(memory_format=torch.channels_last)
t = torch.randint(0, classes, (a.batch,), device=dev)
q = harness.build_qat(a.model, dev, qnmethod=a.method, act_bit=a.bits, weight_bit=a.bits,
                      distillation=True, num_classes=classes, calib_batch=x[:64])
q.model.to(memory_format=torch.channels_last)
if getattr(q, "tmodel", None) is not None:
    q.tmodel.to(memory_format=torch.channels_last)
    q.tmodel.eval()
opt = q.configure_optimizers()
q.train(); q.wrapped_criterion.train()


def step():
    loss = q.training_step((x, t), 0)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = collections.Counter()
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU or not ev.kernels:
        continue
    # only leaf ops (an op whose child also owns the kernels would double count)
    if any(ch.kernels for ch in ev.cpu_children):
        continue
    frames = [f for f in (ev.stack or []) if "/repo/" in f or "torch/nn/" in f or "torch/optim" in f]
    where = frames[0] if frames else ((ev.stack or ["<autograd engine / no python frame>"])[0])
    where = where.replace("/root/repo/", "").split("/site-packages/")[-1]
    shapes = str(ev.input_shapes)[:70]
    rows[(ev.name, len(ev.kernels), shapes, where[:110])] += 1
print(f"# {a.model} {a.method}: ops that launch kernels in ONE eager step, by call site (count >= 2 shown)")
tot = 0
for (name, nk, shapes, where), c in sorted(rows.items(), key=lambda kv: -kv[1] * kv[0][1]):
    tot += c * nk
    if c >= 2:
        print(f"{c * nk:4d} launches  {name:38s} {shapes:70s} {where}")
print("# total launches attributed:", tot)

#!/usr/bin/env python
"""Where the GPU time of one QAT step goes: top kernels by device time (CUPTI via torch.profiler).

    python tools/step_breakdown.py [--model resnet18] [--batch 256] [--channels-last] [--top 25]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--method", default="STE")
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--by-count", action="store_true", help="sort by launches per step instead of time")
    a = ap.parse_args()
    from mhaq_b200 import harness
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.set_float32_matmul_precision("high")
    if a.model == "rfdn":
        x = torch.rand(a.batch, 3, 256, 256, device=dev)
        t = torch.rand(a.batch, 3, 1024, 1024, device=dev)
        q = harness.build_qat("rfdn", dev, qnmethod=a.method, act_bit=a.bits, weight_bit=a.bits,
                              distillation=False, lr=5e-4, calib_batch=x[:4], calib_bits=a.bits)
    else:
        side, classes = (224, 1000) if a.model == "resnet18" else (32, 100)
        x = torch.randn(a.batch, 3, side, side, device=dev)
        t = torch.randint(0, classes, (a.batch,), device=dev)
        q = harness.build_qat(a.model, dev, qnmethod=a.method, act_bit=a.bits, weight_bit=a.bits,
                              distillation=True, num_classes=classes, calib_batch=x[:64])
    if a.channels_last:
        q.model.to(memory_format=torch.channels_last)
        if getattr(q, "tmodel", None) is not None:
            q.tmodel.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    opt = q.configure_optimizers()
    q.train(); q.wrapped_criterion.train()
    if getattr(q, "tmodel", None) is not None:
        q.tmodel.eval()

    def step():
        loss = q.training_step((x, t), 0)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"# {a.model} batch {a.batch} {'channels_last' if a.channels_last else 'NCHW'}: "
          f"{e0.elapsed_time(e1) / 6:.2f} ms/step")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        dt = getattr(ev, "device_time_total", 0.0) or 0.0
        if dt > 0:
            rows.append((dt / 2e3, ev.count // 2, ev.key))
    rows.sort(reverse=True, key=(lambda r: (r[1], r[0])) if a.by_count else None)
    tot = sum(r[0] for r in rows)
    print(f"# GPU kernel time {tot:.2f} ms/step over {sum(r[1] for r in rows)} launches/step")
    for ms, n, k in rows[: a.top]:
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n:<4d} {k[:110]}")


if __name__ == "__main__":
    main()

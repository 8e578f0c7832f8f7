#!/usr/bin/env python
"""Small driver for ncu: the row-resident weight kernels at ResNet-18 weight shapes and the AEWGS
statistics kernel at the microbench shape.  Prints CUDA-event timings when run without ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import mhaq_b200  # noqa: E402
from mhaq_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


for shape in ((512, 512, 3, 3), (256, 256, 3, 3), (64, 64, 3, 3)):
    w = torch.randn(shape, device=dev, requires_grad=True)
    ls = torch.full((shape[0], 1, 1, 1), -6.0, device=dev, requires_grad=True)
    go = torch.randn(shape, device=dev)
    gl = torch.randn(shape[0], device=dev)

    def fused():
        wq, _, _, lr = ops.weight_fake_quant_rows(w, ls, method="STE")
        torch.autograd.backward([wq, lr], [go, gl])

    def streaming():
        wq, mn, mx = ops.weight_fake_quant_log(w, ls, method="STE")
        lr = torch.log2(mx - mn + torch.exp2(ls.ravel()))
        torch.autograd.backward([wq, lr], [go, gl])

    print(f"weight {shape}: fused rows {timed(fused):7.1f} us   streaming + torch range term {timed(streaming):7.1f} us "
          f"(host-launch bound: fwd+bwd per call, back to back)", flush=True)

x = torch.randn(512, 1 << 19, device=dev)
go = torch.randn(512, 1 << 19, device=dev)
mn, mx = x.amin(1, keepdim=True), x.amax(1, keepdim=True)
scale = (mx - mn) / 15
L = ops._Launch(x, scale, mn, None, None)
t = timed(lambda: ops.aewgs_stats(go, x, L, False), 20)
print(f"aewgs stats [512, 2^19]: {t:7.1f} us = {8 * x.numel() / t / 1e3:.0f} GB/s (8 B/elem)")

"""Separate 'clamp' from 'layout' effects in the backward kernel timing (B200)."""
import math, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops
dev = torch.device("cuda")
n = 1 << 28
x = torch.randn(n, device=dev); go = torch.randn(n, device=dev)
def t(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
for rows in (1, 512):
    xs = x.view(rows, -1); gs = go.view(rows, -1)
    for clamp in (False, True):
        for method in ("STE", "LSQ"):
            if rows == 1:
                s = torch.tensor([0.25], device=dev); zp = torch.tensor([-2.0], device=dev)
                lo = zp if clamp else -math.inf; hi = (zp + 4.0 - s) if clamp else math.inf
            else:
                s = torch.full((rows, 1), 0.25, device=dev); zp = torch.full((rows, 1), -2.0, device=dev)
                lo = zp.clone() if clamp else -math.inf; hi = (zp + 4.0 - s) if clamp else math.inf
            L = ops._Launch(xs, s, zp, lo, hi)
            mid = ops._method_id(method)
            ms = t(lambda: ops._backward_impl(gs, xs, L, mid, False, None, True, philox=(1, 2)))
            msf = t(lambda: ops._forward_impl(xs, L, True, False, False))
            print(f"rows={rows:4d} clamp={clamp!s:5} {method}: bwd {12*n/ms/1e6:7.1f} GB/s ({ms:.4f} ms)  fwd {8*n/msf/1e6:7.1f} GB/s")

"""What a per-channel flat kernel could reach: the flat backward WITHOUT clamp at 2^28 (per-tensor,
lo/hi = -/+inf) next to the clamped flat backward and the unclamped per-channel streaming kernel
([512, 2^19], the bench workload).  CUDA-event timing over 20 back-to-back calls on rotating inputs."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops

n = 1 << 28
xs = [torch.randn(n, device="cuda") for _ in range(2)]
gs = [torch.randn(n, device="cuda") for _ in range(2)]
b = torch.tensor([-2.0], device="cuda"); s = torch.tensor([0.25], device="cuda"); hi = b + 4.0 - s


def timeit(f, reps=20):
    for i in range(4): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


Lc = [ops._Launch(x, s, b, b, hi) for x in xs]
Lu = [ops._Launch(x, s, b, -math.inf, math.inf) for x in xs]
t = timeit(lambda i: ops._backward_impl(gs[i % 2], xs[i % 2], Lc[i % 2], 0, False, None, True, philox=(1, 2)))
print(f"flat clamped   per-tensor 2^28: {t:.1f} us  {12*n/t/1e3:.0f} GB/s")
t = timeit(lambda i: ops._backward_impl(gs[i % 2], xs[i % 2], Lu[i % 2], 0, False, None, True, philox=(1, 2)))
print(f"flat unclamped per-tensor 2^28: {t:.1f} us  {12*n/t/1e3:.0f} GB/s")
x2 = [x.view(512, 1 << 19) for x in xs]; g2 = [g.view(512, 1 << 19) for g in gs]
sc = torch.full((512, 1), 0.25, device="cuda"); zp = torch.full((512, 1), -2.0, device="cuda")
Lp = [ops._Launch(x, sc, zp, -math.inf, math.inf) for x in x2]
t = timeit(lambda i: ops._backward_impl(g2[i % 2], x2[i % 2], Lp[i % 2], 0, False, None, True, philox=(1, 2)))
print(f"streaming unclamped [512,2^19] (+finalize): {t:.1f} us  {12*n/t/1e3:.0f} GB/s")

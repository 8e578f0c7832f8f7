"""GPU-side fwd/bwd time at the ResNet-18 activation sizes (per-tensor, clamped, STE).
Cold-cache without a dirty flush: K rotating input sets whose footprint exceeds 2x the L2."""
import math, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops
dev = torch.device("cuda")
def timed(body, reps):
    for i in range(4): body(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): body(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (1 << 20, 1 << 22, 6422528, 12845056, 1 << 24, 25690112, 51380224, 1 << 26):
    K = max(2, int(math.ceil(2 * 126e6 / (4 * n))) + 1)
    xs = [torch.randn(n, device=dev) for _ in range(K)]; gs = [torch.randn(n, device=dev) for _ in range(K)]
    s = torch.tensor([0.25], device=dev); zp = torch.tensor([-2.0], device=dev)
    Ls = [ops._Launch(x, s, zp, zp, zp + 4.0 - s) for x in xs]
    f = lambda i: ops._forward_impl(xs[i % K], Ls[i % K], True, False, False)
    b = lambda i: ops._backward_impl(gs[i % K], xs[i % K], Ls[i % K], 0, False, None, True, philox=(1, 2))
    reps = 60
    mf = timed(f, reps); mb = timed(b, reps)
    print(f"n={n:9d} K={K:3d} fwd {mf*1e3:7.1f} us {8*n/mf/1e6:7.0f} GB/s ({8*n/mf/1e6/6540.8:.2f})   bwd {mb*1e3:7.1f} us {12*n/mb/1e6:7.0f} GB/s ({12*n/mb/1e6/6540.8:.2f})")

"""Per-tensor UNCLAMPED STE backward (lo/hi = -/+inf; per-tensor weights): flat single launch vs
streaming + finalize (run with MHAQ_FQ_FLAT_MAX_LOG2=0), back-to-back calls on rotating inputs."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops
b = torch.tensor([-2.0], device="cuda"); s = torch.tensor([0.25], device="cuda")
for log2n in (22, 24, 25, 26, 27, 28):
    n = 1 << log2n
    k = max(2, (1 << 28) // n)
    k = min(k, 16)
    xs = [torch.randn(n, device="cuda") for _ in range(k)]
    gs = [torch.randn(n, device="cuda") for _ in range(k)]
    Ls = [ops._Launch(x, s, b, -math.inf, math.inf) for x in xs]
    f = lambda i: ops._backward_impl(gs[i % k], xs[i % k], Ls[i % k], 0, False, None, True, philox=(1, 2))
    for i in range(4): f(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    reps = 2 * k
    with torch.cuda.graph(g):
        for i in range(reps): f(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e3
    print(f"2^{log2n}: {t:.1f} us  {12*n/t/1e3:.0f} GB/s  single_launch={ops.lib.mhaq_fq_bwd_single_launch(1, n, 1, 0, 0)}")
    del g, xs, gs, Ls
    torch.cuda.empty_cache()

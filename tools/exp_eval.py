"""Eval-statistics forward (y + min/max code + min/max input, two launches) vs the plain forward and
copy_ at 2^24..2^28 per-tensor clamped, CUDA-graph replays on rotating inputs."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops
from tools.midsize_graph import graph_time, PEAK
dev = torch.device("cuda")
for log2n in (24, 26, 28):
    n = 1 << log2n
    K = min(16, max(3, int(math.ceil(2.2 * 126e6 / (4 * n))) + 1))
    xs = [torch.randn(n, device=dev) for _ in range(K)]
    outs = [torch.empty(n, device=dev) for _ in range(K)]
    s = torch.tensor([0.25], device=dev); zp = torch.tensor([-2.0], device=dev)
    Ls = [ops._Launch(x, s, zp, zp, zp + 4.0 - s) for x in xs]
    tf = graph_time(lambda i: ops._forward_impl(xs[i], Ls[i], True, False, False), K)
    te = graph_time(lambda i: ops._forward_impl(xs[i], Ls[i], True, False, True), K)
    tc = graph_time(lambda i: outs[i].copy_(xs[i]), K)
    print(f"2^{log2n}: fwd {tf*1e3:7.1f} us ({8*n/tf/1e6/PEAK:.3f})  eval {te*1e3:7.1f} us ({8*n/te/1e6/PEAK:.3f})  copy {tc*1e3:7.1f} us", flush=True)

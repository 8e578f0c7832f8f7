import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops
dev = torch.device("cuda")
for n in (1 << 24, 25690112, 1 << 26):
    xs = [torch.randn(n, device=dev) for _ in range(4)]; gs = [torch.randn(n, device=dev) for _ in range(4)]
    s = torch.tensor([0.25], device=dev); zp = torch.tensor([-2.0], device=dev)
    for i in range(8):
        L = ops._Launch(xs[i % 4], s, zp, zp, zp + 4.0 - s)
        ops._forward_impl(xs[i % 4], L, True, False, False)
        ops._backward_impl(gs[i % 4], xs[i % 4], L, 0, False, None, True, philox=(1, 2))
    torch.cuda.synchronize()

"""Host cost of one fake_quant forward / backward call (small tensor: GPU time negligible)."""
import cProfile, pstats, io, sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhaq_b200
from mhaq_b200 import ops
dev = torch.device("cuda")
x = torch.randn(64, 64, 8, 8, device=dev, requires_grad=True); go = torch.randn_like(x)
ls = torch.tensor([-2.0], device=dev, requires_grad=True); lq = torch.tensor([2.0], device=dev, requires_grad=True)
b = torch.tensor([-2.0], device=dev, requires_grad=True)
def fwd_bwd():
    s = torch.exp2(ls); q = torch.exp2(lq)
    y = mhaq_b200.fake_quant(x, s, b, b, b + q - s, method="STE")
    y.backward(go)
def raw():
    L = ops._Launch(x.detach(), sd, b.detach(), b.detach(), hid)
    ops._forward_impl(x.detach(), L, True, False, False)
    ops._backward_impl(go, x.detach(), L, 0, False, None, True, philox=(1, 2))
sd = torch.exp2(ls).detach(); hid = (b + torch.exp2(lq) - sd).detach()
for f in (fwd_bwd, raw):
    for _ in range(20): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): f()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{f.__name__}: host {1e6*(t1-t0)/200:.1f} us per call pair")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): raw()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])

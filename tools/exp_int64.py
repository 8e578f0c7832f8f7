"""Indexing beyond 2^31 elements: per-tensor clamped LSQ forward + backward (the single-launch flat
backward) and per-channel STE at n = 2^31 + 4100; slices of the big run are compared bit for bit
with separate runs on copies of those slices (LSQ: y and gx depend on the element's value only)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhaq_b200 as fq

n = (1 << 31) + 4100
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.empty(n, device="cuda").normal_(generator=g)
go = torch.empty(n, device="cuda").normal_(generator=g)
b = torch.tensor([-2.0], device="cuda"); s = torch.tensor([0.25], device="cuda"); hi = b + 4.0 - s


def run(xx, gg):
    xs = xx.clone().requires_grad_(True)
    s_, b_, h_ = (t.clone().requires_grad_(True) for t in (s, b, hi))
    y = fq.fake_quant(xs, s_, b_, b_, h_, method="LSQ")
    y.backward(gg)
    return y.detach(), xs.grad, s_.grad.double(), b_.grad.double(), h_.grad.double()


y, gx, gs, gb, gh = run(x, go)
torch.cuda.synchronize()
ok = True
sums = [0.0, 0.0, 0.0]
step = 1 << 28
for lo_ in range(0, n, step):
    hi_ = min(lo_ + step, n)
    ys, gxs, a, c, d = run(x[lo_:hi_], go[lo_:hi_])
    e1, e2 = torch.equal(ys, y[lo_:hi_]), torch.equal(gxs, gx[lo_:hi_])
    ok &= e1 and e2
    sums[0] += float(a); sums[1] += float(c); sums[2] += float(d)
    print(f"slice [{lo_}, {hi_}): y {'==' if e1 else '!='}  gx {'==' if e2 else '!='}", flush=True)
for name, big, parts in (("g_scale", gs, sums[0]), ("g_zp+g_lo", gb, sums[1]), ("g_hi", gh, sums[2])):
    rel = abs(float(big) - parts) / max(abs(parts), 1e-30)
    print(f"{name}: whole {float(big):.9g}  sum of slices {parts:.9g}  rel {rel:.2e}")
    ok &= rel < 1e-5
print("INT64 INDEXING", "OK" if ok else "FAILED", f"(n = {n}, single_launch = {fq.ops.lib.mhaq_fq_bwd_single_launch(1, n, 1, 3, 0)})")

# ---- per-channel [4100, 2^19] (2.15e9 elements) through the streaming kernels, LSQ
del x, go, y, gx
torch.cuda.empty_cache()
R, I = 4100, 1 << 19
x = torch.empty(R, I, device="cuda").normal_(generator=g)
go = torch.empty(R, I, device="cuda").normal_(generator=g)
sc = torch.full((R, 1), 0.25, device="cuda") * (1 + torch.arange(R, device="cuda").view(R, 1) % 3)
zp = torch.full((R, 1), -2.0, device="cuda")


def run_pc(xx, gg, s0, z0):
    xs = xx.clone().requires_grad_(True)
    s_, z_ = s0.clone().requires_grad_(True), z0.clone().requires_grad_(True)
    y = fq.fake_quant(xs, s_, z_, -math.inf, math.inf, method="LSQ")
    y.backward(gg)
    return y.detach(), xs.grad, s_.grad, z_.grad


y, gx, gs, gz = run_pc(x, go, sc, zp)
ok2 = True
for r0 in range(0, R, 1024):
    # (always 1024 rows, the last block overlapping: the task policy — 1 or 2 sub-tiles per task,
    # hence the fp32 grouping inside g_scale — depends on the number of tasks of the call)
    r0 = min(r0, R - 1024)
    r1 = r0 + 1024
    ys, gxs, a, c = run_pc(x[r0:r1], go[r0:r1], sc[r0:r1], zp[r0:r1])
    e = [torch.equal(ys, y[r0:r1]), torch.equal(gxs, gx[r0:r1]), torch.equal(a, gs[r0:r1]), torch.equal(c, gz[r0:r1])]
    ok2 &= all(e)
    print(f"rows [{r0}, {r1}): y, gx, g_scale, g_zp equal: {e}", flush=True)
print("INT64 INDEXING per-channel", "OK" if ok2 else "FAILED", f"({R} x {I} = {R * I} elements)")

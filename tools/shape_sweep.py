"""Effective bandwidth of the fake-quant calls at the real tensor shapes of BASELINE configs 3-5
(SURVEY.md Appendix B): ResNet-18 @B=256, ResNet-20 @B=256, RFDN @B=4 (256x256 LR patches).
GPU-side time (back-to-back launches, rotating cold inputs), through the public autograd API's
launch path.  Prints one line per distinct shape and the per-step totals."""
import math, sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops
dev = torch.device("cuda")
PEAK = 6540.8

def timed(body, reps=40):
    for i in range(4): body(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): body(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def act(shape, method="STE"):
    n = math.prod(shape)
    K = max(2, int(math.ceil(2 * 126e6 / (4 * n))) + 1)
    K = min(K, 64)
    xs = [torch.randn(shape, device=dev) for _ in range(K)]; gs = [torch.randn(shape, device=dev) for _ in range(K)]
    s = torch.tensor([0.25], device=dev); zp = torch.tensor([-2.0], device=dev)
    Ls = [ops._Launch(x, s, zp, zp, zp + 4.0 - s) for x in xs]
    mid = ops._method_id(method)
    tf = timed(lambda i: ops._forward_impl(xs[i % K], Ls[i % K], True, False, False))
    tb = timed(lambda i: ops._backward_impl(gs[i % K], xs[i % K], Ls[i % K], mid, False, None, True, philox=(1, 2)))
    return n, tf, tb

def weight(shape, method):
    n = math.prod(shape)
    w = torch.randn(shape, device=dev) * 0.1; go = torch.randn(shape, device=dev)
    s = torch.full((shape[0], 1, 1, 1), 0.01, device=dev)
    mid = ops._method_id(method)
    def f(i):
        return ops._WeightFakeQuantFn.apply(w, s, mid, None, (1, 2))
    wr = w.clone().requires_grad_(True)
    def fb(i):
        wq, mn, mx = ops._WeightFakeQuantFn.apply(wr, s, mid, None, (1, 2))
        wq.backward(go); wr.grad = None
    tf = timed(f); tfb = timed(fb)
    return n, tf, tfb - tf

models = {
 "ResNet-18 @B=256 (STE W4A4)": dict(B=256, acts=[((64,56,56),5),((128,28,28),4),((256,14,14),4),((512,7,7),3)], wmethod="STE",
     weights=[((64,64,3,3),4),((128,64,3,3),1),((128,128,3,3),3),((256,128,3,3),1),((256,256,3,3),3),((512,256,3,3),1),((512,512,3,3),3)]),
 "ResNet-20 @B=256 (AEWGS W1A1)": dict(B=256, acts=[((16,32,32),7),((32,16,16),6),((64,8,8),5)], wmethod="AEWGS",
     weights=[((16,16,3,3),6),((32,16,3,3),1),((32,32,3,3),5),((64,32,3,3),1),((64,64,3,3),5)]),
 "RFDN @B=4, 256x256 LR (LSQ W2A2)": dict(B=4, acts=[((50,256,256),17),((12,256,256),4),((12,41,41),12)], wmethod="LSQ",
     weights=[((50,50,3,3),13),((25,50,3,3),4),((12,50,3,3),4),((12,12,3,3),12)]),
}
# NOTE: ResNet-18 activation inputs of the quantized 3x3 convs: the input of a
# stride-2 conv has the previous stage's resolution; counts follow the appendix totals approximately.
out = {}
for name, m in models.items():
    print(f"== {name}")
    tot_alg = tot_t = 0.0
    for shp, cnt in m["acts"]:
        n, tf, tb = act((m["B"],) + shp)
        print(f"  act  {str((m['B'],)+shp):24s} x{cnt:2d}  {n/1e6:7.2f} M elems  fwd {tf*1e3:7.1f} us ({8*n/tf/1e6/PEAK:.2f})  bwd {tb*1e3:7.1f} us ({12*n/tb/1e6/PEAK:.2f})")
        tot_alg += cnt * 20 * n; tot_t += cnt * (tf + tb)
    for shp, cnt in m["weights"]:
        n, tf, tb = weight(shp, m["wmethod"])
        print(f"  wght {str(shp):24s} x{cnt:2d}  {n/1e6:7.3f} M elems  fwd {tf*1e3:7.1f} us  bwd {tb*1e3:7.1f} us   (launch-bound)")
        tot_alg += cnt * 20 * n; tot_t += cnt * (tf + tb)
    print(f"  per step: {tot_alg/1e9:.2f} GB algorithmic in {tot_t:.3f} ms  ->  {tot_alg/tot_t/1e6:.0f} GB/s  ({tot_alg/tot_t/1e6/PEAK:.2f} of measured peak)")
    out[name] = {"GB_algorithmic": tot_alg / 1e9, "ms": tot_t, "GBps": tot_alg / tot_t / 1e6}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "shape_sweep.json"), "w"), indent=1)

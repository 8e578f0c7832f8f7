"""GPU-side time of the fake-quant calls at mid-size tensors, host launch cost excluded.

Each measurement captures K calls on K rotating input sets (footprint > 2x the 126 MB L2, so
every call sees cold inputs without a dirty flush) into ONE CUDA graph and times its replays:
what a graph-captured training step pays per quantizer.  The eager (Python-launched) time of
the same calls is printed beside it, and the size-matched ceilings: torch `copy_` (1R:1W,
8 B/element) for the forward and `torch.add(a, b, out=c)` (2R:1W, 12 B/element) for the
backward, timed the same way.

    python tools/midsize_graph.py [--quick] [--out gpurun_out/midsize.json]
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mhaq_b200 import ops  # noqa: E402

dev = torch.device("cuda")
PEAK = 6540.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def graph_time(body, K, reps=20):
    """body(i) for i in range(K) captured once; returns ms per call over `reps` replays."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(K):
            body(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(K):
            body(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * K)


def eager_time(body, K, reps=3):
    for i in range(K):
        body(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(K):
            body(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * K)


def case(name, shape, per_channel, method="STE", clamp=True):
    n = math.prod(shape)
    K = min(64, max(3, int(math.ceil(2.2 * 126e6 / (4 * n))) + 1))
    xs = [torch.randn(shape, device=dev) for _ in range(K)]
    gs = [torch.randn(shape, device=dev) for _ in range(K)]
    outs = [torch.empty(shape, device=dev) for _ in range(K)]
    if per_channel:
        C = shape[0]
        s = torch.full((C, 1), 0.05, device=dev)
        zp = torch.full((C, 1), -4.0, device=dev)
        Ls = [ops._Launch(x, s, zp, None, None) for x in xs]
    else:
        s = torch.tensor([0.25], device=dev)
        zp = torch.tensor([-2.0], device=dev)
        Ls = [ops._Launch(x, s, zp, zp, zp + 4.0 - s) if clamp else ops._Launch(x, s, zp, None, None) for x in xs]
    mid = ops._method_id(method)
    f = lambda i: ops._forward_impl(xs[i], Ls[i], True, False, False)
    b = lambda i: ops._backward_impl(gs[i], xs[i], Ls[i], mid, False, None, True, philox=(1, 2))
    ev = lambda i: ops._forward_impl(xs[i], Ls[i], True, False, True)     # eval: y + code / input min-max
    cp = lambda i: outs[i].copy_(xs[i])
    ad = lambda i: torch.add(xs[i], gs[i], out=outs[i])
    r = {"name": name, "shape": list(shape), "n": n, "K": K, "method": method}
    for key, fn, by in (("fwd", f, 8), ("eval", ev, 8), ("bwd", b, 12 + (8 if method == "AEWGS" else 0)), ("copy", cp, 8),
                        ("add3", ad, 12)):
        tg = graph_time(fn, K)
        te = eager_time(fn, K)
        r[key] = {"graph_us": round(tg * 1e3, 2), "eager_us": round(te * 1e3, 2),
                  "GBps": round(by * n / tg / 1e6, 1), "frac_peak": round(by * n / tg / 1e6 / PEAK, 3)}
    r["fwd_vs_copy"] = round(r["copy"]["graph_us"] / r["fwd"]["graph_us"], 3)
    r["bwd_vs_add3"] = round(r["add3"]["graph_us"] / r["bwd"]["graph_us"], 3)
    print(f"{name:34s} n={n/1e6:7.2f}M  fwd {r['fwd']['graph_us']:7.1f} us ({r['fwd']['frac_peak']:.2f}; copy {r['copy']['graph_us']:6.1f} us; eager {r['fwd']['eager_us']:6.1f})"
          f"  eval {r['eval']['graph_us']:7.1f} us ({r['eval']['frac_peak']:.2f})"
          f"  bwd {r['bwd']['graph_us']:7.1f} us ({r['bwd']['frac_peak']:.2f}; add3 {r['add3']['graph_us']:6.1f} us; eager {r['bwd']['eager_us']:6.1f})", flush=True)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--large", action="store_true", help="per-tensor clamped STE at 12.8 M .. 2^28 elements only")
    ap.add_argument("--out", default=os.path.join("gpurun_out", "midsize.json"))
    a = ap.parse_args()
    rows = []
    acts = [("resnet18 act (256,64,56,56)", (256, 64, 56, 56)), ("resnet18 act (256,128,28,28)", (256, 128, 28, 28)),
            ("resnet18 act (256,256,14,14)", (256, 256, 14, 14)), ("resnet18 act (256,512,7,7)", (256, 512, 7, 7)),
            ("resnet20 act (256,16,32,32)", (256, 16, 32, 32)), ("resnet20 act (256,32,16,16)", (256, 32, 16, 16)),
            ("resnet20 act (256,64,8,8)", (256, 64, 8, 8)), ("rfdn act (4,50,256,256)", (4, 50, 256, 256)),
            ("rfdn act (4,12,256,256)", (4, 12, 256, 256))]
    if a.quick:
        acts = acts[2:6]
    if a.large:
        acts = [("resnet18 act (256,256,14,14)", (256, 256, 14, 14)), ("resnet18 act (256,128,28,28)", (256, 128, 28, 28)),
                ("resnet18 act (256,64,56,56)", (256, 64, 56, 56)), ("2^26", (1 << 26,)), ("2^27", (1 << 27,)), ("2^28", (1 << 28,))]
    for nm, shp in acts:
        rows.append(case(nm, shp, False))
    if not a.quick and not a.large:
        for log2n in (20, 22, 24, 26):
            for C in (0, 64, 512, 4096):
                n = 1 << log2n
                shp = (n,) if C == 0 else (C, n // C)
                for m in ("STE", "LSQ") + (("AEWGS",) if C else ()):
                    rows.append(case(f"sweep 2^{log2n} C={C} {m}", shp, C != 0, m))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"peak_GBps": PEAK, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — fake-quant fwd+bwd throughput (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our sm_100a path
    python bench.py --impl reference [...]                        # reference CPU path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic input: one
fake-quant forward + one backward (incl. the in-kernel deterministic reduction) over a
[C, N/C] fp32 tensor.  Default workload = BASELINE.json configs[1] headline
point: per-channel (C=512) N=2^28 elements, GDNSQ/STE estimator with fused
Philox noise, 4 bits.  GB/s = 20 B/element (8 fwd + 12 bwd) x N / t.

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BYTES_FWD, BYTES_BWD = 8, 12          # algorithmic bytes / element (fp32)
METRIC = "fake-quant fwd+bwd GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--channels", type=int, default=512, help="0 = per-tensor activation style")
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--method", default="STE", choices=["STE", "LSQ", "AEWGS", "EWGS"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="also run the config-2 sweep -> gpurun_out/")
    ap.add_argument("--cpu-log2n", type=int, default=24)
    ap.add_argument("--no-resnet", action="store_true", help="skip the ResNet-18 W4A4 QAT leg")
    ap.add_argument("--resnet-batch", type=int, default=256, help="per-GPU batch (config 4)")
    ap.add_argument("--resnet-steps", type=int, default=12)
    ap.add_argument("--qat-model", default="resnet18", choices=["resnet18", "resnet20", "rfdn"],
                    help="resnet18 = configs[3] (ImageNet-shaped, STE W4A4); resnet20 = configs[2] "
                         "(CIFAR-100-shaped; use --qat-method AEWGS --qat-bits 1); rfdn = configs[4] "
                         "(x4 super-resolution on 256x256 patches, L1, no teacher; use --qat-method LSQ "
                         "--qat-bits 2 --resnet-batch 16)")
    ap.add_argument("--qat-method", default="STE", choices=["STE", "LSQ", "AEWGS", "EWGS"])
    ap.add_argument("--qat-bits", type=int, default=4)
    ap.add_argument("--ddp-reference-flags", action="store_true",
                    help="wrap DDP exactly like the reference's Trainer (find_unused_parameters=True, "
                         "buffer broadcast every step) instead of the lean wrapping")
    ap.add_argument("--graph", action="store_true", default=True, help=argparse.SUPPRESS)
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="QAT leg: launch the training step eagerly from Python.  Default: the whole step "
                         "(teacher+student forward, loss, backward, DDP all-reduces, RAdam) is captured "
                         "once in a CUDA graph and replayed")
    ap.add_argument("--nchw", dest="channels_last", action="store_false",
                    help="QAT leg: keep model and batch in row-major NCHW (the reference Trainer's layout). "
                         "Default is torch.channels_last (NHWC): cuDNN's Blackwell convolution / batch-norm "
                         "kernels are NHWC-native (NCHW costs ~8 ms of layout transposes and 2x slower BN per "
                         "step), and the fake-quant kernels walk dense NHWC storage directly, no copies")
    ap.add_argument("--no-eager-ref", action="store_true",
                    help="skip timing the reference's ATen chain on the GPU (second denominator)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(a):
    lay = f"per-channel [C={a.channels}, N/C]" if a.channels else "per-tensor"
    return (f"configs[1] quantizer microbench: {lay} fp32 N=2^{a.log2n}, {a.method}"
            f"{' (GDNSQ, fused Philox noise)' if a.method == 'STE' else ''}, {a.bits} bits, fwd+bwd")


def make_inputs(a, device, log2n=None):
    """x~N(0,1) seed 0, go~N(0,1) seed 1."""
    n = 1 << (log2n if log2n is not None else a.log2n)
    gx = torch.Generator(device=device).manual_seed(0)
    gg = torch.Generator(device=device).manual_seed(1)
    if a.channels:
        shape = (a.channels, n // a.channels)
    else:
        shape = (n,)
    x = torch.randn(shape, device=device, generator=gx)
    go = torch.randn(shape, device=device, generator=gg)
    if a.channels:
        mn = x.amin(1, keepdim=True)
        mx = x.amax(1, keepdim=True)
        scale = (mx - mn) / (2 ** a.bits - 1)        # calibration formula, minmaxobserver.py:82
        zp, lo, hi = mn, None, None
    else:
        b = torch.tensor([-2.0], device=device)
        scale = torch.exp2(torch.tensor([2.0 - a.bits], device=device))
        zp, lo, hi = b, b, b + 4.0 - scale
    return x, go, scale, zp, lo, hi


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def time_region(fn, steps, sync_dist):
    """K steps between barrier+synchronize on both sides, CUDA events on the launching stream."""
    import torch.distributed as dist
    torch.cuda.synchronize()
    if sync_dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if sync_dist:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if sync_dist:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# ---------------------------------------------------------------------------
class _EagerReferenceBackend:
    """Bench-only A/B switch: route Quantizer.fake_quant through the reference's eager ATen
    chain (the oracle port) on the GPU, to time the SAME QAT step the way the reference runs
    it.  Lives in bench.py on purpose — the product has no such switch."""

    def __enter__(self):
        from mhaq_b200.quantization.gdnsq import gdnsq as G
        from mhaq_b200 import ops
        from oracle import fq_oracle as O
        self.G, self.ops = G, ops
        self.saved = (G.Quantizer.fake_quant, G.Quantizer.fake_quant_eval, G.Quantizer.fake_quant_weight,
                      ops.act_fake_quant)

        def fake_quant(q, value, noise=None):
            return O.fake_quant(value, q.scale, q.zero_point, q.min_val, q.max_val,
                                method=q.qnmethod.name, noise=noise)

        def fake_quant_eval(q, value):
            codes = O.quantize(value, q.scale, q.zero_point, q.min_val, q.max_val)
            mm = codes.aminmax()
            return O.dequantize(codes, q.scale, q.zero_point), torch.stack(
                [mm.min, mm.max, torch.zeros((), device=value.device)])

        def fake_quant_weight(q, weight, log_scale=None, noise=None):
            # NoisyConv2d.forward's weight lines (gdnsq_conv2d.py:72-84,98) op for op; no row
            # range is returned, so ModelHelper re-reduces the weight like the reference does
            if log_scale is not None:
                q.scale = torch.exp2(log_scale)
            q.zero_point = weight.amin(tuple(range(1, weight.dim())), keepdim=True)
            wq = O.fake_quant(weight, q.scale, q.zero_point, -math.inf, math.inf,
                              method=q.qnmethod.name, noise=noise)
            return wq, None, None, None

        def act_fake_quant(x, log_act_s, log_act_q, act_b, method="STE", noise=None, philox=None):
            return O.act_fake_quant(x, log_act_s, log_act_q, act_b, noise=noise, method=method)

        G.Quantizer.fake_quant, G.Quantizer.fake_quant_eval = fake_quant, fake_quant_eval
        G.Quantizer.fake_quant_weight = fake_quant_weight
        ops.act_fake_quant = act_fake_quant
        return self

    def __exit__(self, *a):
        G = self.G
        (G.Quantizer.fake_quant, G.Quantizer.fake_quant_eval, G.Quantizer.fake_quant_weight,
         self.ops.act_fake_quant) = self.saved


def qat_key(a):
    if (a.qat_model, a.qat_method, a.qat_bits) == ("resnet18", "STE", 4):
        return "resnet18_w4a4_qat"
    return f"{a.qat_model}_{a.qat_method.lower()}_w{a.qat_bits}a{a.qat_bits}_qat"


def resnet18_leg(a, dev, world, rank, use_dist, profile_share=True):
    """BASELINE configs[3]: torchvision ResNet-18, ImageNet-shaped synthetic batch (256 per GPU,
    224x224), GDNSQ/STE W4A4 per-channel, distillation (Symmetrical KL) from a frozen FP copy,
    RAdam lr 3e-4, fp32 + TF32 convolutions, DDP over NCCL for N > 1 — through
    Quantizer(config)().quantize(lmodel) and the patched training_step."""
    from mhaq_b200 import harness
    torch.backends.cudnn.benchmark = True
    torch.set_float32_matmul_precision("high")
    B = a.resnet_batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    sr = a.qat_model == "rfdn"
    if sr:      # LR patch in [0,1] (denormalised x255 inside the module), HR target x4
        x = torch.rand(B, 3, 256, 256, device=dev, generator=g)
        t = torch.rand(B, 3, 1024, 1024, device=dev, generator=g)
        q = harness.build_qat("rfdn", dev, qnmethod=a.qat_method, act_bit=a.qat_bits, weight_bit=a.qat_bits,
                              distillation=False, lr=5e-4, calib_batch=x[: min(B, 4)], calib_bits=a.qat_bits)
    else:
        side, classes = (224, 1000) if a.qat_model == "resnet18" else (32, 100)
        x = torch.randn(B, 3, side, side, device=dev, generator=g)
        t = torch.randint(0, classes, (B,), device=dev, generator=g)
        q = harness.build_qat(a.qat_model, dev, qnmethod=a.qat_method, act_bit=a.qat_bits,
                              weight_bit=a.qat_bits, distillation=True, num_classes=classes,
                              calib_batch=x[: min(B, 64)])
    if a.channels_last:
        q.model.to(memory_format=torch.channels_last)
        if getattr(q, "tmodel", None) is not None:
            q.tmodel.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    if use_dist:
        wrap = harness.ddp_side_stream if a.graph else harness.wrap_ddp
        q.model = wrap(q.model, dev, lean=not a.ddp_reference_flags)
    q.train(); q.wrapped_criterion.train()
    if getattr(q, "tmodel", None) is not None:
        q.tmodel.eval()
    opt = q.configure_optimizers()

    def eager_step(batch=None):
        loss = q.training_step(batch or (x, t), 0)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.detach()

    res = {}
    k = a.resnet_steps
    if a.graph and not use_dist:
        # the same step launched eagerly from Python first: its time, and (CUPTI) which share of
        # the GPU time the fake-quant kernels take — a graph replay hides kernel names from CUPTI
        for _ in range(4):
            eager_step()
        res["eager_ms_per_step"] = round(time_region(eager_step, k, False) / k, 2)
    if rank == 0 and not use_dist and profile_share:   # (a rank-local DDP step would dead-lock the other ranks)
        try:
            from torch.profiler import profile, ProfilerActivity
            eager_step()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                eager_step(); eager_step()
                torch.cuda.synchronize()
            tot = fq = 0.0
            for ev in prof.key_averages():
                dt = getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0)
                tot += dt
                if "fq_" in ev.key:
                    fq += dt
            if tot > 0:
                res["fake_quant_kernel_share"] = round(fq / tot, 4)
                res["fake_quant_ms_per_step"] = round(fq / 2 / 1e3, 3)
                res["gpu_kernel_ms_per_step"] = round(tot / 2 / 1e3, 2)   # rest of an eager step = GPU idle
        except Exception as exc:   # profiler unavailable: the throughput numbers stand alone
            res["fake_quant_kernel_share"] = None
            res["profiler_error"] = str(exc)[:80]

    graphed = None
    if a.graph:
        # capture; if it fails on ANY rank every rank falls back to eager launches (the bench
        # line must not depend on a capture succeeding) and says so
        note = None
        try:
            graphed = harness.GraphedTrainStep(q, (x, t), seed=1234 + rank)
        except Exception as exc:
            note = f"{type(exc).__name__}: {str(exc)[:160]}"
        ok = torch.tensor([0 if graphed is None else 1], device=dev)
        if use_dist:
            import torch.distributed as dist
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if graphed is not None:
                graphed.close()
            graphed = None
            from mhaq_b200 import ops as _ops
            _ops.set_device_philox_state(None)
            torch.cuda.synchronize()
            opt = q.configure_optimizers()
            res["graph_capture_failed"] = note or "capture failed on another rank"
    step = graphed if graphed is not None else eager_step

    # end to end: every step's batch comes from pinned host memory (H2D inside the timed region,
    # double-buffered on a copy stream so the copy of batch k+1 overlaps step k) and the step's
    # loss is read back to the host
    hx = x.cpu().pin_memory(); ht = t.cpu().pin_memory()      # (keeps x's memory format)
    hloss = torch.empty((), dtype=torch.float32).pin_memory()
    feed = harness.BatchPrefetcher((x, t))
    feed.put((hx, ht))

    def step_e2e():
        batch = feed.get()
        feed.put((hx, ht))                  # next step's copy, in flight while this step computes
        hloss.copy_(step(batch), non_blocking=True)
        feed.release()

    for _ in range(4):
        step()
    ms = time_region(step, k, use_dist) / k
    ke = max(3, k // 2)
    step_e2e()
    ms_e = time_region(step_e2e, ke, use_dist) / ke
    cfg = {"resnet18": "configs[3] ResNet-18 224x224", "resnet20": "configs[2] ResNet-20 32x32 (CIFAR-100 shaped)",
           "rfdn": "configs[4] RFDN x4 SR, 256x256 LR patches, L1"}[a.qat_model]
    res = {"workload": f"{cfg} {a.qat_method} W{a.qat_bits}A{a.qat_bits} QAT, {'no teacher' if sr else 'distillation'}, RAdam, fp32/TF32, "
                       f"batch {B}/GPU, {'channels_last' if a.channels_last else 'NCHW'}, "
                       f"{'DDP dp%d' % world if use_dist else 'single GPU'}, "
                       f"{'whole step (NCCL all-reduces included) replayed from a CUDA graph' if graphed is not None else 'eager launches'}",
           "img_per_s": round(world * B / (ms * 1e-3), 1), "ms_per_step": round(ms, 2),
           "e2e_img_per_s": round(world * B / (ms_e * 1e-3), 1),
           "e2e_h2d_bytes_per_step": hx.numel() * hx.element_size() + ht.numel() * ht.element_size(),
           "n_gpus": world,
           "quantized_act_elems_per_step": {"resnet18": 1680896, "resnet20": 184320, "rfdn": 59093392}[a.qat_model] * B,
           **res}
    if use_dist:   # replicas must hold identical parameters after the DDP steps
        import torch.distributed as dist
        chk = torch.stack([p.detach().double().sum() for p in q.model.parameters()]).sum().reshape(1)
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        res["ddp_replicas_in_sync"] = bool((lo_ == hi_).item())
    if graphed is not None:
        graphed.close()
    del q, graphed, step, feed
    torch.cuda.empty_cache()
    return res


def eager_reference_leg(a, dev):
    """The reference's own ATen chain (the oracle port, same ops) executed eagerly on the B200:
    the honest speed-up denominator for the fused kernels."""
    from oracle import fq_oracle as O
    b = argparse.Namespace(**vars(a))
    b.log2n = min(a.log2n, 26)
    n = 1 << b.log2n
    x, go, scale, zp, lo, hi = make_inputs(b, dev)
    lo_ = -math.inf if lo is None else lo
    hi_ = math.inf if hi is None else hi
    sp = scale.clone().requires_grad_(True)

    def step():
        xs = x.detach().requires_grad_(True)
        sp.grad = None
        y = O.fake_quant(xs, sp, zp, lo_, hi_, method=a.method)
        y.backward(go)

    for _ in range(3):
        step()
    ms = time_region(step, 10, False) / 10
    return {"value": round(20 * n / (ms * 1e-3) / 1e9, 1), "unit": "GB/s", "ms_per_step": round(ms, 3),
            "what": f"reference ATen op chain (oracle port) run eagerly on this GPU, N=2^{b.log2n}"}


# ---------------------------------------------------------------------------
def run_ours(a):
    import mhaq_b200
    from mhaq_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        import datetime
        if a.graph:   # torch's recipe for capturing DDP's NCCL all-reduces in a CUDA graph
            os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    n = 1 << a.log2n
    x, go, scale, zp, lo, hi = make_inputs(a, dev)
    lo_ = -math.inf if lo is None else lo
    hi_ = math.inf if hi is None else hi
    scale_p = scale.clone().requires_grad_(True)
    launches = {"n": 0}

    def step():
        xs = x.detach().requires_grad_(True)
        scale_p.grad = None
        y = mhaq_b200.fake_quant(xs, scale_p, zp, lo_, hi_, method=a.method)
        y.backward(go)
        launches["n"] += 3 + (2 if a.method == "AEWGS" else 0)   # fwd + bwd + finalize (+ AEWGS stats x2)
        return y, xs.grad, scale_p.grad

    # clocks / throttle reasons are sampled from before the warm-up to the end of the timed
    # region (nvidia-smi needs ~0.1 s to start; the timed region alone can be shorter than that)
    with ClockSampler(local) as cs:
        t0, n_warm = time.perf_counter(), 0
        while n_warm < max(a.warmup, 3) or time.perf_counter() - t0 < 0.35:
            step()                          # (keeps the GPU under load while the sampler starts)
            n_warm += 1
            if n_warm % 8 == 0:
                torch.cuda.synchronize()
        launches["n"] = 0                   # gpu_launches counts the timed region only
        ms = time_region(step, a.steps, use_dist)
    clocks = cs.summary()
    clocks["window"] = f"{n_warm} untimed warm-up steps + the timed region"
    n_l = launches["n"]
    per_step = ms / a.steps
    gbs = world * (BYTES_FWD + BYTES_BWD) * n / (per_step * 1e-3) / 1e9

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        # ---- per-kernel roofline, measured live with CUDA events on the launch stream ----
        L = ops._Launch(x.detach(), scale, zp, lo_, hi_)
        xd = x.detach()

        def fwd_only():
            ops._forward_impl(xd, L, True, False, False)

        def bwd_only():
            ops._backward_impl(go, xd, L, ops._method_id(a.method), False, None, True, philox=(1, 2))

        for f in (fwd_only, bwd_only):
            for _ in range(3):
                f()
        k = max(a.steps, 10)
        t_f = time_region(fwd_only, k, False) / k
        t_b = time_region(bwd_only, k, False) / k
        bwd_bytes = BYTES_BWD + (8 if a.method == "AEWGS" else 0)
        ach_b = bwd_bytes * n / (t_b * 1e-3) / 1e9
        ach_f = BYTES_FWD * n / (t_f * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("fq_bwd_kernel_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": "fq_bwd_kernel (+ deterministic finalize)", "achieved": round(ach_b, 1),
                    "peak": peak, "unit": "GB/s", "frac": round(ach_b / peak, 4), "traffic": traffic,
                    "peak_source": peak_src, "frac_of_8TBps_nominal": round(ach_b / 8000.0, 4),
                    "bytes_per_elem": bwd_bytes, "ms_per_launch": round(t_b, 4),
                    "fwd_kernel": {"achieved": round(ach_f, 1), "frac": round(ach_f / peak, 4),
                                   "bytes_per_elem": BYTES_FWD, "ms_per_launch": round(t_f, 4)}}
        out = {"metric": METRIC, "value": round(gbs, 1), "unit": "GB/s", "n_gpus": world,
               "steps": a.steps, "warmup": max(a.warmup, 3), "warmup_actual": n_warm,
               "ms_per_step": round(per_step, 4),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "impl": "ours",
               "config": {"workload": workload_name(a), "elements_per_gpu": n,
                          "layout": [a.channels, n // a.channels] if a.channels else [n],
                          "method": a.method, "bits": a.bits, "bytes_per_element": 20,
                          "l2": "inputs (1 GiB each) far exceed the 126 MB L2; no flush needed",
                          "multi_gpu": "replicas only (independent tensors per GPU, no data-path collective)"},
               "frac_of_hbm_peak": round(gbs / world / peak, 4),
               "roofline": roofline, "clocks": clocks, "gpu_launches": n_l}

    # ---- end to end through the public API with HOST buffers -------------------------
    if not a.no_e2e:
        try:
            from mhaq_b200.host import fake_quant_fwd_bwd_host
            hx = torch.empty(x.shape, dtype=torch.float32).pin_memory()
            hg = torch.empty(x.shape, dtype=torch.float32).pin_memory()
            hx.copy_(x.detach()); hg.copy_(go)
            hy = torch.empty(x.shape, dtype=torch.float32).pin_memory()
            hgx = torch.empty(x.shape, dtype=torch.float32).pin_memory()
            hgs = torch.empty(scale.shape, dtype=torch.float32).pin_memory()
            del x, go                       # the e2e leg owns its (staged) device memory
            torch.cuda.empty_cache()

            def e2e_step():
                _, _, grads = fake_quant_fwd_bwd_host(hx, hg, scale, zp, lo, hi, method=a.method,
                                                      y_host=hy, gx_host=hgx, chunks=16)
                hgs.copy_(grads["scale"], non_blocking=True)

            ke = max(3, min(a.steps, 5))
            for _ in range(2):
                e2e_step()
            ms_e = time_region(e2e_step, ke, use_dist) / ke
            if rank == 0:
                out["e2e"] = {"value": round(world * 20 * n / (ms_e * 1e-3) / 1e9, 2), "unit": "GB/s",
                              "h2d_bytes_per_step": 2 * 4 * n, "d2h_bytes_per_step": 2 * 4 * n + 4 * scale.numel(),
                              "ms_per_step": round(ms_e, 3), "steps": ke,
                              "api": "mhaq_b200.host.fake_quant_fwd_bwd_host: pinned host x/go in, y/gx/g_scale out, "
                                     "16 row chunks pipelined over full-duplex PCIe (copies inside the timed region)"}
            del hx, hg, hy, hgx
        except Exception as exc:       # e.g. the box refuses 4 GiB of pinned host memory
            if use_dist:
                raise
            if rank == 0:
                out["e2e"] = {"value": None, "unit": "GB/s", "error": f"{type(exc).__name__}: {str(exc)[:200]}"}

    x = go = None
    torch.cuda.empty_cache()
    if not a.no_resnet:
        try:
            rn = resnet18_leg(a, dev, world, rank, use_dist)
        except Exception as exc:       # the QAT leg must never take the headline line down with it
            if use_dist:
                raise
            rn = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            torch.cuda.empty_cache()
        if rank == 0:
            out[qat_key(a)] = rn
    if rank == 0:
        if not a.no_eager_ref:
            out["reference_eager_gpu"] = eager_reference_leg(a, dev)
            if not a.no_resnet and not use_dist:
                key = qat_key(a)
                for cl in ((False, True) if a.channels_last else (False,)):
                    b = argparse.Namespace(**vars(a))
                    b.channels_last, b.graph = cl, False
                    try:
                        with _EagerReferenceBackend():
                            rr = resnet18_leg(b, dev, 1, 0, False, profile_share=False)
                    except Exception as exc:
                        out["reference_eager_gpu"][key + ("_channels_last" if cl else "")] = {
                            "error": f"{type(exc).__name__}: {str(exc)[:200]}"}
                        torch.cuda.empty_cache()
                        continue
                    out["reference_eager_gpu"][key + ("_channels_last" if cl else "")] = {
                        "img_per_s": rr["img_per_s"], "ms_per_step": rr["ms_per_step"],
                        "what": "same QAT step with every fake-quant routed through the reference's eager "
                                "ATen chain on this GPU, " + ("channels_last like our leg" if cl else
                                "row-major NCHW as the reference's Trainer runs it")}
        if not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(a, steps=5, warmup=2)
        if a.sweep:
            sweep(a, dev)
        emit(out)
    if use_dist:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
def cpu_baseline(a, steps, warmup):
    """The reference's CPU quantizer path (oracle port, same ATen op sequence) on the host cores."""
    from oracle import fq_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    log2n = min(a.cpu_log2n, a.log2n)
    n = 1 << log2n
    x, go, scale, zp, lo, hi = make_inputs(a, "cpu", log2n)
    lo_ = -math.inf if lo is None else lo
    hi_ = math.inf if hi is None else hi
    sp = scale.clone().requires_grad_(True)

    def step():
        xs = x.detach().requires_grad_(True)
        sp.grad = None
        y = O.fake_quant(xs, sp, zp, lo_, hi_, method=a.method)
        y.backward(go)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(20 * n / dt / 1e9, 3), "unit": "GB/s", "cores": cores, "kind": "port",
            "ms_per_step": round(dt * 1e3, 2),
            "sample": f"same workload cut to N=2^{log2n} elements, {warmup} warm-up + {steps} timed steps, "
                      f"torch {torch.__version__} CPU, {cores} threads"}


def cpu_config0_step():
    """BASELINE configs[0]: ResNet-20 CIFAR-10 GDNSQ(STE) W4A4 QAT step on the CPU, synthetic
    32x32 batch of 128, through Quantizer(config)().quantize() with every fake-quant routed
    to the reference's eager ATen chain (oracle port) — the reference's own CPU-runnable case."""
    from mhaq_b200 import harness
    torch.manual_seed(0)
    x = torch.randn(128, 3, 32, 32)
    t = torch.randint(0, 10, (128,))
    with _EagerReferenceBackend():
        q = harness.build_qat("resnet20", "cpu", qnmethod="STE", act_bit=4, weight_bit=4,
                              distillation=False, num_classes=10, calib_batch=x[:32])
        opt = q.configure_optimizers()
        q.train(); q.wrapped_criterion.train()
        ts = []
        for i in range(5):
            t0 = time.perf_counter()
            loss = q.training_step((x, t), 0)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            ts.append(time.perf_counter() - t0)
    ts = sorted(ts[2:])
    med = ts[len(ts) // 2]
    return {"workload": "configs[0] ResNet-20 CIFAR-10 GDNSQ(STE) W4A4 QAT step, batch 128, CPU",
            "s_per_step": round(med, 3), "img_per_s": round(128 / med, 1),
            "threads": torch.get_num_threads(), "protocol": "2 warm-up + median of 3"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, a.steps)
    cb = cpu_baseline(a, steps=min(steps, 10), warmup=max(1, min(a.warmup, 3)))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"metric": METRIC, "value": cb["value"], "unit": "GB/s", "n_gpus": world, "steps": min(steps, 10),
           "warmup": max(1, min(a.warmup, 3)), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": workload_name(a), "note": "reference CPU path = oracle port of the "
                      "reference's ATen op sequence (the Python reference cannot travel to the GPU box)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    if not a.no_resnet:
        try:
            out["cpu_config0_step"] = cpu_config0_step()
        except Exception as exc:   # never let the extra leg break the contract line
            out["cpu_config0_step"] = {"error": str(exc)[:120]}
    emit(out)


# ---------------------------------------------------------------------------
def sweep(a, dev):
    """BASELINE config 2: N in 2^20..2^30, layouts, methods, bits.  Writes gpurun_out/sweep.json."""
    import mhaq_b200
    from mhaq_b200 import ops
    peak, _ = peaks()
    rows = []
    for log2n in (20, 22, 24, 26, 28, 30):
        for ch in (0, 64, 512, 4096):
            for method in ("STE", "LSQ", "AEWGS"):
                if method == "AEWGS" and ch == 0:
                    continue
                for bits in ((4,) if log2n != 28 else (1, 2, 4, 8)):
                    b = argparse.Namespace(**vars(a))
                    b.log2n, b.channels, b.method, b.bits = log2n, ch, method, bits
                    try:
                        x, go, scale, zp, lo, hi = make_inputs(b, dev)
                    except torch.OutOfMemoryError:
                        continue
                    n = 1 << log2n
                    lo_ = -math.inf if lo is None else lo
                    hi_ = math.inf if hi is None else hi
                    L = ops._Launch(x, scale, zp, lo_, hi_)
                    mid = ops._method_id(method)
                    f = lambda: ops._forward_impl(x, L, True, False, False)
                    g = lambda: ops._backward_impl(go, x, L, mid, False, None, True, philox=(1, 2))
                    reps = 50 if log2n <= 24 else (20 if log2n <= 28 else 5)
                    res = {}
                    # tensors that fit the 126 MB L2 are evicted between iterations by a 256 MB
                    # write; its own time is measured the same way and subtracted.  Launches are
                    # enqueued back to back so the figure is GPU time, not host launch latency.
                    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev) if log2n <= 24 else None

                    def timed(body):
                        for _ in range(3):
                            body()
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(reps):
                            body()
                        e1.record()
                        torch.cuda.synchronize()
                        return e0.elapsed_time(e1) / reps

                    t_flush = timed(lambda: flush.zero_()) if flush is not None else 0.0
                    for nm, fn, by in (("fwd", f, 8), ("bwd", g, 12 + (8 if method == "AEWGS" else 0))):
                        if flush is not None:
                            med = max(timed(lambda: (flush.zero_(), fn())) - t_flush, 1e-6)
                        else:
                            med = timed(fn)
                        res[nm] = {"ms_median": round(med, 5), "GBps": round(by * n / med / 1e6, 1),
                                   "frac": round(by * n / med / 1e6 / peak, 4)}
                    tot = res["fwd"]["ms_median"] + res["bwd"]["ms_median"]
                    by = 20 + (8 if method == "AEWGS" else 0)
                    rows.append({"log2n": log2n, "channels": ch, "method": method, "bits": bits, **res,
                                 "fwd_bwd_GBps": round(by * n / tot / 1e6, 1),
                                 "fwd_bwd_frac": round(by * n / tot / 1e6 / peak, 4)})
                    print("sweep", json.dumps(rows[-1]), file=sys.stderr, flush=True)
                    del x, go, L
                    torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"peak_GBps": peak, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


_JSON_FD = None


def _claim_stdout():
    """ONE JSON line on stdout: NCCL (NCCL_DEBUG=VERSION/INFO) and other native libraries print
    to fd 1; keep a private copy of the real stdout for the JSON line and point fd 1 at stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    a = parse_args()
    _claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                             "(use --impl reference for the CPU baseline)")
        run_ours(a)


if __name__ == "__main__":
    main()

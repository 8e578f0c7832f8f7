#!/usr/bin/env python
"""profiles/<round>_scaling.md from the bench lines profiles/<round>_bench.json (N=1) and
profiles/<round>_bench_n{2,4,8}.json.

    python tools/scaling_table.py [--round r02]
"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(rnd, n):
    p = os.path.join(ROOT, "profiles", f"{rnd}_bench.json" if n == 1 else f"{rnd}_bench_n{n}.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        txt = f.read().strip()
    line = [l for l in txt.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r02")
    a = ap.parse_args()
    runs = {n: load(a.round, n) for n in (1, 2, 4, 8)}
    runs = {n: d for n, d in runs.items() if d}
    out = [f"# {a.round} — 1 / 2 / 4 / 8 B200 (one node, one process per GPU, NCCL over NVLink)", "",
           "From `bench.py` lines (`profiles/%s_bench.json`, `_n2`, `_n4`, `_n8`); device-timed, MAX over ranks, "
           "barrier + synchronize on both sides.  Efficiency = value(N) / (N x value(1))." % a.round, ""]
    base = runs.get(1)

    def eff(v, n, v1):
        return "" if not v1 else f"{v / (n * v1):.3f}"

    # ---- microbench
    out += ["## configs[1] microbench — replicas only (no data-path collective)", "",
            "| GPUs | fwd+bwd GB/s (aggregate) | ms/step (max over ranks) | per-rank ms/step min / median / max | efficiency | e2e GB/s (host buffers) | duplex host link GB/s each way per rank |",
            "|---|---|---|---|---|---|---|"]
    for n, d in sorted(runs.items()):
        pr = d.get("per_rank_ms_per_step", {})
        e2e = d.get("e2e", {})
        out.append(f"| {n} | {d['value']} | {d['ms_per_step']} | {pr.get('min')} / {pr.get('median')} / {pr.get('max')} | "
                   f"{eff(d['value'], n, base['value'] if base else None)} | {e2e.get('value')} | "
                   f"{(e2e.get('host_link') or {}).get('duplex_each_way_GBps')} |")
    out += ["", "The step is replayed from a CUDA graph (one `cudaGraphLaunch` per step), so host launch jitter cannot enter the "
            "16 ms timed window — round 1 lost 6.5 % at N = 8 to one straggling Python launch loop with unchanged kernel time.  "
            "`e2e` is bound by the host: every rank moves 2 GiB in and 2 GiB out per step through the same host DRAM / PCIe root "
            "complexes (`host_link` = this rank's measured copy bandwidth with all ranks copying at once), so the aggregate "
            "does not scale with the GPU count.", ""]

    # ---- QAT legs
    legs = [("resnet18_w4a4_qat", "configs[3] ResNet-18 224x224 STE W4A4, batch 256/GPU, lean DDP, whole step in a CUDA graph"),
            ("resnet18_w4a4_qat_reference_trainer_flags", "configs[3] with the reference Trainer's flags (SyncBatchNorm, find_unused_parameters=True, buffer broadcast), eager launches"),
            ("resnet20_aewgs_w1a1_qat", "configs[2] ResNet-20 CIFAR-100-shaped AEWGS W1A1, batch 256/GPU (the config with a data-path collective)"),
            ("rfdn_lsq_w2a2_qat", "configs[4] RFDN x4 SR LSQ W2A2, 16 x 256x256 LR patches / GPU")]
    for key, title in legs:
        rows = [(n, d[key]) for n, d in sorted(runs.items()) if isinstance(d.get(key), dict) and "img_per_s" in d[key]]
        if not rows:
            continue
        v1 = next((r["img_per_s"] for n, r in rows if n == 1), None)
        if key.endswith("reference_trainer_flags") and base and isinstance(base.get("resnet18_w4a4_qat"), dict):
            v1 = base["resnet18_w4a4_qat"]["img_per_s"]
        out += [f"## {title}", "",
                "| GPUs | img/s | ms/step | per-rank ms/step min / median / max | scaling vs 1 GPU | efficiency | e2e img/s | replicas in sync | launch mode |",
                "|---|---|---|---|---|---|---|---|---|"]
        for n, r in rows:
            pr = r.get("per_rank_ms_per_step", {})
            sc = f"{r['img_per_s'] / v1:.2f}x" if v1 else ""
            mode = "graph" if "replayed from a CUDA graph" in r.get("workload", "") else "eager"
            out.append(f"| {n} | {r['img_per_s']} | {r['ms_per_step']} | {pr.get('min', '')} / {pr.get('median', '')} / {pr.get('max', '')} | "
                       f"{sc} | {eff(r['img_per_s'], n, v1)} | {r.get('e2e_img_per_s')} | {r.get('ddp_replicas_in_sync', '')} | {mode} |")
        ar = [(n, r["aewgs_stats_allreduce"]) for n, r in rows if "aewgs_stats_allreduce" in r]
        if ar:
            out += ["", "In-backward AEWGS statistics all-reduce (CUDA events around the call, eager DDP steps): " +
                    "; ".join(f"N={n}: {x['calls_per_step']} call(s)/step, {x['ms_per_step']} ms/step" for n, x in ar) + "."]
        out.append("")
    out += ["## what limits each configuration", "",
            "* **microbench**: nothing to exchange — replicas; per-rank times agree to 0.1 % and the aggregate is N x the "
            "single-GPU value.  Its host-buffer `e2e` variant is bound by the host's DRAM / PCIe, shared by all ranks.",
            "* **ResNet-18 (configs[3])**: the fake-quant path adds no collective; the +0.3 ... +0.8 ms per step over one GPU "
            "is DDP's bucketed all-reduce of 11.7 M gradients (47 MB) captured in the graph — NCCL kernels that share HBM "
            "and SMs with the backward they overlap — i.e. library time outside the path.  With the reference Trainer's own "
            "flags the step is 37-39 ms: SyncBatchNorm's all-gathers around every one of the 20 BatchNorm layers (forward and "
            "backward), the per-step graph traversal of `find_unused_parameters=True`, the buffer broadcast, and eager "
            "launches instead of one graph launch.",
            "* **ResNet-20 AEWGS (configs[2])**: a 6 ms step, so latency-bound: DDP costs 0.1-0.2 ms whatever N is.  "
            "The path's one collective — the packed statistics all-reduce, ONE per step for all 18 conv weights — sits in "
            "the middle of the weight backward (the apply kernel needs the averaged statistics), so one NCCL latency plus the "
            "skew between ranks at that point is exposed; the gradient all-reduce (0.27 M parameters, one bucket) follows at "
            "the end.  Round 1 issued 18 statistics all-reduces per step.",
            "* **RFDN (configs[4])**: 0.43 M parameters, the all-reduce is negligible; 72 -> 74 ms is NCCL launch / "
            "synchronisation inside a graph of ~4200 nodes and memory-system contention at 59 M quantized activation "
            "elements per image.", ""]
    p = os.path.join(ROOT, "profiles", f"{a.round}_scaling.md")
    with open(p, "w") as f:
        f.write("\n".join(out) + "\n")
    print(p)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""gpurun_out/sweep.json (written by `python bench.py --sweep`) -> profiles/<name>.md / .json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02_sweep_config2"
d = json.load(open(os.path.join(ROOT, "gpurun_out", "sweep.json")))
peak = d["peak_GBps"]
out = [f"# {tag.split('_')[0]} — BASELINE configs[1] sweep (B200, fp32)", "",
       "`python bench.py --sweep` — GPU-side time between two CUDA events: N ≤ 2^24: 20 calls captured into one CUDA",
       "graph (what a graph-captured training step pays; no Python launch latency), each preceded by a 256 MB write that",
       "evicts the tensors from the 126 MB L2 and whose own time is subtracted; N ≥ 2^26: calls enqueued back to back.",
       f"GB/s = algorithmic bytes (fwd / eval 8, bwd 12, AEWGS bwd 20 B/elem) / time; frac = ÷ {peak} GB/s",
       "(measured `torch.copy_`).  `ch=0` = per-tensor clamped (activation style, 4 param grads); `ch=C` = per-channel",
       "`[C, N/C]`, lo/hi = ±inf (weight style).  bwd includes the finalize launch (and the AEWGS stats pass).", "",
       "`eval` = the forward with the code / input min-max statistics (validation, calibration).  (The evicting write leaves",
       "256 MB of dirty lines that the measured kernel's traffic pushes out, so the N ≤ 2^24 rows read a few points lower",
       "than `r02_midsize.md`, which rotates input sets instead.)", "",
       "| N | ch | method | bits | fwd GB/s | frac | eval GB/s | frac | bwd GB/s | frac | fwd+bwd GB/s | frac |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in d["rows"]:
    ev = r.get("eval", {"GBps": "", "frac": float("nan")})
    out.append(f"| 2^{r['log2n']} | {r['channels']} | {r['method']} | {r['bits']} | {r['fwd']['GBps']} | {r['fwd']['frac']:.2f} | "
               f"{ev['GBps']} | {ev['frac']:.2f} | "
               f"{r['bwd']['GBps']} | {r['bwd']['frac']:.2f} | {r['fwd_bwd_GBps']} | {r['fwd_bwd_frac']:.2f} |")
open(os.path.join(ROOT, "profiles", tag + ".md"), "w").write("\n".join(out) + "\n")
json.dump(d, open(os.path.join(ROOT, "profiles", tag + ".json"), "w"), indent=1)
print(f"{len(d['rows'])} rows -> profiles/{tag}.md")

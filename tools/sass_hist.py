#!/usr/bin/env python
"""SASS opcode histogram and register / shared-memory table of the built library ->
profiles/<round>_sass_histogram.md.  Runs anywhere (cuobjdump on the .so, no GPU needed).

    python tools/sass_hist.py [--round r02]
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mhaq_b200", "csrc", "libmhaq_fq.so")


def demangle_short(name):
    out = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    out = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", out)
    out = re.sub(r"^void ", "", out)
    out = re.sub(r"\((?:int|bool)\)", "", out)
    return out.split("(")[0][:90]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r02")
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    per_fn, cur = collections.OrderedDict(), None
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_fn[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            per_fn[cur][op] += 1
            total[op] += 1
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = tuple(int(v) for v in m.groups())
    lines = [f"# {a.round} — SASS opcode histogram and resource table of `libmhaq_fq.so` (sm_100a)", "",
             "`python tools/sass_hist.py` (cuobjdump -sass / -res-usage on the in-tree build; static counts per",
             "kernel instantiation, not executed counts).", "",
             "## whole library: packed fp32x2, memory, TMA / mbarrier opcodes", "",
             "| opcode | static count |", "|---|---|"]
    keys = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FSETP", "FSEL", "FMNMX", "FRND", "MUFU.RCP", "MUFU.EX2",
            "LDG.E.NA.128.CONSTANT", "STG.E.NA.128", "LDS.128", "UBLKCP.S.G", "SYNCS.ARRIVE.TRANS64",
            "SYNCS.PHASECHK.TRANS64.TRYWAIT", "ATOMG.E.ADD.STRONG.GPU", "MEMBAR.ALL.GPU", "SHFL.BFLY", "BAR.SYNC.DEFER_BLOCKING"]
    for k in keys:
        n = sum(v for op, v in total.items() if op == k or op.startswith(k + "."))
        lines.append(f"| `{k}` | {n} |")
    fl_atomics = sum(v for op, v in total.items() if op.startswith(("ATOMG", "RED", "ATOMS")) and (".F32" in op or ".F64" in op))
    lines += ["", f"Floating-point atomics in the library: **{fl_atomics}** (the only atomics are the integer tickets).", "",
              "## per kernel instantiation", "",
              "| kernel | regs | stack | static smem | FFMA2 | FADD2 | FMUL2 | LDG.128 | LDS.128 | STG.128 | UBLKCP | total instr |",
              "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    want = ("fq_fwd_kernel", "fq_bwd_kernel", "fq_bwd_flat_kernel", "fq_wrow", "fq_aewgs_stats_kernel", "fq_bwd_finalize",
            "fq_potential_loss", "fq_minmax_finalize")
    for fn, c in per_fn.items():
        short = demangle_short(fn)
        if not any(w in short for w in want):
            continue
        r = regs.get(fn, ("?", "?", "?"))
        g = lambda p: sum(v for op, v in c.items() if op.startswith(p))
        lines.append(f"| `{short}` | {r[0]} | {r[1]} | {r[2]} | {g('FFMA2')} | {g('FADD2')} | {g('FMUL2')} | "
                     f"{sum(v for op, v in c.items() if op.startswith('LDG') and '128' in op)} | {g('LDS.128')} | "
                     f"{sum(v for op, v in c.items() if op.startswith('STG') and '128' in op)} | {g('UBLKCP')} | {sum(c.values())} |")
    out = os.path.join(ROOT, "profiles", f"{a.round}_sass_histogram.md")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(out, len(per_fn), "kernels")


if __name__ == "__main__":
    main()

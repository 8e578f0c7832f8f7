#!/bin/bash
# HISTORICAL (results: profiles/r02_exp_flat3.txt).  The variants were compile-time switches
# (-DMHAQ_FLAT_SYNC=0|1 -DMHAQ_FLAT_POLL=0|1) of the work tree between commits 5fbdd13 and f20cc5e, built into
# tools/_exp/ (git-ignored); the losing paths were removed from the source afterwards.
# Flat backward v2: "empty"-mbarrier ring (S) and polled self-validating records (P) against the
# block-barrier ring + fence/ticket epilogue (s0p0 = what r02_midsize.md's "final" row measured).
echo "== correctness (default build = S1 P1)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_reference.py -m gpu -q -x 2>&1 | tail -3
for cfg in "X=1" "MHAQ_FQ_LIB=/root/repo/tools/_exp/libmhaq_fq_s0p0.so" "MHAQ_FQ_LIB=/root/repo/tools/_exp/libmhaq_fq_s1p0.so" "MHAQ_FQ_LIB=/root/repo/tools/_exp/libmhaq_fq_s0p1.so" "MHAQ_FQ_FLAT_EXP=1"; do
  echo "== quick: $cfg"
  env $cfg timeout 200 python tools/midsize_graph.py --quick --out gpurun_out/tmp_exp.json 2>&1
done
for cfg in "X=1" "MHAQ_FQ_LIB=/root/repo/tools/_exp/libmhaq_fq_s0p0.so"; do
  echo "== large: $cfg"
  env $cfg timeout 300 python tools/midsize_graph.py --large --out gpurun_out/tmp_exp.json 2>&1
done

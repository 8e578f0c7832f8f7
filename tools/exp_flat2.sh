#!/bin/bash
# Aux-warp flat backward: defaults (4 blocks/SM, interleave >= 2^24, 2 stages) vs 3 stages, epilogue cost.
for cfg in "X=1" "MHAQ_FQ_FLAT_EXP=1" "MHAQ_FQ_LIB=/root/repo/tools/_exp/libmhaq_fq_s3.so" "MHAQ_FQ_FLAT_INTERLEAVE_LOG2=23" "MHAQ_FQ_FLAT_CTAS_PER_SM=5"; do
  echo "== quick: $cfg"
  env $cfg timeout 200 python tools/midsize_graph.py --quick --out gpurun_out/tmp_exp.json 2>&1
done
for cfg in "X=1" "MHAQ_FQ_LIB=/root/repo/tools/_exp/libmhaq_fq_s3.so" "MHAQ_FQ_FLAT_CTAS_PER_SM=3"; do
  echo "== large: $cfg"
  env $cfg timeout 300 python tools/midsize_graph.py --large --out gpurun_out/tmp_exp.json 2>&1
done

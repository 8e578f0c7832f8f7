#!/bin/bash
# Flat backward experiments: epilogue cost, CTAs per SM, vs the streaming kernel (MAX_LOG2=0),
# and at large sizes blocked vs interleaved partition vs the streaming kernel.
for cfg in "X=1" "MHAQ_FQ_FLAT_EXP=1" "MHAQ_FQ_FLAT_CTAS_PER_SM=4" "MHAQ_FQ_FLAT_CTAS_PER_SM=3" "MHAQ_FQ_FLAT_MAX_LOG2=0"; do
  echo "== $cfg"
  env $cfg timeout 200 python tools/midsize_graph.py --quick --out gpurun_out/tmp_exp.json 2>&1
done
for cfg in "MHAQ_FQ_FLAT_MAX_LOG2=0" "MHAQ_FQ_FLAT_MAX_LOG2=28" "MHAQ_FQ_FLAT_MAX_LOG2=28 MHAQ_FQ_FLAT_INTERLEAVE_LOG2=20" "MHAQ_FQ_FLAT_MAX_LOG2=28 MHAQ_FQ_FLAT_INTERLEAVE_LOG2=20 MHAQ_FQ_FLAT_CTAS_PER_SM=4"; do
  echo "== large: $cfg"
  env $cfg timeout 300 python tools/midsize_graph.py --large --out gpurun_out/tmp_exp.json 2>&1
done

#!/bin/bash
# In-situ comparison (ResNet-18 b256 eager step, CUPTI): flat TMA backward (default) vs the streaming
# kernel + finalize (MHAQ_FQ_FLAT_MAX_LOG2=0) vs flat with a blocked partition only.
for cfg in "X=1" "MHAQ_FQ_FLAT_MAX_LOG2=0" "MHAQ_FQ_FLAT_INTERLEAVE_LOG2=62" "MHAQ_FQ_FLAT_INTERLEAVE_LOG2=22"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/step_breakdown.py --model resnet18 --batch 256 --channels-last --top 60 2>&1 | grep "^#\|fq_"
done

#!/bin/bash
# Programmatic dependent launch of the finalize kernels: on (default) vs MHAQ_FQ_NO_PDL=1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for cfg in "X=1" "MHAQ_FQ_NO_PDL=1" "X=1" "MHAQ_FQ_NO_PDL=1"; do
  echo "== $cfg"
  env $cfg timeout 400 python tools/midsize_graph.py --out gpurun_out/tmp.json 2>&1 | grep "sweep\|act (256,512,7,7)\|act (256,64,56"
done

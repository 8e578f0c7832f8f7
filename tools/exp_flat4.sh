#!/bin/bash
# HISTORICAL (results: profiles/r02_exp_flat4.txt).  Variants = -DMHAQ_FLAT_SYNC / -DMHAQ_FLAT_POLL / -DMHAQ_FLAT_READER0
# builds of the work tree just before commit f20cc5e (tools/_exp/, git-ignored); the losing paths were removed afterwards.
# flat backward v2 diagnosis: which of {ring release (S), polled records (P), reader block (R)} costs what, at which size
E=/root/repo/tools/_exp
for lib in s0p0 s1p0 s0p1r0 s0p1r1 s1p1r0; do
  for mode in quick large; do
    echo "== $mode: $lib"
    MHAQ_FQ_LIB=$E/libmhaq_fq_$lib.so timeout 300 python tools/midsize_graph.py --$mode --out gpurun_out/tmp_exp.json 2>&1
  done
done
for lib in s0p1r0 s0p1r1; do
  echo "== large, epilogue skipped: $lib"
  MHAQ_FQ_FLAT_EXP=1 MHAQ_FQ_LIB=$E/libmhaq_fq_$lib.so timeout 300 python tools/midsize_graph.py --large --out gpurun_out/tmp_exp.json 2>&1
done

#!/bin/bash
# HISTORICAL (results: profiles/r02_exp_flat5.txt).  sync1 = -DMHAQ_FLAT_SYNC=1 build of the work tree just before commit
# b5afa74; both variants now live in the library as the MBAR template parameter (MHAQ_FQ_FLAT_MBAR_LOG2 switches at run time).
# flat backward: block barrier per batch (default) vs the same refill point signalled through an
# "empty" mbarrier (no block barrier, still ONE batch in flight during compute)
E=/root/repo/tools/_exp
MHAQ_FQ_LIB=$E/libmhaq_fq_sync1.so timeout 300 python -m pytest tests/test_gpu_live_reference.py tests/test_gpu_parity.py -m gpu -q -x -k "flat or full_size or determin" 2>&1 | tail -2
for lib in default sync1 default sync1; do
  for mode in quick large; do
    echo "== $mode: $lib"
    if [ $lib = default ]; then L=/root/repo/mhaq_b200/csrc/libmhaq_fq.so; else L=$E/libmhaq_fq_$lib.so; fi
    MHAQ_FQ_LIB=$L timeout 300 python tools/midsize_graph.py --$mode --out gpurun_out/tmp_exp.json 2>&1
  done
done

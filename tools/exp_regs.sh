for m in ${MS:-6 7 8}; do for spt in ${SPTS:-1 2 4}; do for ch in 0 512; do LIB=""; [ $m != 7 ] && LIB=$PWD/tools/_exp/libm$m.so; MHAQ_FQ_LIB=$LIB MHAQ_FQ_BWD_SPT=$spt python bench.py --steps 10 --warmup 3 --channels $ch --no-e2e --no-cpu-baseline > gpurun_out/e.json 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/e.json').read().strip().splitlines()[-1]); r=d['roofline']
print('min_ctas=$m bwd_spt=$spt channels=$ch fwd+bwd GB/s',d['value'],'| bwd GB/s',r['achieved'],'ms',r['ms_per_launch'],'| fwd GB/s',r['fwd_kernel']['achieved'])"; done; done; done

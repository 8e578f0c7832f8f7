# task-granularity experiment: MHAQ_FQ_BWD_SPT / MHAQ_FQ_FWD_SPT = sub-tiles (4096 elements) per CTA task
for spt in ${SPTS:-1 2 4 8 16}; do for ch in 0 512; do MHAQ_FQ_BWD_SPT=$spt MHAQ_FQ_FWD_SPT=$spt python bench.py --steps 10 --warmup 3 --channels $ch --no-e2e --no-cpu-baseline > gpurun_out/e.json 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/e.json').read().strip().splitlines()[-1]); r=d['roofline']
print('spt=$spt channels=$ch fwd+bwd GB/s',d['value'],'| bwd GB/s',r['achieved'],'ms',r['ms_per_launch'],'| fwd GB/s',r['fwd_kernel']['achieved'],'ms',r['fwd_kernel']['ms_per_launch'])"; done; done

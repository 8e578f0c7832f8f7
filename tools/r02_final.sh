#!/bin/bash
# Final round-2 measurement pass on one B200: tests, smoke, the bench line (both arms), the
# mid-size table, step breakdowns, the per-tensor launch list and one ncu --set full capture of the
# flat backward (the streaming kernels are unchanged since r02_profile.sh).
set -x
B="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-resnet --no-eager-ref --no-graph-microbench"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo rc=$?
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; echo rc=$?
timeout 600 python tools/midsize_graph.py --out gpurun_out/r02_midsize_final.json > gpurun_out/r02_midsize_final.txt 2>&1
timeout 300 python tools/midsize_graph.py --large --out gpurun_out/tmp.json > gpurun_out/r02_midsize_final_large.txt 2>&1
timeout 300 python tools/exp_flat_huge.py > gpurun_out/r02_exp_flat_huge.txt 2>&1
timeout 300 python tools/step_breakdown.py --model resnet20 --batch 256 --channels-last --top 40 > gpurun_out/r02_resnet20_step_breakdown.txt 2>&1
timeout 300 python tools/step_breakdown.py --model resnet18 --batch 256 --channels-last --top 30 > gpurun_out/r02_resnet18_step_breakdown.txt 2>&1
timeout 200 python bench.py $B --channels 0 > gpurun_out/r02_pre_ncu_pt.json 2> gpurun_out/r02_pre_ncu_pt.err; echo rc=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_pertensor.csv python bench.py $B --channels 0 > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fq_bwd_flat_kernel|fq_fwd_kernel" -s 6 -c 2 -f -o gpurun_out/r02_pertensor python bench.py $B --channels 0 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# Round-2 profile run (1 GPU): shape sweep with size-matched ceilings, ResNet-20 launch breakdown,
# ncu launch list and full captures of the dominant kernels.  Outputs -> gpurun_out/.
set -x
B="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-resnet --no-eager-ref --no-graph-microbench"
timeout 400 python tools/midsize_graph.py --out gpurun_out/r02_midsize_after.json > gpurun_out/r02_midsize_after.txt 2>&1
timeout 300 python tools/step_breakdown.py --model resnet20 --batch 256 --channels-last --top 40 > gpurun_out/r02_resnet20_step_breakdown.txt 2>&1
timeout 300 python tools/step_breakdown.py --model resnet18 --batch 256 --channels-last --top 30 > gpurun_out/r02_resnet18_step_breakdown.txt 2>&1
# the commands exit 0 without ncu first
timeout 200 python bench.py $B > gpurun_out/r02_pre_ncu_pc.json 2> gpurun_out/r02_pre_ncu_pc.err; echo rc=$?
timeout 200 python bench.py $B --channels 0 > gpurun_out/r02_pre_ncu_pt.json 2> gpurun_out/r02_pre_ncu_pt.err; echo rc=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $B > gpurun_out/ncu_a.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_pertensor.csv python bench.py $B --channels 0 > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fq_bwd_kernel|fq_fwd_kernel|fq_bwd_finalize" -s 9 -c 3 -f -o gpurun_out/r02_perchannel python bench.py $B > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fq_bwd_flat_kernel|fq_fwd_kernel" -s 6 -c 2 -f -o gpurun_out/r02_pertensor python bench.py $B --channels 0 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
W="--workload cornell_neuralq --steps 1 --warmup 3 --width 128 --height 128 --batch 4096 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_dqn_backward' -s 100 -c 1 -f -o gpurun_out/r2_prof_bwd python bench.py $W > gpurun_out/r2_ncu_bwd.log 2>&1; tail -2 gpurun_out/r2_ncu_bwd.log

#!/bin/bash
mkdir -p gpurun_out
W="--workload cornell_neuralq --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dqn_forward -c 2 -f -o gpurun_out/r2_prof_dqn_fwd_b python bench.py $W > gpurun_out/r2_ncu_dqn_b.log 2>&1; tail -2 gpurun_out/r2_ncu_dqn_b.log

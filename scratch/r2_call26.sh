#!/bin/bash
mkdir -p gpurun_out
W="--workload cornell_neuralq --steps 1 --warmup 3 --width 128 --height 128 --batch 4096 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 3000 -c 200 --csv --log-file gpurun_out/r2_launches_nq_26.csv python bench.py $W > gpurun_out/r2_ncu_nq_26.log 2>&1
echo done

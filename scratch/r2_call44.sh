#!/bin/bash
mkdir -p gpurun_out
for w in cornell_neuralq archway_neuralq; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final2_${w}.json 2> gpurun_out/r2_bench_final2_${w}.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final2_${w}.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame','gpu_launches')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
W="--workload archway_neuralq --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dqn_forward|k_dqn_backward' -c 3 -f -o gpurun_out/r2_prof_dqn_final python bench.py $W > gpurun_out/r2_ncu_dqn_final.log 2>&1; tail -1 gpurun_out/r2_ncu_dqn_final.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dqn.py tests/test_gpu_host_mirror.py -m gpu -x -q > gpurun_out/r2_pytest_dqn19.log 2>&1; echo "pytest dqn rc=$?"; tail -8 gpurun_out/r2_pytest_dqn19.log | cut -c1-800
for w in cornell_neuralq archway_neuralq; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_19.json 2> gpurun_out/r2_bench_${w}_19.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_19.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
W="--workload cornell_neuralq --steps 1 --warmup 3 --width 128 --height 128 --batch 4096 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/r2_launches_nq_19.csv python bench.py $W > gpurun_out/r2_ncu_nq_19.log 2>&1

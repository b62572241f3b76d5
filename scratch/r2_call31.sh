#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q > gpurun_out/r2_pytest_dqn31.log 2>&1; echo "pytest dqn rc=$?"; tail -4 gpurun_out/r2_pytest_dqn31.log | cut -c1-1200
for w in cornell_neuralq archway_neuralq; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_31.json 2> gpurun_out/r2_bench_${w}_31.err; echo "$w rc=$?"; tail -2 gpurun_out/r2_bench_${w}_31.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_31.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame','gpu_launches')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
timeout 200 python bench.py --workload cornell_neuralq --steps 2 --warmup 3 --no-cpu-baseline --width 200 --height 200 --batch 4096 > gpurun_out/r2_bench_ragged_31.json 2> gpurun_out/r2_bench_ragged_31.err; echo "ragged rc=$?"; tail -2 gpurun_out/r2_bench_ragged_31.err | cut -c1-300

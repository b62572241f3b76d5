#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q > gpurun_out/r2_pytest_dqn24.log 2>&1; echo "pytest dqn rc=$?"; tail -4 gpurun_out/r2_pytest_dqn24.log | cut -c1-800
for lib in librlpt.so librlpt_r1.so; do
for w in cornell_neuralq; do
  RLPT_LIB_NAME=$lib timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_24.json 2> gpurun_out/r2_bench_${w}_24.err; echo "$lib $w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_24.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done; done

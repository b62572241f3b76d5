#!/bin/bash
mkdir -p gpurun_out
for lib in librlpt.so librlpt_k32.so librlpt_p4.so librlpt_k32p2.so; do
RLPT_LIB_NAME=$lib timeout 400 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q -k "forward" > gpurun_out/r2_pytest_dqn17.log 2>&1; echo "$lib pytest dqn rc=$?"; tail -1 gpurun_out/r2_pytest_dqn17.log | cut -c1-300
for w in cornell_neuralq; do
  RLPT_LIB_NAME=$lib timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_17.json 2> gpurun_out/r2_bench_${w}_17.err; echo "$lib $w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_17.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done; done

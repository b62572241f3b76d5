#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q > gpurun_out/r2_pytest_dqn25.log 2>&1; echo "pytest dqn rc=$?"; tail -6 gpurun_out/r2_pytest_dqn25.log | cut -c1-1200
for f in 1 0; do
for w in cornell_neuralq; do
  RLPT_NQ_FUSED_BWD=$f timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_25.json 2> gpurun_out/r2_bench_${w}_25.err; echo "fused=$f $w rc=$?"; tail -2 gpurun_out/r2_bench_${w}_25.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_25.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done; done
RLPT_NQ_FUSED_BWD=0 timeout 600 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q -k "training or adam or neuralq" > gpurun_out/r2_pytest_dqn25b.log 2>&1; echo "pytest dqn (unfused) rc=$?"; tail -3 gpurun_out/r2_pytest_dqn25b.log | cut -c1-600

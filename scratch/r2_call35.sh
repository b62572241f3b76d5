#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q > gpurun_out/r2_pytest_dqn35.log 2>&1; echo "pytest dqn rc=$?"; tail -4 gpurun_out/r2_pytest_dqn35.log | cut -c1-1200
for lib in librlpt.so librlpt_k64.so; do
  RLPT_LIB_NAME=$lib timeout 300 python bench.py --workload cornell_neuralq --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$lib', {k:d[k] for k in ('ms_per_step','us_per_optimiser_step')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done

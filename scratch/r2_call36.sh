#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q > gpurun_out/r2_pytest_dqn36.log 2>&1; echo "pytest dqn rc=$?"; tail -4 gpurun_out/r2_pytest_dqn36.log | cut -c1-1500
for w in cornell_neuralq archway_neuralq; do
  timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$w', {k:d[k] for k in ('ms_per_step','us_per_optimiser_step')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done

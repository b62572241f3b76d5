#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all6.log 2>&1; echo "pytest all rc=$?"; tail -15 gpurun_out/r2_pytest_all6.log | cut -c1-400
for w in cornell_neuralq archway_neuralq; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2_bench_${w}_b.json 2> gpurun_out/r2_bench_${w}_b.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_b.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'])"
done
timeout 300 bash scratch/kstats.sh "X=1" --workload cornell_sarsa 2>&1 | tail -1

#!/bin/bash
# forward kernel rewrite (warp-specialised, per-chunk epilogue) + BVH push variants
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q > gpurun_out/r2_pytest_dqn12.log 2>&1; echo "pytest dqn rc=$?"; tail -5 gpurun_out/r2_pytest_dqn12.log | cut -c1-400
for w in cornell_neuralq archway_neuralq; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_12.json 2> gpurun_out/r2_bench_${w}_12.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_12.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "closest" > gpurun_out/r2_pytest_bvh12.log 2>&1; echo "pytest bvh rc=$?"; tail -3 gpurun_out/r2_pytest_bvh12.log | cut -c1-300
RLPT_LIB_NAME=librlpt_u.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "closest" > gpurun_out/r2_pytest_bvh12u.log 2>&1; echo "pytest bvh unified rc=$?"; tail -3 gpurun_out/r2_pytest_bvh12u.log | cut -c1-300
REPS=2 BENCH_ARGS="--workload medieval_inside_default --no-exclusive --steps 16" timeout 600 bash scratch/ab.sh "RLPT_LIB_NAME=librlpt_old.so" "RLPT_LIB_NAME=librlpt.so" "RLPT_LIB_NAME=librlpt_u.so"
REPS=1 BENCH_ARGS="--workload archway_sarsa --no-exclusive --steps 16" timeout 600 bash scratch/ab.sh "RLPT_LIB_NAME=librlpt_old.so" "RLPT_LIB_NAME=librlpt.so" "RLPT_LIB_NAME=librlpt_u.so"

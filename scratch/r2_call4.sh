#!/bin/bash
mkdir -p gpurun_out
for lib in librlpt.so librlpt_upop.so; do
  for w in medieval_inside_default archway_sarsa medieval_default; do
    RLPT_LIB_NAME=$lib timeout 300 bash scratch/kstats.sh "RLPT_LIB_NAME=$lib" --workload $w 2>&1 | tail -1
  done
done
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_all4.log 2>&1; echo "pytest all rc=$?"; tail -12 gpurun_out/r2_pytest_all4.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench_default.json
timeout 900 python scratch/mape_ablation.py > gpurun_out/r2_ablation.log 2>&1; echo "ablation rc=$?"; tail -5 gpurun_out/r2_ablation.log

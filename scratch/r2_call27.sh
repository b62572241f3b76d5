#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q -k "training or adam or neuralq" > gpurun_out/r2_pytest_dqn27.log 2>&1; echo "pytest dqn rc=$?"; tail -3 gpurun_out/r2_pytest_dqn27.log | cut -c1-1200
for f in 1; do
for w in cornell_neuralq; do
  RLPT_NQ_FUSED_BWD=$f timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_27.json 2> gpurun_out/r2_bench_${w}_27.err; echo "fused=$f $w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_27.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done; done
W="--workload cornell_neuralq --steps 1 --warmup 3 --width 128 --height 128 --batch 4096 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 3000 -c 200 --csv --log-file gpurun_out/r2_launches_nq_27.csv python bench.py $W > gpurun_out/r2_ncu_nq_27.log 2>&1

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dqn.py tests/test_gpu_host_mirror.py -m gpu -q -x > gpurun_out/r2_pytest_dqn8.log 2>&1; echo "pytest dqn rc=$?"; tail -12 gpurun_out/r2_pytest_dqn8.log | cut -c1-400
for w in cornell_neuralq archway_neuralq; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2_bench_${w}_c.json 2> gpurun_out/r2_bench_${w}_c.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_c.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'])"
done
W="--workload cornell_neuralq --steps 1 --warmup 3 --width 128 --height 128 --batch 4096"
python bench.py $W > gpurun_out/r2_plain_nq.json 2> gpurun_out/r2_plain_nq.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1200 --csv --log-file gpurun_out/r2_launches_nq_b.csv python bench.py $W > gpurun_out/r2_ncu_nq.log 2>&1
tail -2 gpurun_out/r2_ncu_nq.log

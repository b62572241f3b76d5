#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all9.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/r2_pytest_all9.log | cut -c1-400
for w in cornell_neuralq archway_neuralq; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2_bench_${w}_d.json 2> gpurun_out/r2_bench_${w}_d.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_d.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'])"
done
for w in cornell_sarsa door_room_sarsa archway_sarsa medieval_sarsa; do
  timeout 600 python bench.py --workload $w --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$w.json')); print({k:d[k] for k in ('value','ms_per_step','mean_path_length','radiance_map_build_ms')}, d['roofline']['frac'], d['roofline'].get('frac_exclusive'))"
done

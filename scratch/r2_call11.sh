#!/bin/bash
mkdir -p gpurun_out
for w in cornell_neuralq archway_neuralq; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2_bench_${w}_e.json 2> gpurun_out/r2_bench_${w}_e.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_e.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame')}, d['roofline']['frac'])"
done
W="--workload medieval_inside_default --steps 2 --warmup 3 --no-cpu-baseline --no-exclusive"
python bench.py $W > gpurun_out/r2_plain_med_inside_b.json 2> gpurun_out/r2_plain_med_inside_b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_med_inside_b.csv python bench.py $W > gpurun_out/r2_ncu_list_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_isect_bvh -s 160 -c 4 -f -o gpurun_out/r2_prof_bvh4_b python bench.py $W > gpurun_out/r2_ncu_full_b.log 2>&1; tail -2 gpurun_out/r2_ncu_full_b.log
W="--workload cornell_neuralq --steps 1 --warmup 3 --width 128 --height 128 --batch 4096"
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1000 --csv --log-file gpurun_out/r2_launches_nq_c.csv python bench.py $W > gpurun_out/r2_ncu_nq_c.log 2>&1

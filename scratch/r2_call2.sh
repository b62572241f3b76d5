#!/bin/bash
# round 2, call 2: full GPU suite on the BVH4 build + ncu of k_isect_bvh on the closed-scene workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all2.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/r2_pytest_all2.log
W="--workload medieval_inside_default --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py $W > gpurun_out/r2_plain_med_inside.json 2> gpurun_out/r2_plain_med_inside.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_med_inside.csv python bench.py $W > gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_isect_bvh -s 160 -c 6 -f -o gpurun_out/r2_prof_bvh4_a python bench.py $W > gpurun_out/r2_ncu_full.log 2>&1; tail -2 gpurun_out/r2_ncu_full.log

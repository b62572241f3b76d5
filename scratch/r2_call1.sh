#!/bin/bash
# round 2, call 1: BVH4 correctness + first numbers
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "closest_hit or bvh" > gpurun_out/r2_pytest_bvh.log 2>&1; echo "pytest bvh rc=$?"; tail -5 gpurun_out/r2_pytest_bvh.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/r2_pytest_all.log
for w in medieval_inside_default medieval_default archway_sarsa complex_light_room_default cornell_sarsa; do
  timeout 300 bash scratch/kstats.sh "X=1" --workload $w 2>&1 | tail -1
done
for l in 1 2; do RLPT_BVH_LEAF=$l timeout 300 bash scratch/kstats.sh "RLPT_BVH_LEAF=$l" --workload medieval_inside_default 2>&1 | tail -1; done

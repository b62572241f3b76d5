#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "closest_hit or bvh or through_the_bvh" > gpurun_out/r2_pytest_bvh3.log 2>&1; echo "pytest bvh rc=$?"; tail -3 gpurun_out/r2_pytest_bvh3.log
for lib in librlpt.so librlpt_b8r8.so librlpt_b2r4.so librlpt_b4r12.so librlpt_mb4.so librlpt_mb2.so; do
  RLPT_LIB_NAME=$lib timeout 300 bash scratch/kstats.sh "RLPT_LIB_NAME=$lib" --workload medieval_inside_default 2>&1 | tail -1
done
for w in medieval_default archway_sarsa; do timeout 300 bash scratch/kstats.sh "X=1" --workload $w 2>&1 | tail -1; done
for l in 1 2; do timeout 300 bash scratch/kstats.sh "RLPT_BVH_LEAF=$l" --workload medieval_inside_default 2>&1 | tail -1; done

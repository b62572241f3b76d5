#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dqn.py -m gpu -x -q -k "fused_backward" > gpurun_out/r2_pytest_dqn37.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_dqn37.log | cut -c1-1500

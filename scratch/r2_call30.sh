#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus or gpus_2 or 2gpu" > gpurun_out/r2_pytest_2gpu_30.log 2>&1; echo "pytest 2gpu rc=$?"; tail -5 gpurun_out/r2_pytest_2gpu_30.log | cut -c1-600

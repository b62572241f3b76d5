#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "two_gpus or host_mirror or offline_trainer or example_program" > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -15 gpurun_out/r2_pytest_2gpu.log | cut -c1-600

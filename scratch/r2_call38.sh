#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
REPS=3 BENCH_ARGS="--no-exclusive --steps 32" timeout 600 bash scratch/ab.sh "RLPT_LIB_NAME=librlpt_prev.so" "RLPT_LIB_NAME=librlpt.so"

#!/bin/bash
REPS=2 BENCH_ARGS="--no-exclusive --steps 32" timeout 900 bash scratch/ab.sh "RLPT_LIB_NAME=librlpt.so" "RLPT_LIB_NAME=librlpt_i7.so" "RLPT_LIB_NAME=librlpt_i8.so" "RLPT_LIB_NAME=librlpt_i5.so"

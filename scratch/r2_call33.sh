#!/bin/bash
for lib in librlpt.so librlpt_p2.so librlpt_p4.so; do
  RLPT_LIB_NAME=$lib timeout 300 python bench.py --workload cornell_neuralq --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$lib', {k:d[k] for k in ('ms_per_step','us_per_optimiser_step')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done

#!/bin/bash
# A/B on one box: scratch/ab.sh libA.so libB.so ...   (each built with RLPT_LIB_NAME=...); prints Mpaths/s per lib, interleaved reps
for rep in 1 2; do
  for lib in "$@"; do
    v=$(RLPT_LIB_NAME=$lib python bench.py --no-cpu-baseline --steps 32 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f ms=%.3f e2e=%.1f frac=%.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))")
    echo "$lib rep$rep: $v"
  done
done

#!/bin/bash
# A/B on one box: scratch/ab.sh "VAR=val VAR2=val" "VAR=val" ...  -- each argument is an environment for one bench variant
# (e.g. RLPT_LIB_NAME=librlpt_x.so RLPT_SPLIT=0); prints Mpaths/s per variant, two interleaved repetitions
REPS=${REPS:-2}
for rep in $(seq 1 $REPS); do
  for spec in "$@"; do
    v=$(env $spec python bench.py --no-cpu-baseline --steps 32 --warmup 3 ${BENCH_ARGS:-} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f ms=%.3f e2e=%.1f frac=%.3f launches=%d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']))")
    echo "[$spec] rep$rep: $v"
  done
done

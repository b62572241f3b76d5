#!/bin/bash
# scratch/abw.sh WORKLOAD "ENV=.." "ENV=.." ...: one short bench per environment on the given workload
w=$1; shift
for spec in "$@"; do
  v=$(env $spec timeout 300 python bench.py --no-cpu-baseline --workload $w --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f ms=%.3f frac=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))")
  echo "[$w $spec]: $v"
done

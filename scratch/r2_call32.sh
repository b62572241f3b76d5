#!/bin/bash
REPS=1 BENCH_ARGS="--workload door_room_sarsa --no-exclusive --steps 16" timeout 900 bash scratch/ab.sh "RLPT_TAIL=8192" "RLPT_TAIL=16384" "RLPT_TAIL=32768" "RLPT_TAIL=65536" "RLPT_TAIL=131072"
python bench.py --workload door_room_sarsa --no-exclusive --no-cpu-baseline --steps 16 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','mean_path_length','gpu_launches')}); print('tail', d.get('tail')); print('isect', d['roofline'].get('share_of_kernel_time'), 'shade', d['roofline_shade'].get('share_of_kernel_time'))"

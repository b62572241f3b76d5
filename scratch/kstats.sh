#!/bin/bash
# scratch/kstats.sh "VAR=val ..." [bench args]: one bench run, per-kernel device times from the JSON line
spec="$1"; shift
env $spec python bench.py --no-cpu-baseline --steps 8 --warmup 3 "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r, s, t = d['roofline'], d['roofline_shade'], d['tail']
fr = d['steps']
rays = d['mean_path_length'] * 512 * 512 * d['config']['spp_per_frame'] * fr
print('[$spec $*] %.1f Mpaths/s %.3f ms/frame | per frame: isect %.2f ms (%d launches) shade %.2f ms tail %.2f ms (%d) | per ray: %.1f box %.1f tri | frac %.3f' % (
  d['value'], d['ms_per_step'], r['avg_launch_ms'] * r['launches'] / fr, r['launches'] / fr, s['avg_launch_ms'] * s['launches'] / fr, t['seconds'] * 1e3 / fr, t['launches'] / fr,
  r['box_tests'] / rays, r['triangle_tests'] / rays, r['frac']))"

#!/bin/bash
mkdir -p gpurun_out
n=8
run() { tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $n --steps 48 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_ab_$tag.json 2> gpurun_out/r2_scale_ab_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_scale_ab_$tag.json')); print('$tag', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value']), 'merge GB/s', round(d['roofline_merge']['achieved']), d['clocks'].get('samples'))"; }
run nosmi RLPT_X=1
run smi RLPT_BENCH_SMI=1

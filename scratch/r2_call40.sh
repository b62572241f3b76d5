#!/bin/bash
mkdir -p gpurun_out
n=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $n --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_final_cornell_weak_n$n.json 2> gpurun_out/r2_scale_final_cornell_weak_n$n.err
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_scale_final_cornell_weak_n$n.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['roofline_merge']['achieved'], d['clocks'])"

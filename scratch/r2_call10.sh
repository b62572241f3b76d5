#!/bin/bash
# 8-GPU box: scaling lines (weak + strong) for the configs that name multi-GPU, and the 2-GPU Neural-Q test
mkdir -p gpurun_out
run() { n=$1; shift; tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 32 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2_scale_${tag}_n$n.json 2> gpurun_out/r2_scale_${tag}_n$n.err
  echo "$tag n=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_scale_${tag}_n$n.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['config']['exchange'][:40])"; }
run 8 cornell_weak --workload cornell_sarsa
run 8 cornell_strong --workload cornell_sarsa --scaling strong --spp 32
run 8 door_room_weak --workload door_room_sarsa
run 8 medieval_weak --workload medieval_sarsa
run 4 cornell_weak --workload cornell_sarsa
run 4 cornell_strong --workload cornell_sarsa --scaling strong --spp 32
run 2 cornell_strong --workload cornell_sarsa --scaling strong --spp 32
timeout 300 python -m pytest tests/test_gpu_dqn.py -m gpu -q -k two_gpus > gpurun_out/r2_pytest_nq2gpu.log 2>&1; echo "pytest nq 2gpu rc=$?"; tail -4 gpurun_out/r2_pytest_nq2gpu.log | cut -c1-300

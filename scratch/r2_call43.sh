#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final2.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2_pytest_final2.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final2_cornell_sarsa.json 2> gpurun_out/r2_bench_final2.err; echo "bench default rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final2_cornell_sarsa.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_exclusive'), d['cpu_baseline']['value'], d['clocks'])"

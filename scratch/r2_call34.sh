#!/bin/bash
# end-of-round validation + captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2_pytest_final.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final_cornell_sarsa.json 2> gpurun_out/r2_bench_final.err; echo "bench default rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final_cornell_sarsa.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_exclusive'), d['cpu_baseline']['value'])"
for w in cornell_neuralq archway_neuralq; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final_${w}.json 2> gpurun_out/r2_bench_final_${w}.err; echo "$w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final_${w}.json')); print({k:d[k] for k in ('value','ms_per_step','us_per_optimiser_step','train_share_of_frame','gpu_launches')}, d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done
bash scratch/profile_cornell.sh r2final
W="--workload archway_neuralq --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dqn_forward|k_dqn_backward' -c 3 -f -o gpurun_out/r2_prof_dqn_final python bench.py $W > gpurun_out/r2_ncu_dqn_final.log 2>&1; tail -1 gpurun_out/r2_ncu_dqn_final.log
W="--workload archway_neuralq --steps 1 --warmup 3 --width 128 --height 128 --batch 4096 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/r2_launches_nq_final.csv python bench.py $W > gpurun_out/r2_ncu_nq_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 3000 -c 300 --csv --log-file gpurun_out/r2_launches_nq_final_warm.csv python bench.py $W > gpurun_out/r2_ncu_nq_final_warm.log 2>&1
echo done

#!/bin/bash
# scratch/full_check.sh TAG: the whole GPU suite, smoke, the default bench line, per-workload kernel breakdowns, and the ncu launch
# list + one --set full capture of the longest secondary k_isect_bvh launch (Medieval_House); everything lands in gpurun_out/
T=${1:-x}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$T.log 2>&1; tail -3 gpurun_out/pytest_gpu_$T.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$T.log 2>&1; tail -1 gpurun_out/smoke_$T.log
python bench.py --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
for w in medieval_default medieval_sarsa archway_sarsa door_room_sarsa cornell_sarsa cornell_default; do bash scratch/kstats.sh "A=1" --workload $w; done > gpurun_out/kstats_$T.log 2>&1; cat gpurun_out/kstats_$T.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
ncu --metrics $M --clock-control none -k regex:k_isect_bvh -c 24 --csv --log-file gpurun_out/launches_bvh_$T.csv python bench.py --workload medieval_default --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bvh_list_$T.log 2>&1
IDX=$(python scratch/ncu_pick.py gpurun_out/launches_bvh_$T.csv k_isect_bvh "0, 0>"); echo idx $IDX
ncu --set full --clock-control none --import-source on -k regex:k_isect_bvh -s $IDX -c 1 -f -o gpurun_out/prof_bvh_$T python bench.py --workload medieval_default --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bvh_full_$T.log 2>&1; tail -2 gpurun_out/ncu_bvh_full_$T.log

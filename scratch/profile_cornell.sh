#!/bin/bash
# scratch/profile_cornell.sh TAG: the headline workload under ncu -- (1) launch list of `bench.py --steps 2 --warmup 3` (durations only),
# (2) one --set full capture of eight consecutive k_isect / k_shade launches starting at the first primary k_isect of the fourth frame
# (both lanes' primary and bounce-1 launches). Numbers printed by bench.py under ncu are not bench values.
T=${1:-x}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cornell_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cornell_list_$T.log 2>&1
IDX=$(python - <<P
import csv, re
rows = list(csv.reader(l for l in open("gpurun_out/launches_cornell_$T.csv") if l.startswith('"')))
h = rows[0]; ni = h.index("Kernel Name")
names = [r[ni] for r in rows[1:] if re.search("k_isect|k_shade", r[ni])]
prim = [i for i, n in enumerate(names) if "k_isect<1, 1>" in n]
print(prim[6] if len(prim) > 6 else prim[-2])
P
)
echo "full capture from matching launch $IDX"
ncu --set full --clock-control none --import-source on -k regex:"k_isect|k_shade" -s $IDX -c 8 -f -o gpurun_out/prof_cornell_$T python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cornell_full_$T.log 2>&1; tail -2 gpurun_out/ncu_cornell_full_$T.log

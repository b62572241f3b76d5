#!/usr/bin/env python
"""scratch/ncu_pick.py launches.csv regex [secondary]: index (among the launches whose name matches regex) of the longest one --
feeds `ncu -k regex:... -s IDX -c 1` so that the --set full capture lands on the launch that matters."""
import csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ni = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value")
pat = re.compile(sys.argv[2]); want = sys.argv[3] if len(sys.argv) > 3 else ""
k = -1; best = (-1.0, 0)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum" or not pat.search(r[ni]): continue
    k += 1
    if want and want not in r[ni]: continue
    v = float(r[vi].replace(",", ""))
    if v > best[0]: best = (v, k)
print(best[1])

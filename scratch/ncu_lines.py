#!/usr/bin/env python
"""scratch/ncu_lines.py rep.ncu-rep 'kernel substring' [top]: warp instructions, active lanes and stall samples per SOURCE line of one
kernel of an ncu --set full --import-source on capture (the per-line view of profiles/ncu_summary.py's SASS regions)."""
import csv, subprocess, sys, io
rep, want = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = file = hdr = None; out = {}
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": file = r[1]; continue
    if r[0] in ("Function Name", "Kernel Name"): fn = r[1]; continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr and fn and want in fn and r[0].isdigit() and len(r) >= len(hdr) - 2:
        off = len(r) - len(hdr)
        o = out.setdefault((file.split('/')[-1], int(r[0])), [0, 0, 0, r[1].strip()[:100]])
        o[0] += num(r[ix["Instructions Executed"] + off]); o[1] += num(r[ix["Thread Instructions Executed"] + off]); o[2] += num(r[ix["# Samples"] + off])
tot = sum(v[0] for v in out.values()); ts = sum(v[2] for v in out.values())
print("total warp inst", tot, "samples", ts)
for k, v in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s:%4d inst=%6.2fM (%4.1f%%) lanes=%4.1f samp=%4.1f%% | %s" % (k[0], k[1], v[0] / 1e6, 100 * v[0] / max(tot, 1), v[1] / max(v[0], 1), 100 * v[2] / max(ts, 1), v[3]))

#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics + hot SASS regions. usage: ncu_summary.py rep [kernel_index]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'sm__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum', 'sm__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want or h in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        print("%-75s %-12s %s" % (h, units[i], [r[i] for r in rows[2:]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernels = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {"name": r[1], "rows": []}; kernels.append(cur)
    elif r and r[0] == 'Address': cur["hdr"] = r
    elif cur is not None and len(r) > 10: cur["rows"].append(r)
k = kernels[kidx]; ix = {h: i for i, h in enumerate(k["hdr"])}; data = k["rows"]
print("\n== hot SASS regions of", k["name"][:80], "kernel", kidx)
tot = sum(int(r[ix['Instructions Executed']]) for r in data); samp = sum(int(r[ix['# Samples']]) for r in data)
print("total warp inst", tot, "samples", samp)
out = [(i, int(r[ix['Instructions Executed']]), r[ix['Avg. Threads Executed']], int(r[ix['# Samples']]), r[1].strip()) for i, r in enumerate(data)]
stalls = [h for h in k["hdr"] if h.startswith('stall_') and 'Not Issued' not in h]
start = 0
for i in range(1, len(out) + 1):
    if i == len(out) or abs(out[i][1] - out[start][1]) > 0.15 * max(out[start][1], 1) + 1000:
        c = sum(x[1] for x in out[start:i]); s = sum(x[3] for x in out[start:i])
        if c > 0.01 * tot or s > 0.01 * samp:
            ops = collections.Counter((x[4].split()[1] if x[4].startswith('@') else x[4].split()[0]) for x in out[start:i]).most_common(5)
            st = collections.Counter()
            for r in data[start:i]:
                for h in stalls: st[h] += int(r[ix[h]] or 0)
            print("%4d-%4d n=%3d inst=%7.1fM (%4.1f%%) samp=%4.1f%% thr=%s %s | %s" % (start, i - 1, i - start, c / 1e6, 100 * c / tot, 100 * s / samp, out[start][2], ops, st.most_common(3)))
        start = i

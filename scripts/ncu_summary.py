#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key counters + top stalled SASS instructions."""
import csv, subprocess, sys, io

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %-12s %s" % (w, units[i], [r[i] for r in data]))
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
        v = [float(r[i].replace(',', '')) for r in data]
        if max(v) > 0.05:
            print("stall %-40s %s" % (h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
data = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ci = {n: i for i, n in enumerate(h)}
tot = sum(float(r[ci['# Samples']] or 0) for r in data)
texec = sum(float(r[ci['Instructions Executed']] or 0) for r in data)
print("samples", tot, "sass instructions", len(data), "warp-instructions executed", texec)
for r in sorted(data, key=lambda r: -float(r[ci['# Samples']] or 0))[:topn]:
    s = float(r[ci['# Samples']])
    st = {k[6:]: int(r[ci[k]]) for k in ci if k.startswith('stall_') and '(' not in k and r[ci[k]] not in ('', '0')}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print('%5.1f%% %-60s exec=%-9s %s' % (100 * s / tot, r[ci['Source']][:60], r[ci['Instructions Executed']], top))

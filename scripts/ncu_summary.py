#!/usr/bin/env python
"""Summarise an .ncu-rep (raw metrics + SASS-level executed-instruction histogram).  Usage: ncu_summary.py rep [--sass]"""
import csv, io, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pat = re.compile(r'Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__warps_active.avg.pct|launch__(grid_size|block_size|registers_per_thread$|occupancy_limit|waves)|smsp__inst_executed.sum$|issue_active.avg.pct|smsp__average_warps_issue_stalled.*_per_issue_active|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|sm__inst_executed_pipe_(lsu|alu|fma|xu|uniform).*pct_of_peak_sustained_active|lts__t_sector_hit_rate.pct|launch__shared_mem_per_block')
for r in rows[2:]:
    for k, u, v in zip(hdr, units, r):
        if pat.search(k):
            try:
                fv = float(v.replace(',', ''))
                if 'stalled' in k and fv < 0.05: continue
            except ValueError:
                pass
            print(f"{k} = {v} {u}")
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc, iex = hdr.index('Source'), hdr.index('Instructions Executed')
tot = sum(int(r[iex]) for r in data)
warps = int(data[0][iex])
print('static SASS', len(data), 'executed warp-inst', tot, 'per warp', round(tot / warps, 1))
h = Counter()
for r in data:
    op = r[isrc].strip().split()
    if op and op[0].startswith('@'): op = op[1:]
    h[op[0].split('.')[0] if op else '?'] += int(r[iex])
print(' '.join(f"{k}:{v/warps:.0f}" for k, v in h.most_common(30)))
if '--sass' in sys.argv:
    for i, r in enumerate(data): print(i, r[iex], r[isrc])

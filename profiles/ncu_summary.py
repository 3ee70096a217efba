#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key metrics + opcode histogram.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.sum']
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print("kernel:", r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
    for k in keys + stall:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = None
c, s = Counter(), Counter()
tot = 0
for r in srows:
    if 'Source' in r and 'Instructions Executed' in r:
        h = r; continue
    if h is None or len(r) != len(h): continue
    ex = r[h.index('Instructions Executed')]
    if not ex.isdigit(): continue
    txt = r[h.index('Source')].split()
    op = (txt[1] if txt[0].startswith('@') else txt[0]).split('.')[0]
    c[op] += int(ex); tot += int(ex)
    sm = r[h.index('# Samples')]
    s[op] += int(sm) if sm.isdigit() else 0
print("opcode histogram (warp instructions executed, stall samples):")
for op, n in c.most_common(25):
    print(f"  {op:12s} {n:12d} {100*n/max(tot,1):5.1f}%  samples {s[op]}")

#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel key metrics + top stall SASS lines.  usage: ncu_summary.py rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'launch__grid_size']
for r in rows[2:]:
    print('---')
    for w in want:
        if w in hdr:
            print(f"  {w:62s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur = None; data = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'Kernel Name':
        if cur is not None: data.append(cur)
        cur = [r[1], None, []]; continue
    if cur is None: continue
    if r and r[0] == 'Address': cur[1] = r; continue
    if cur[1] and len(r) == len(cur[1]): cur[2].append(r)
if cur is not None: data.append(cur)
for name, h, rs in data:
    si = h.index('# Samples'); sc = h.index('Source')
    tot = sum(int(x[si] or 0) for x in rs) or 1
    print(f"=== {name}  total samples {tot}")
    for idx, x in sorted(enumerate(rs), key=lambda ix: -int(ix[1][si] or 0))[:topn]:
        print(f"   {int(x[si]):7d} {100*int(x[si])/tot:5.1f}%  #{idx:4d} {x[sc].strip()[:100]}")

#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + top stalled SASS instructions (needs ncu on PATH).
usage: tools/ncu_top.py report.ncu-rep [kernel-index] [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kid = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; r = rows[2 + kid]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size']
for w in want:
    if w in hdr: print(f"{w:70s} {r[hdr.index(w)][:60]} {rows[1][hdr.index(w)]}")
st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in h and r[i].replace('.', '').isdigit()]
tot = sum(v for v, _ in st) or 1
print("stall samples:", ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')}={100 * v / tot:.0f}%" for v, h in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page concatenates kernels: split on "Kernel Name" lines
blocks = src.split('"Kernel Name"')[1:]
blk = blocks[min(kid, len(blocks) - 1)]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
hdr = rows[1]; ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
data = [(int(x[isamp]), x[ia].strip(), int(x[iex]), i) for i, x in enumerate(rows[2:]) if len(x) > isamp and x[isamp].isdigit()]
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
for s, text, ex, i in sorted(data, reverse=True)[:topn]:
    print(f"{s:7d} {100 * s / tot:5.1f}%  #{i:5d} ex={ex:10d}  {text[:100]}")

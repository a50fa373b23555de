#!/usr/bin/env python
"""Aggregate an .ncu-rep's warp-stall samples by CUDA source line (needs -lineinfo builds).
usage: tools/ncu_lines.py report.ncu-rep [kernel-substring] [topN]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; ksub = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict(); cur_file = None; hdr = None; func = None; seen_funcs = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or (ksub and ksub not in (func or "")): continue
    if func not in seen_funcs: seen_funcs.append(func)
    if func != seen_funcs[0]: continue           # first matching kernel instance only
    ln = r[0]
    if not ln.isdigit():  # SASS rows have an empty line-number column; CUDA line rows carry the per-line totals
        continue
    try:
        # index from the end: source text with embedded quotes can shift the leading columns
        isamp = hdr.index("# Samples") - len(hdr); iex = hdr.index("Instructions Executed") - len(hdr)
        s = int(r[isamp] or 0); ex = int(r[iex] or 0)
    except (ValueError, IndexError):
        continue
    key = (cur_file, ln, r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += ex
tot = sum(v[0] for v in agg.values()) or 1; totex = sum(v[1] for v in agg.values()) or 1
print("kernel:", (seen_funcs[0] if seen_funcs else "?")[:100]); print("samples", tot, "instructions", totex)
for (f, ln, src), (s, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*s/tot:5.1f}% smp {100*ex/totex:5.1f}% ins  {f}:{ln:>4s}  {src}")

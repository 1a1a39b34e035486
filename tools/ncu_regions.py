#!/usr/bin/env python
"""Developer tool: split the SASS of an .ncu-rep's first kernel into runs of equal execution count and print the share of
executed instructions and stall samples of each run (which loop costs what).  Usage: ncu_regions.py report.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, iall, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name' and data:
        break
    if len(r) < len(hdr) or r[0] in ('Kernel Name', 'Address'):
        continue
    data.append((r[isrc].strip(), int(r[iex]), int(r[iall])))
tot = sum(d[1] for d in data) or 1
stot = sum(d[2] for d in data) or 1
runs, cur = [], None
for i, (s, c, st) in enumerate(data):
    if cur and abs(c - cur[2]) <= 0.03 * max(c, cur[2], 1):
        cur[1] = i; cur[3] += c; cur[4] += st
    else:
        if cur:
            runs.append(cur)
        cur = [i, i, c, c, st, s]
runs.append(cur)
print(f"{len(data)} SASS instructions, {tot} warp-instructions executed, {stot} stall samples")
for a, b, c, t, st, s in runs:
    if 100.0 * t / tot >= min_pct or 100.0 * st / stot >= min_pct:
        print(f"{a:5d}-{b:5d} n={b - a + 1:4d} exec/line={c:9d} inst={100 * t / tot:5.1f}% stalls={100 * st / stot:5.1f}%  {s[:60]}")

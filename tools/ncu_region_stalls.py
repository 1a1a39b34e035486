#!/usr/bin/env python
"""Developer tool: per SASS region (runs of equal execution count) of an .ncu-rep's first kernel, the share of executed
instructions and the stall samples by reason, plus shared-memory wavefronts (ideal / actual).
Usage: ncu_region_stalls.py report.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
reasons = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name' and data:
        break
    if len(r) < len(hdr) or r[0] in ('Kernel Name', 'Address'):
        continue
    data.append(r)


def num(r, k):
    try:
        return float(r[ix[k]] or 0)
    except ValueError:
        return 0.0


tot = sum(num(r, 'Instructions Executed') for r in data) or 1
stot = sum(num(r, 'Warp Stall Sampling (All Samples)') for r in data) or 1
runs, cur = [], None
for i, r in enumerate(data):
    c = num(r, 'Instructions Executed')
    if cur and abs(c - cur['c']) <= 0.03 * max(c, cur['c'], 1):
        cur['b'] = i
    else:
        if cur:
            runs.append(cur)
        cur = {'a': i, 'b': i, 'c': c}
runs.append(cur)
print(f"{len(data)} SASS instructions, {tot:.0f} warp-instructions, {stot:.0f} stall samples")
for run in runs:
    seg = data[run['a']:run['b'] + 1]
    inst = sum(num(r, 'Instructions Executed') for r in seg)
    st = sum(num(r, 'Warp Stall Sampling (All Samples)') for r in seg)
    if 100 * inst / tot < min_pct and 100 * st / stot < min_pct:
        continue
    by = sorted(((sum(num(r, k) for r in seg), k.replace('stall_', '')) for k in reasons), reverse=True)[:5]
    wf, wfi = sum(num(r, 'L1 Wavefronts Shared') for r in seg), sum(num(r, 'L1 Wavefronts Shared Ideal') for r in seg)
    thr = sum(num(r, 'Thread Instructions Executed') for r in seg)
    print(f"{run['a']:5d}-{run['b']:5d} n={run['b'] - run['a'] + 1:4d} x{run['c']:9.0f} inst={100 * inst / tot:5.1f}% lanes={thr / max(inst, 1):4.1f} "
          f"stalls={100 * st / stot:5.1f}% smem_wf={wf / 1e6:6.2f}M (ideal {wfi / 1e6:6.2f}M) | "
          + ' '.join(f"{k}={100 * v / max(st, 1):.0f}%" for v, k in by if v > 0))

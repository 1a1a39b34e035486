#!/usr/bin/env python
"""Developer tool: list the loops (backward branches) of one kernel in a built library with their instruction count and
opcode mix.  Usage: sass_loops.py lib.so kernel-name-substring [min_instructions]"""
import collections
import re
import subprocess
import sys

lib, name = sys.argv[1], sys.argv[2]
min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
fn, rows = None, []
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m and fn and name in fn:
        rows.append((int(m.group(1), 16), m.group(2).strip(), fn))
fns = sorted(set(r[2] for r in rows))
for f in fns:
    ins = [(a, t) for a, t, g in rows if g == f]
    addr = {a: i for i, (a, t) in enumerate(ins)}
    print(f, len(ins), 'instructions')
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA(?:\.U)?\s+(?:\S+,\s*)?(0x[0-9a-f]+)', t)
        if m and int(m.group(1), 16) in addr and int(m.group(1), 16) <= a:
            j = addr[int(m.group(1), 16)]
            n = i - j + 1
            if n < min_n:
                continue
            c = collections.Counter()
            for _, u in ins[j:i + 1]:
                w = u.split()
                op = w[1] if w[0].startswith('@') else w[0]
                c[op.split('.')[0]] += 1
            print(f'  loop {ins[j][0]:#x}..{a:#x}: {n} instructions', dict(c.most_common(14)))

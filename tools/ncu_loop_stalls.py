#!/usr/bin/env python
"""Where a kernel's hot loop waits: `ncu -i X.ncu-rep --page source --csv --print-source sass > X.csv`,
then `python tools/ncu_loop_stalls.py X.csv [min_executions]` prints the stall mix of the instructions
executed at least that often and the most-sampled ones."""
import csv, sys
name = sys.argv[1]
least = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
rows = list(csv.reader(open(name)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    try:
        a = int(r[idx['Address']], 16)
    except Exception:
        continue
    data.append((a, r[idx['Source']], int(r[idx['# Samples']] or 0), int(r[idx['Instructions Executed']] or 0), r))
base = data[0][0]
print(name, 'total samples', sum(d[2] for d in data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
loop = [d for d in data if d[3] >= least]
print('loop instrs', len(loop), 'samples', sum(d[2] for d in loop), 'executions', loop[0][3] if loop else None)
agg = {}
for d in loop:
    for c in stall_cols:
        agg[c] = agg.get(c, 0) + int(d[4][idx[c]] or 0)
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for d in sorted(loop, key=lambda x: -x[2])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    st = sorted(((int(d[4][idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(hex(d[0] - base), d[2], d[1][:70], st)

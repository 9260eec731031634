#!/usr/bin/env python
"""Group a tools/ncu_sass_dump.py listing into runs of equal execution frequency.
usage: tools/sass_regions.py counts.txt <units> [min_instr_per_unit]   (units = e.g. warp-tiles, to normalise counts)"""
import sys
rows = []
for ln in open(sys.argv[1]):
    if ln.startswith('=='): print(ln.strip()); continue
    p = ln.split(None, 4)
    rows.append((int(p[0]), int(p[1]), p[2], int(p[3]), p[4].strip()))
WT = float(sys.argv[2]); lo = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
tot = sum(r[1] for r in rows); stt = sum(r[3] for r in rows) or 1
print("total warp instr", tot, "per unit", tot / WT)
reg = []; cur = None
for r in rows:
    f = r[1] / WT
    if cur and abs(cur['f'] - f) < 0.02 * max(cur['f'], 0.05):
        cur['n'] += 1; cur['sum'] += f; cur['end'] = r[0]; cur['st'] += r[3]
    else:
        cur = {'start': r[0], 'end': r[0], 'f': f, 'n': 1, 'sum': f, 'st': r[3]}; reg.append(cur)
for c in reg:
    if c['sum'] > lo:
        print(f"#{c['start']:4d}-{c['end']:4d} n={c['n']:4d} freq={c['f']:.3f} instr/unit={c['sum']:.1f} stall%={100*c['st']/stt:.1f}")

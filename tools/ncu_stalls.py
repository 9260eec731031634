#!/usr/bin/env python
"""Stall reasons per opcode over the hottest loop of a kernel (instructions executed as often as the most common count).
usage: tools/ncu_stalls.py file.ncu-rep"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = None
recs = []
for r in rows:
    if r and r[0] == 'Address': H = r; continue
    if not H or len(r) < len(H): continue
    try: ie = int(r[H.index('Instructions Executed')])
    except ValueError: continue
    recs.append((ie, r))
w = collections.Counter()
for ie, r in recs:
    if ie > 0: w[ie] += int(r[H.index('Warp Stall Sampling (All Samples)')])
common = w.most_common(1)[0][0]   # the execution count that collects the most stall samples = the hot loop
agg = collections.defaultdict(collections.Counter); cnt = collections.Counter()
KEYS = ['stall_math', 'stall_mio', 'stall_short_sb', 'stall_long_sb', 'stall_wait', 'stall_selected', 'stall_not_selected', 'stall_dispatch',
        'stall_sleep', 'stall_branch_resolving', 'stall_no_inst', 'stall_lg', 'stall_barrier']
for ie, r in recs:
    if ie != common: continue
    op = [o for o in r[1].strip().split() if not o.startswith('@')][0].split('.')[0]
    cnt[op] += 1
    for k in KEYS: agg[op][k] += int(r[H.index(k)])
tot = sum(sum(v.values()) for v in agg.values())
print(f'loop instructions executed {common} times each: {sum(cnt.values())} instructions, {tot} stall samples')
for op, v in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
    s = sum(v.values()) or 1
    print(f'{op:8s} n={cnt[op]:4d} {100*s/tot:5.1f}%  ' + ' '.join(f'{k[6:]}={100*x/s:.0f}%' for k, x in v.most_common(5)))

#!/usr/bin/env python
"""Summarise an .ncu-rep source page: hottest SASS instructions by stall samples / executed count.
usage: tools/ncu_top.py file.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; H = None; data = []
def flush():
    if not data: return
    tot = sum(d[0] for d in data) or 1; ti = sum(d[1] for d in data)
    print(f"== {kern}: {len(data)} SASS instrs, {ti} warp-instr executed, {tot} stall samples")
    for i, d in enumerate(data):
        d.append(i)
    for d in sorted(data, key=lambda d: -d[0])[:N]:
        print(f"{100*d[0]/tot:5.1f}%  exec={d[1]:>12} thr/inst={d[2]:>5}  #{d[4]:<4} {d[3]}")
for r in rows:
    if r and r[0] == "Kernel Name":
        flush(); data = []; kern = r[1][:90]; H = None; continue
    if r and r[0] == "Address":
        H = r; si = H.index("Source"); ws = H.index("Warp Stall Sampling (All Samples)"); ie = H.index("Instructions Executed"); at = H.index("Avg. Threads Executed"); continue
    if H and len(r) > ie:
        try: data.append([int(r[ws]), int(r[ie]), r[at], r[si].strip()[:100]])
        except ValueError: pass
    if kern and len(data) and False: pass
flush()

#!/usr/bin/env python
"""Dump every SASS instruction of the profiled kernels with its executed count, active threads and stall samples.
usage: tools/ncu_sass_dump.py file.ncu-rep out.txt"""
import csv, subprocess, sys
rep, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
H = None
with open(dst, "w") as f:
    idx = 0
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            f.write(f"== {r[1][:120]}\n"); H = None; idx = 0; continue
        if r and r[0] == "Address":
            H = r; si = H.index("Source"); ws = H.index("Warp Stall Sampling (All Samples)")
            ie = H.index("Instructions Executed"); at = H.index("Avg. Threads Executed"); continue
        if H and len(r) > ie:
            try:
                f.write(f"{idx:5d} {int(r[ie]):>12d} {r[at]:>5} {int(r[ws]):>8d}  {r[si].strip()[:110]}\n"); idx += 1
            except ValueError:
                pass

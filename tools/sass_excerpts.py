#!/usr/bin/env python
"""SASS excerpts of the shipped libtcsdn.so around its Blackwell-only instructions (runs here: cuobjdump needs no GPU).
usage: tools/sass_excerpts.py > profiles/<tag>_sass_excerpts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "traffic_classifier_sdn_b200", "libtcsdn.so")], capture_output=True, text=True).stdout
parts = re.split(r'\n\s*Function : ', txt)
out = ["# SASS excerpts of the shipped libtcsdn.so (cuobjdump -sass): the Blackwell-only instructions of the distance engine and the scorers",
       "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UBLKCP = cp.async.bulk (TMA unit), UTCBAR = tcgen05.commit,",
       "# UTCATOMSWS = tcgen05.alloc, SYNCS = mbarrier ops, MUFU.EX2 = ex2.approx, FFMA2 = fma.rn.f32x2.",
       "# Per kernel: instruction counts, then the first three sites of each kind with +-3 instructions of context.", ""]
WANT = ("engine_kernelIfLb1ELi5ELb0", "engine_kernelIfLb0ELi1", "scorer_tiled_kernelIfLi8ELi6ELi2ELi128ELi2", "scorer_tiled_kernelIfLi12ELi6ELi0ELi256ELi2")
for p in parts[1:]:
    name = p.split('\n', 1)[0].strip()
    if not any(w in name for w in WANT):
        continue
    ins = [m.group(2).strip() for m in (re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l) for l in p.split('\n')) if m]
    c = collections.Counter([t for t in x.split() if not t.startswith('@')][0].split('.')[0] for x in ins)
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    out.append(f"== {dem}")
    out.append(f"   {len(ins)} instructions; " + ", ".join(f"{k} {c[k]}" for k in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU", "FFMA2", "FMNMX3", "DADD", "DFMA", "STL", "LDL") if c[k]))
    keys = ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "UTCATOMSWS")
    shown = set(); n_sites = collections.Counter()
    for i, x in enumerate(ins):
        k = next((k for k in keys if k in x), None)
        if k is None:
            continue
        n_sites[k] += 1
        if n_sites[k] > 3:
            continue
        for j in range(max(0, i - 3), min(len(ins), i + 4)):
            if j not in shown:
                out.append(f"   {j:5d}  {ins[j]}")
                shown.add(j)
        out.append("   ...")
    out.append("")
sys.stdout.write("\n".join(out) + "\n")

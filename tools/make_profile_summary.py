#!/usr/bin/env python
"""Turn gpurun_out/ ncu captures into the committed profiles/ summaries (text + json).
usage: tools/make_profile_summary.py <tag>"""
import csv, json, os, subprocess, sys, collections
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, sys.argv[2] if len(sys.argv) > 2 else "profiles")
os.makedirs(P, exist_ok=True)
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct']
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
summary = {}
lines = [f"# ncu summaries, round {tag} (B200, --clock-control none; from gpurun_out/prof_{tag}_*.ncu-rep)\n"]
for name in ("gnb", "logistic", "forest", "forest_hbm", "forest_hbm2", "svc", "svc_refine", "knn"):
    rep = os.path.join(G, f"prof_{tag}_{name}.ncu-rep")
    if not os.path.exists(rep): continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); H = rows[0]; U = rows[1]; r = rows[2]
    kern = r[H.index("Kernel Name")]
    lines.append(f"\n## {name}: `{kern[:110]}`\n")
    d = {}
    for k in KEYS:
        if k in H:
            v, u = r[H.index(k)], U[H.index(k)]
            lines.append(f"    {k:72s} {v:>16} {u}")
            try: d[k] = float(v.replace(",", "")) * UNIT.get(u, 1)
            except ValueError: pass
    summary[name] = {"kernel": kern, "duration_s": d.get('gpu__time_duration.sum'),
                     "dram_bytes": (d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)),
                     "dram_read_bytes": d.get('dram__bytes_read.sum'), "dram_write_bytes": d.get('dram__bytes_write.sum'),
                     "tensor_pipe_pct": d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                     "xu_pipe_pct": d.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
                     "fp64_pipe_pct": d.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
                     "fma_pipe_pct": d.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
                     "alu_pipe_pct": d.get('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'),
                     "dram_pct_of_peak": d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                     "issue_active_pct": d.get('smsp__issue_active.avg.pct_of_peak_sustained_active')}
    # hottest SASS by stall samples
    top = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_top.py"), rep, "12"], capture_output=True, text=True).stdout
    lines.append("\n  hottest SASS instructions (share of warp-stall samples):\n")
    lines += ["    " + l for l in top.splitlines()[:14]]
    # SASS mnemonics that prove the Blackwell paths
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    mn = collections.Counter()
    for l in src.splitlines():
        for m in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "MUFU.EX2", "SYNCS", "DFMA", "FMNMX3", "UTMALDG"):
            if m in l: mn[m] += 1
    lines.append(f"\n  SASS mnemonics present: {dict(mn)}")
open(os.path.join(P, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump(summary, open(os.path.join(P, f"{tag}_ncu_summary.json"), "w"), indent=1)
# launch list
lf = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lf):
    rows = list(csv.reader(open(lf)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi: continue
        k = r[ki][:100]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as fh:
        fh.write(f"# every kernel launch of `python bench.py --steps 5 --warmup 3 --gpu-only` under ncu (gpu__time_duration.sum, cold-cache, serialised)\n\n")
        fh.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {c} | {t/1e3:.1f} | {100*t/tot:.1f}% |\n")
    import shutil; shutil.copy(lf, os.path.join(P, f"{tag}_launches.csv"))
print(open(os.path.join(P, f"{tag}_ncu_summary.md")).read()[:200]); print(json.dumps(summary, indent=0)[:1500])

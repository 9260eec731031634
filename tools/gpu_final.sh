#!/bin/bash
# final numbers of the round: GPU suite, smoke, bench line, reference arm, full launch list
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1800 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}.stderr; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2>/dev/null; echo "ref rc=$?"
if [ "${2:-launches}" = launches ]; then
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --gpu-only > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_${TAG}.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    a = agg.setdefault(r[ki][:110], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/${TAG}_launches.md", "w") as fh:
    fh.write("# every kernel launch of \`python bench.py --steps 5 --warmup 3 --gpu-only\` under ncu (gpu__time_duration.sum, cold-cache, serialised)\n\n")
    fh.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| \`{k}\` | {c} | {t/1e3:.1f} | {100*t/tot:.1f}% |\n")
print(open("gpurun_out/${TAG}_launches.md").read()[:3000])
PY
fi
timeout 120 tools/tmem_probe > gpurun_out/${TAG}_probe.txt 2>&1; tail -3 gpurun_out/${TAG}_probe.txt

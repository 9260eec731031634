#!/bin/bash
# per-kernel device times of the SVC workload (engine kernel + fp64 re-evaluation of uncertified rows)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"engine_kernel|svc_exact" --csv \
    --log-file gpurun_out/svc_launches.csv python tools/run_workload.py svc 10000000 1 > gpurun_out/svc_launches.stdout 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/svc_launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows:
    print(r[4][:60], r[-1], r[-2])
PY
tail -1 gpurun_out/svc_launches.stdout

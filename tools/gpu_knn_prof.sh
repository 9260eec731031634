#!/bin/bash
# KNN engine: per-kernel device times of one predict (sort kernels, engine, tie re-run), then ONE --set full capture of the
# engine kernel with a per-instruction dump grouped into regions (summarised on the box)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_|engine_kernel" --csv \
    --log-file gpurun_out/knn_launches.csv python tools/run_workload.py knn 10000000 1 > gpurun_out/knn_launches.stdout 2>&1
python - <<'PY'
import csv
for r in csv.reader(open("gpurun_out/knn_launches.csv")):
    if len(r) > 5 and r[0].isdigit(): print(r[4][:70], r[-1], r[-2])
PY
tail -1 gpurun_out/knn_launches.stdout
timeout 600 ncu --set full --clock-control none --import-source on -k regex:engine_kernel -s 1 -c 1 -f -o gpurun_out/prof_knn \
   python tools/run_workload.py knn 10000000 1 > gpurun_out/prof_knn.stdout 2> gpurun_out/prof_knn.stderr
python tools/ncu_sass_dump.py gpurun_out/prof_knn.ncu-rep gpurun_out/knn_sass_counts.txt
python tools/ncu_raw.py gpurun_out/prof_knn.ncu-rep > gpurun_out/knn_raw.txt 2>&1
python tools/sass_regions.py gpurun_out/knn_sass_counts.txt ${KNN_UNITS:-312512} 0 > gpurun_out/knn_sass_regions_all.txt 2>&1
head -2 gpurun_out/knn_sass_regions_all.txt; awk -F'stall%=' 'NF>1 && $2+0 >= 1.5' gpurun_out/knn_sass_regions_all.txt
grep -E "time_duration|tensor|issue_active|alu_cycles|warps_active|inst_executed.sum|cycles_elapsed" gpurun_out/knn_raw.txt
rm -f gpurun_out/prof_knn.ncu-rep; du -sh gpurun_out

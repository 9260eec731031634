#!/bin/bash
# One --set full capture of the KNN engine kernel + a full per-instruction dump (summarised on the box).
mkdir -p gpurun_out
for f in ${KNN_FLUSH_LIST:-31}; do
  echo "flush $f: $(TCSDN_TOOL_OPTS=7=$f timeout 300 python tools/run_workload.py knn 10000000 2 2>&1 | tail -1)"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:engine_kernel -s 2 -c 1 -f -o gpurun_out/prof_knn \
   python tools/run_workload.py knn 10000000 2 > gpurun_out/prof_knn.stdout 2> gpurun_out/prof_knn.stderr
python tools/ncu_sass_dump.py gpurun_out/prof_knn.ncu-rep gpurun_out/knn_sass_counts.txt
python tools/ncu_raw.py gpurun_out/prof_knn.ncu-rep > gpurun_out/knn_raw.txt 2>&1
rm -f gpurun_out/prof_knn.ncu-rep; du -sh gpurun_out

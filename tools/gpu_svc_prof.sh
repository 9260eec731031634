#!/bin/bash
# one full ncu capture of the SVC engine kernel (2M rows), report kept under gpurun_out/
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:engine_kernel -s 1 -c 1 \
    -f -o gpurun_out/prof_svc python tools/run_workload.py svc 2000000 1 > gpurun_out/prof_svc.stdout 2>&1
tail -2 gpurun_out/prof_svc.stdout; ls -la gpurun_out/prof_svc.ncu-rep

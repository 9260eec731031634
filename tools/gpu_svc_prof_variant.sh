#!/bin/bash
# ncu capture of the SVC engine kernel of an experiment build: tools/gpu_svc_prof_variant.sh <variant>
cp traffic_classifier_sdn_b200/libtcsdn.so /tmp/libtcsdn_product.so
cp variants/libtcsdn_$1.so traffic_classifier_sdn_b200/libtcsdn.so
timeout 800 ncu --set full --clock-control none --import-source on -k regex:engine_kernel -s 1 -c 1 \
    -f -o gpurun_out/prof_svc_$1 python tools/run_workload.py svc 2000000 1 > gpurun_out/prof_svc_$1.stdout 2>&1
cp /tmp/libtcsdn_product.so traffic_classifier_sdn_b200/libtcsdn.so
ls -la gpurun_out/prof_svc_$1.ncu-rep

#!/bin/bash
# KNN engine: parity tests, then the 10M x 50k workload's time with pruning on and off and the engine's own counters
# (tiles multiplied per pass, tie rows, exact evaluations)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -q --timeout 150 -k "knn or engine_ragged or nonfinite or full_size_properties" 2>&1 | tail -25
timeout 300 python tools/knn_time.py "$@" 2>&1 | tail -12

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "gnb or golden or ragged or large_batch or device_pointers" > gpurun_out/pytest_gnb.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gnb.log; grep "pre-pass" gpurun_out/pytest_gnb.log
timeout 300 python bench.py --workload gnb --no-extras --steps 20 --warmup 3 > gpurun_out/gnb_fast.json 2>/dev/null; python tools/show_bench.py gpurun_out/gnb_fast.json | head -1
timeout 300 python bench.py --workload gnb_100m --no-extras --steps 5 --warmup 3 > gpurun_out/gnb_100m.json 2>/dev/null; python tools/show_bench.py gpurun_out/gnb_100m.json | head -1

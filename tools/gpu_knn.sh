#!/bin/bash
# KNN engine check: engine + parity tests, then the flush-period sweep on the 10M x 50k workload
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "knn or KNN or kneigh or engine" > gpurun_out/pytest_knn.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_knn.log
for f in ${KNN_FLUSH_LIST:-8 16 31}; do
  echo "flush $f: $(TCSDN_TOOL_OPTS=7=$f timeout 300 python tools/run_workload.py knn 10000000 2 2>&1 | tail -1)"
done

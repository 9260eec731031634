#!/bin/bash
# distance-engine check: engine + parity tests, then KNN / SVC timings at the BASELINE sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_engine.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_engine.log
for f in ${KNN_FLUSH_LIST:-31}; do
  echo "flush $f: $(TCSDN_TOOL_OPTS=7=$f timeout 300 python tools/run_workload.py knn 10000000 2 2>&1 | tail -1)"
done
echo "$(timeout 300 python tools/run_workload.py svc 10000000 2 2>&1 | tail -1)"
timeout 300 python tools/stress_svc.py 2>&1 | tail -2

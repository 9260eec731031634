#!/bin/bash
# forest check: parity tests, then both forest workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "forest or Forest" > gpurun_out/pytest_forest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_forest.log
for w in forest forest_hbm; do
  timeout 300 python bench.py --workload $w --no-extras --steps 10 --warmup 3 > gpurun_out/forest_$w.json 2>/dev/null
  python tools/show_bench.py gpurun_out/forest_$w.json | head -1
done

#!/bin/bash
# Round check on a GPU box: full -m gpu suite, smoke(), then a scorer thread-configuration sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in 0 1 2 3; do for w in gnb logistic kmeans; do   # TCSDN_OPT_SCORER_SHAPE (key 4): 0 auto, 1 = 128x4, 2 = 256x2, 3 = 128x2
  timeout 200 python bench.py --workload $w --no-extras --gpu-only --set-option 4=$c --steps 20 --warmup 3 > gpurun_out/sweep_${c}_$w.json 2>/dev/null
  echo "cfg $c $w: $(python tools/show_bench.py gpurun_out/sweep_${c}_$w.json | head -1)"
done; done

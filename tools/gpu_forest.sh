#!/bin/bash
# forest check: parity tests, then both forest workloads with and without the coherence sort
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "forest or Forest or golden or ragged or large_batch" > gpurun_out/pytest_forest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_forest.log
for s in 1 0; do for w in forest forest_hbm; do     # TCSDN_OPT_FOREST_SORT (key 6)
  timeout 300 python bench.py --workload $w --no-extras --gpu-only --set-option 6=$s --steps 10 --warmup 3 > gpurun_out/forest_${s}_$w.json 2>/dev/null
  echo "sort=$s $w: $(python tools/show_bench.py gpurun_out/forest_${s}_$w.json | head -1)"
done; done
for c in 0 1 3; do for w in forest_hbm forest_hbm2; do   # TCSDN_OPT_FOREST_SHAPE (key 5): 0 auto, 1 = 512x2 (two CTAs per SM), 3 = 1024x1
  timeout 400 python bench.py --workload $w --no-extras --gpu-only --set-option 5=$c --steps 5 --warmup 3 > gpurun_out/forest_shape${c}_$w.json 2>/dev/null
  echo "shape=$c $w: $(python tools/show_bench.py gpurun_out/forest_shape${c}_$w.json | head -1)"
done; done

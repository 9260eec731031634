#!/bin/bash
# round-2 tuning sweeps: scorer CTA shapes on the 1M-row headline, forest CTA shapes on the L2- and HBM-resident forests
mkdir -p gpurun_out
for c in 1 2 3; do
  timeout 200 python bench.py --workload gnb --no-extras --gpu-only --set-option 4=$c --steps 20 --warmup 3 > gpurun_out/sweep_${c}_gnb.json 2>/dev/null
  echo "scorer shape $c gnb: $(python tools/show_bench.py gpurun_out/sweep_${c}_gnb.json | head -1)"
done
for c in 1 3; do for w in forest_hbm forest_hbm2; do
  timeout 600 python bench.py --workload $w --no-extras --gpu-only --set-option 5=$c --steps 5 --warmup 3 > gpurun_out/forest_shape${c}_$w.json 2>/dev/null
  echo "forest shape=$c $w: $(python tools/show_bench.py gpurun_out/forest_shape${c}_$w.json | head -1)"
done; done

#!/bin/bash
# after the KNN engine change: GPU suite, smoke, bench line, then the KNN engine's own evidence (role cycle breakdown from the
# experiment build, per-kernel times of one predict, one --set full capture with the per-instruction regions)
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}.stderr; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_1gpu.json
bash tools/gpu_knn_timing.sh 2>&1 | tee gpurun_out/${TAG}_knn_roles.txt
bash tools/gpu_knn_prof.sh 2>&1 | tee gpurun_out/${TAG}_knn_prof.txt

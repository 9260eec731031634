#!/bin/bash
# Round-end evidence on one B200: full GPU suite, smoke(), the default bench line, the reference arm; profiles: tools/gpu_profile.sh
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1800 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}.stderr; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2>/dev/null; echo "ref rc=$?"; tail -c 300 gpurun_out/bench_${TAG}_reference_arm.json
timeout 600 python tools/stress_svc.py 20 2>&1 | tail -1
bash tools/gpu_profile.sh ${TAG}

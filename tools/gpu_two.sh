#!/bin/bash
# multi-GPU evidence (gpurun --gpus N): comm tests (incl. the two-rank torchrun check), then the N-GPU bench line
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm.py -m gpu -q -x > gpurun_out/pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${N}gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --gpu-only > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${N}gpu.stderr; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_${N}gpu.json; python -c "
import json; j=json.loads(open('gpurun_out/bench_${TAG}_${N}gpu.json').read().strip().splitlines()[-1]); print('with gather:', j.get('with_label_allgather')); print('value_with_gather', j.get('value_with_gather'), 'gather_efficiency', j.get('gather_efficiency'), 'numa', j.get('numa_binding'))"
tail -3 gpurun_out/bench_${N}gpu.stderr

#!/bin/bash
# two-GPU evidence: comm tests (incl. the two-rank torchrun check), then the 2-GPU bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm.py tests/test_parity_gpu.py -m gpu -q -x -k "comm or two_rank or gnb" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r01c_2gpu.json 2> gpurun_out/bench_2gpu.stderr; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_r01c_2gpu.json; python -c "
import json; j=json.loads(open('gpurun_out/bench_r01c_2gpu.json').read().strip().splitlines()[-1]); print(j.get('with_label_allgather'))"
tail -3 gpurun_out/bench_2gpu.stderr

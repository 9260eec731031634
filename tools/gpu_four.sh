#!/bin/bash
# four-GPU evidence: sharded predict + gather check, then the headline bench line at N=4
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29671 tools/two_gpu_gather.py 2>&1 | grep -E "gather|Error" | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 4 --steps 20 --warmup 3 --extras logistic,forest,knn > gpurun_out/bench_r01e_4gpu.json 2> gpurun_out/bench_4gpu.stderr; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_r01e_4gpu.json

#!/bin/bash
# run one pytest selection on the GPU box: tools/gpu_one.sh <pytest args...>
mkdir -p gpurun_out
timeout 1200 python -m pytest "$@" -m gpu -q -x > gpurun_out/pytest_one.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_one.log

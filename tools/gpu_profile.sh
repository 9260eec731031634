#!/bin/bash
# Run on the GPU box (under gpurun): launch list of the default bench + one `--set full` capture per hot kernel.
# Usage: tools/gpu_profile.sh <tag> [kernel-regex workload]...
set -u
TAG=${1:-r01}; shift || true
mkdir -p gpurun_out
NCU=${NCU:-ncu}
# every launch of the default bench with its device time (cold-cache, serialised: compare shares)
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
while [ $# -ge 2 ]; do
    RE=$1; WL=$2; shift 2
    timeout 900 $NCU --set full --clock-control none --import-source on -k regex:$RE -s 3 -c 2 \
        -f -o gpurun_out/prof_${TAG}_${WL} python bench.py --workload $WL --no-extras --steps 5 --warmup 3 \
        > gpurun_out/prof_${TAG}_${WL}.stdout 2> gpurun_out/prof_${TAG}_${WL}.stderr
done

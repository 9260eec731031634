#!/bin/bash
# Run on the GPU box (under gpurun): launch list of the default bench + one `--set full` capture per hot kernel.
# Usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
NCU=${NCU:-ncu}
# every launch of the default bench with its device time (cold-cache, serialised: compare shares, not absolutes)
timeout 1200 $NCU --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --gpu-only > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
prof() {  # name kernel-regex command...
    local NAME=$1 RE=$2; shift 2
    timeout 900 $NCU --set full --clock-control none --import-source on -k regex:$RE -s 3 -c 1 \
        -f -o gpurun_out/prof_${TAG}_${NAME} "$@" > gpurun_out/prof_${TAG}_${NAME}.stdout 2> gpurun_out/prof_${TAG}_${NAME}.stderr
}
prof gnb scorer_tiled python bench.py --workload gnb --no-extras --gpu-only --steps 5 --warmup 3
prof logistic scorer_tiled python bench.py --workload logistic --no-extras --gpu-only --steps 5 --warmup 3
prof forest forest_kernel python bench.py --workload forest --no-extras --gpu-only --steps 3 --warmup 3
prof forest_hbm forest_kernel python bench.py --workload forest_hbm --no-extras --gpu-only --steps 3 --warmup 3
prof forest_hbm2 forest_kernel python bench.py --workload forest_hbm2 --no-extras --gpu-only --steps 3 --warmup 3
prof svc engine_kernel python tools/run_workload.py svc 10000000 4
prof svc_refine svc_exact12 python tools/run_workload.py svc 10000000 4
prof knn engine_kernel python tools/run_workload.py knn 10000000 4
# per-instruction execution counts of the KNN engine kernel, grouped into regions (unit = one warp x one reference tile)
python tools/ncu_sass_dump.py gpurun_out/prof_${TAG}_knn.ncu-rep gpurun_out/knn_sass_counts_${TAG}.txt
mkdir -p gpurun_out/profiles
python tools/sass_regions.py gpurun_out/knn_sass_counts_${TAG}.txt $((19532*782*16)) 2.0 > gpurun_out/profiles/${TAG}_knn_sass_regions.txt 2>&1
# summarise on the box (only gpurun_out/ travels back, 64 MiB at most) and drop the bulky reports that are not needed again
python tools/make_profile_summary.py ${TAG} gpurun_out/profiles > gpurun_out/summary_${TAG}.log 2>&1
rm -f gpurun_out/prof_${TAG}_*.ncu-rep
du -sh gpurun_out

#!/bin/bash
# the HBM-resident forest: bench line for both CTA shapes, one ncu capture (DRAM bytes, sectors per request, L2 hit rate), SVC e2e
mkdir -p gpurun_out
for c in 1 3; do
  timeout 900 python bench.py --workload forest_hbm2 --no-extras --gpu-only --set-option 5=$c --steps 5 --warmup 3 > gpurun_out/forest_shape${c}_forest_hbm2.json 2>gpurun_out/hbm2_$c.err
  echo "forest shape=$c hbm2: $(python tools/show_bench.py gpurun_out/forest_shape${c}_forest_hbm2.json | head -1)"
done
timeout 900 ncu --set full --clock-control none -k regex:forest_kernel -s 3 -c 1 -f -o gpurun_out/prof_hbm2 python bench.py --workload forest_hbm2 --no-extras --gpu-only --steps 3 --warmup 3 > gpurun_out/prof_hbm2.stdout 2>&1
ncu -i gpurun_out/prof_hbm2.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); H=rows[0]; r=rows[2]
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size']:
    if k in H: print(f'{k:70s} {r[H.index(k)]:>18} {rows[1][H.index(k)]}')
" | tee gpurun_out/hbm2_ncu.txt
rm -f gpurun_out/prof_hbm2.ncu-rep
timeout 900 python bench.py --workload svc --no-extras --gpu-only --steps 3 --warmup 3 > gpurun_out/svc_only.json 2>/dev/null; python tools/show_bench.py gpurun_out/svc_only.json | head -3

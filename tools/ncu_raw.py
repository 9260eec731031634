#!/usr/bin/env python
"""Key raw metrics of every kernel in an .ncu-rep.  usage: tools/ncu_raw.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__grid_size','launch__block_size','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum',
 'sm__cycles_elapsed.avg','smsp__cycles_active.avg','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_bytes.sum','lts__t_sector_hit_rate.pct']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); H = rows[0]
for r in rows[2:]:
    print("==", r[H.index('Kernel Name')][:100])
    for w in WANT:
        if w in H: print(f"   {w:75s} {r[H.index(w)]:>18} {rows[1][H.index(w)]}")

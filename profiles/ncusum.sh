#!/bin/bash
# usage: ncusum.sh file.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import sys,csv
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__grid_size','launch__block_size','sm__inst_executed_pipe_fma.sum','smsp__inst_executed.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.sum','smsp__cycles_active.avg','lts__t_sector_hit_rate.pct','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_membar_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__thread_inst_executed_per_inst_executed.ratio','launch__local_mem_per_thread' if False else 'x']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i=hdr.index(w); print('%-80s %s %s'%(w,r[i],units[i]))
"

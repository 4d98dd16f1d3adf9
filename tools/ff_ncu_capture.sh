#!/bin/bash
# hardware-counter capture (no SASS patching: a few replays) of k_ff_tiles on every bench workload -> gpurun_out/ffncu_<workload>.csv
cd ${GRAFT_REPO_ROOT:-.}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,sm__cycles_elapsed.avg.per_second
for w in "$@"; do
  timeout 900 python tools/ff_build_only.py $w > gpurun_out/ffncu_plain_$w.log 2>&1 || { echo "plain run of $w failed"; continue; }
  timeout 1500 ncu --metrics $M --clock-control none --print-units base -k regex:k_ff_tiles -c 1 --csv --log-file gpurun_out/ffncu_$w.csv python tools/ff_build_only.py $w > gpurun_out/ffncu_$w.log 2>&1
  tail -2 gpurun_out/ffncu_$w.csv | cut -c1-300
done

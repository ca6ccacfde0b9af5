#!/bin/bash
# Round-2 profiling pass (run under gpurun from the repo root).  Produces, under gpurun_out/:
#   r2_launches.csv     per-launch device time of one warm-up + one timed bench step of every leg (ncu, serialised, cold cache)
#   r2_full_*.csv       raw-page metrics of one --set full capture per dominant kernel
set -u
OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --skip-cfg5 --csmc-chains 592"
$CMD > $OUT/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/r2_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_launches.csv $CMD > $OUT/r2_ncu_launch.log 2>&1
METRICS="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
i=0
for K in "sample_elements_kernel<float, 32, true>" "filter_warp_kernel<float, 32>" "csmc_multi_kernel" "csmc_wide_kernel" "kalman_sweep_fused_kernel" "sample_replay_kernel<float, 32>" "logpdf_fast_kernel"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$K" -s 1 -c 1 -f -o $OUT/r2_full_$i $CMD > $OUT/r2_ncu_full_$i.log 2>&1
  if [ -f $OUT/r2_full_$i.ncu-rep ]; then
    ncu -i $OUT/r2_full_$i.ncu-rep --page raw --csv --metrics $METRICS > $OUT/r2_full_$i.csv 2>/dev/null
    rm -f $OUT/r2_full_$i.ncu-rep
  fi
done
ls -la $OUT | tail -20

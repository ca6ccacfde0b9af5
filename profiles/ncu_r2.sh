#!/bin/bash
# Round-2 profiling pass (run under gpurun from the repo root).  Produces, under gpurun_out/:
#   r2_plain.json       the plain (un-profiled) run of the same command: its numbers are the ones to quote
#   r2_launches.csv     per-launch device time of one warm-up + one timed bench step of every leg (ncu, serialised, cold cache)
#   r2_full_*.csv       counters of one capture per dominant kernel (a number printed under ncu is never a bench value)
set -u
OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --skip-cfg5 --csmc-chains 592"
$CMD > $OUT/r2_plain.json 2> $OUT/r2_plain.err || { echo "plain run failed"; tail -5 $OUT/r2_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/r2_launches.csv $CMD > $OUT/r2_ncu_launch.log 2>&1
METRICS="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
i=0
# kernel regex : launches of it to skip (the structured / general instantiations of the Kalman kernels alternate behind
# the diagonal-F gate; a late launch is a warm one)
for SPEC in "sample_elements_kernel:4" "filter_warp_kernel:4" "sample_replay_kernel:2" "loglik_rows_kernel:4" "csmc_multi_kernel:1" "csmc_wide_kernel:1" "kalman_sweep_fused_kernel:2" "pit_stitch_kernel:12"; do
  K=${SPEC%%:*}; S=${SPEC##*:}
  i=$((i+1))
  ncu --metrics $METRICS --clock-control none --kernel-name-base demangled -k regex:"$K" -s $S -c 1 --csv --log-file $OUT/r2_full_$i.csv $CMD > $OUT/r2_ncu_full_$i.log 2>&1
done
python - <<PY
import csv, glob, json
out = {}
for f in sorted(glob.glob("$OUT/r2_full_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    if len(rows) < 2 or "Kernel Name" not in rows[0]:
        continue
    h = rows[0]
    name = rows[1][h.index("Kernel Name")]
    m = {r[h.index("Metric Name")]: (r[h.index("Metric Value")], r[h.index("Metric Unit")]) for r in rows[1:]}
    out[name] = m
    print("==", name[:100])
    for k, (v, u) in m.items():
        print("   %-90s %s %s" % (k, v, u))
json.dump(out, open("$OUT/r2_full_summary.json", "w"), indent=1)
PY

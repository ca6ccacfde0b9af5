#!/bin/bash
# Profile of the configs[4] kernels (d = 64, T = 2048): per-call times of one sweep, then the ncu counters of the
# dominant kernels.  Run under gpurun from the repo root; outputs under gpurun_out/.
set -u
OUT=gpurun_out
TAG=${1:-c5}
CMD="python profiles/bench_configs.py --config 5 --chains 592 --steps 2 --warmup 1"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { tail -5 $OUT/${TAG}_plain.err; exit 1; }
cat $OUT/${TAG}_plain.json
METRICS="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers"
i=0
for K in "sample_elements_kernel"; do
  i=$((i+1))
  SKIP=1; [ "$K" = "sample_elements_kernel" ] && SKIP=2  # launches alternate lean, general (gate): capture a lean one
  ncu --metrics $METRICS --clock-control none --kernel-name-base demangled -k regex:"$K" -s $SKIP -c 1 --csv --log-file $OUT/${TAG}_k$i.csv python profiles/bench_configs.py --config 5 --chains 592 --steps 1 --warmup 1 > $OUT/${TAG}_ncu$i.log 2>&1
done
python - <<PY
import csv,glob
for f in sorted(glob.glob("$OUT/${TAG}_k*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if len(rows)<2: print(f,"no data"); continue
    h=rows[0]
    name=rows[1][h.index("Kernel Name")]
    print("==",name[:90])
    for r in rows[1:]:
        print("  %-95s %s"%(r[h.index("Metric Name")], r[h.index("Metric Value")]))
PY

#!/usr/bin/env python
"""Turns the outputs of profiles/ncu_r2.sh (gpurun_out/r2_launches.csv, r2_full_summary.json) into the committed
summaries: profiles/r2_launches_summary.txt, profiles/traffic_r2.json (read by bench.py), profiles/r2_ncu_counters.json."""
import collections
import csv
import json
import re
import shutil

rows = [r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r) > 10]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('auxssm::', '')
    v, u = float(r[vi].replace(',', '')), r[ui]
    tot[name] += v * 1e3 if u.startswith('us') else (v * 1e6 if u.startswith('ms') else v)
    cnt[name] += 1
T = sum(tot.values())
lines = ["ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu --skip-cfg5 --csmc-chains 592` (round 2 final state;",
         "--metrics gpu__time_duration.sum --clock-control none: serialised, cold-cache per-launch times; the SHARES are what to compare",
         "with the CUDA-event shares bench.py prints).  All legs of the command are in the list (headline sweep x4, cSMC configs[2]/[3] at",
         "592 chains, the few-chain fused sweep and its CUDA graph, the parallel-in-time cSMC leg).", "",
         "%-60s %8s %12s %7s" % ("kernel", "launches", "total ms", "share")]
for k, v in tot.most_common(30):
    lines.append("%-60s %8d %12.3f %6.1f%%" % (k[:60], cnt[k], v / 1e6, 100 * v / T))
kal = [k for k in tot if any(s in k for s in ('filter_warp', 'sample_elements', 'sample_replay', 'logpdf', 'loglik_rows', 'sv_', 'auxiliary_draw',
                                              'kalman_mh', 'count_offdiag', 'ell_reduce'))]
Tk = sum(tot[k] for k in kal)
lines += ["", "Kernels of the headline aux-Kalman sweep only (share of their summed time, ms per launch):"]
for k in sorted(kal, key=lambda k: -tot[k])[:9]:
    lines.append("  %-58s %6.1f%%  %8.3f" % (k[:58], 100 * tot[k] / Tk, tot[k] / cnt[k] / 1e6))
open('profiles/r2_launches_summary.txt', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines[-10:]))

full = json.load(open('gpurun_out/r2_full_summary.json'))
num = lambda m, k: float(m[k][0].replace(',', ''))
ub = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
per = {}
for name, m in full.items():
    short = re.sub(r'\(.*', '', name).replace('void ', '')
    e = {"dram_read_bytes": num(m, 'dram__bytes_read.sum') * ub.get(m['dram__bytes_read.sum'][1], 1),
         "dram_write_bytes": num(m, 'dram__bytes_write.sum') * ub.get(m['dram__bytes_write.sum'][1], 1),
         "time_ms": num(m, 'gpu__time_duration.sum') * {'ns': 1e-6, 'us': 1e-3, 'ms': 1}.get(m['gpu__time_duration.sum'][1], 1e-6),
         "issue_active_pct": num(m, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
         "fma_pipe_pct": num(m, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
         "warps_active_pct": num(m, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
         "lsu_wavefronts_pct_of_peak": num(m, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
         "shared_wavefronts": num(m, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
         "warp_instructions": num(m, 'smsp__inst_executed.sum'),
         "tensor_pipe_instructions": num(m, 'sm__inst_executed_pipe_tensor.sum')}
    for tag, cfg in (('csmc_multi', 'configs[2]'), ('csmc_wide', 'configs[3]')):
        if tag in short:
            short += " (%s, 592 chains)" % cfg
            e["chains"] = 592
    per[short] = e
    print("%-52s t=%.3f ms lsu=%.1f%% issue=%.1f%% fma=%.1f%%" % (short[:52], e["time_ms"], e["lsu_wavefronts_pct_of_peak"],
                                                                 e["issue_active_pct"], e["fma_pipe_pct"]))
json.dump({"source": "profiles/ncu_r2.sh (ncu counters, --clock-control none; bench.py --steps 1 --warmup 3 --no-cpu --skip-cfg5 "
                     "--csmc-chains 592), round 2 final state", "per_launch": per}, open('profiles/traffic_r2.json', 'w'), indent=1)
shutil.copy('gpurun_out/r2_launches.csv', 'profiles/r2_launches.csv')
shutil.copy('gpurun_out/r2_full_summary.json', 'profiles/r2_ncu_counters.json')

#!/usr/bin/env python
"""Headline benchmark: time-steps x chains / second per Gibbs sweep (BASELINE.json `metric`).

Workload at N=1 (BASELINE.json configs[1]): multivariate stochastic volatility, d=30, T=1024, 1024
chains, SECOND-ORDER auxiliary Kalman sampler, fp32.  One "step" = one full MH sweep
`kernel(key, state, delta)` over all chains (aux draw, 2 x [proposal construction, Kalman filter,
trajectory log-densities], 1 backward sample, accept/reject).  Chains are independent, so N GPUs run
N x 1024 chains with no data-path collective ("scaling": "weak"); NCCL is used once, after the timed
region, for the cross-chain diagnostics all-gather (split R-hat / ESS of two tracked scalars).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the reference's algorithm on the host CPU cores.  The reference itself is
Python/JAX and JAX is not installable in this image (no network), so that arm runs the NumPy oracle
port of it (oracle/, sequential mode = the mode the authors use on CPU), one worker process per core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "time-steps x chains / second per Gibbs sweep (aux-Kalman)"
UNIT = "time-steps*chains/s"
SV = dict(nu=0.0, phi=0.9, tau=2.0, rho=0.25)
WORKLOAD = dict(d=30, T=1024, chains_per_gpu=1024, order=2, delta=0.005, dtype="f32")  # delta ~ 40 % acceptance


def make_data(d, T, seed):
    from oracle import sv as osv
    rng = np.random.default_rng(seed)
    m0, P0, F, Q, b = osv.get_dynamics(SV["nu"], SV["phi"], SV["tau"], SV["rho"], d)
    xs_true, ys = osv.get_data(rng, SV["nu"], SV["phi"], SV["tau"], SV["rho"], d, T)
    return (m0, P0, F, Q, b), xs_true, ys


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ---------------------------------------------------------------------------------------------

def _cpu_worker(args):
    d, T, chains, order, delta, seed, steps, warmup = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import kalman as ok
    from oracle import sv as osv
    (m0, P0, F, Q, b), xs_true, ys = make_data(d, T, 20250002)
    rng = np.random.default_rng(seed)
    x = xs_true[None] + 0.3 * rng.standard_normal((chains, T, d))
    facs = osv.make_factories(ys, m0, P0, F, Q, b, order)
    times = []
    with np.errstate(all="ignore"):
        for it in range(warmup + steps):
            n_aux = rng.standard_normal((chains, T, d))
            n_eps = rng.standard_normal((chains, T, d))
            un = rng.uniform(size=chains)
            t0 = time.perf_counter()
            x, _, _ = ok.kalman_sweep(x, delta, n_aux, n_eps, un, *facs, False)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def cpu_reference(steps, warmup, chains_per_worker, workers):
    """Returns (value, ms_per_step, sample description).  One process per core, each sweeping its own chains."""
    import concurrent.futures as cf
    w = WORKLOAD
    jobs = [(w["d"], w["T"], chains_per_worker, w["order"], w["delta"], 1000 + i, steps, warmup)
            for i in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with cf.ProcessPoolExecutor(max_workers=workers) as ex:
            res = list(ex.map(_cpu_worker, jobs))
    wall = time.perf_counter() - t0
    per_step = max(sum(r) for r in res) / steps  # slowest worker's mean step time
    value = w["T"] * chains_per_worker * workers / per_step
    sample = ("%d worker processes x %d chains, T=%d, d=%d, second-order sweep, fp64 NumPy oracle port, "
              "sequential-in-time (wall %.1fs)" % (workers, chains_per_worker, w["T"], w["d"], wall))
    return value, per_step * 1e3, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    steps, warmup = max(1, min(args.steps, 2)), 0
    value, ms, sample = cpu_reference(steps, warmup, chains_per_worker=16, workers=cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "SV d=30 T=1024 second-order aux-Kalman sweep (bounded sample of chains)",
                   **{k: WORKLOAD[k] for k in ("d", "T", "order", "delta")}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Python/JAX; JAX is not installable here, so this is the NumPy oracle port of it",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                          ("sw_power_cap", 6)):
            if any(s[col].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from aux_ssm_samplers_b200 import _lib
    from aux_ssm_samplers_b200 import stochastic_volatility as sv
    from aux_ssm_samplers_b200 import diagnostics
    from aux_ssm_samplers_b200.draws import ExternalDraws
    from aux_ssm_samplers_b200.kalman import KalmanSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    w = WORKLOAD
    d, T, C, order, delta = w["d"], w["T"], w["chains_per_gpu"], w["order"], w["delta"]
    if args.chains:
        C = args.chains
    if args.delta > 0:
        delta = args.delta
    dtype = torch.float32
    es = 4
    (m0, P0, F, Q, b), xs_true, ys = make_data(d, T, 20250002)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=device)
    _, kernel = sv.get_kernel(to(ys), to(m0), to(P0), to(F), to(Q), to(b), parallel=True, order=order,
                              dense_R=args.dense_r)
    rng = np.random.default_rng(20250002 + 17 * rank)
    x0 = (xs_true[None] + 0.3 * rng.standard_normal((C, T, d))).astype(np.float32)

    # Pinned host buffers for the end-to-end arm (state + this step's draws in, new state + flags out).
    n_sets = 2
    pin = lambda shape: torch.empty(shape, dtype=dtype).pin_memory()
    h_aux = [pin((C, T, d)) for _ in range(n_sets)]
    h_eps = [pin((C, T, d)) for _ in range(n_sets)]
    h_un = [pin((C,)) for _ in range(n_sets)]
    for i in range(n_sets):
        h_aux[i].copy_(torch.from_numpy(rng.standard_normal((C, T, d), dtype=np.float32)))
        h_eps[i].copy_(torch.from_numpy(rng.standard_normal((C, T, d), dtype=np.float32)))
        h_un[i].copy_(torch.from_numpy(rng.uniform(size=C).astype(np.float32)))
    h_x = pin((C, T, d))
    h_x.copy_(torch.from_numpy(x0))
    h_x_out = pin((C, T, d))
    h_acc = torch.empty((C,), dtype=torch.bool).pin_memory()
    d_aux = [t.to(device) for t in h_aux]
    d_eps = [t.to(device) for t in h_eps]
    d_un = [t.to(device) for t in h_un]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident_step(state, i):
        k = i % n_sets
        return kernel(ExternalDraws([d_aux[k], d_eps[k]], [d_un[k]]), state, delta)

    def e2e_step(i):
        """Serial end-to-end step: H2D, sweep and D2H back to back on one stream."""
        k = i % n_sets
        x = h_x.to(device, non_blocking=True)
        na = h_aux[k].to(device, non_blocking=True)
        ne = h_eps[k].to(device, non_blocking=True)
        un = h_un[k].to(device, non_blocking=True)
        st = kernel(ExternalDraws([na, ne], [un]), KalmanSampler(x, True), delta)
        h_x_out.copy_(st.x, non_blocking=True)
        h_acc.copy_(st.updated, non_blocking=True)
        return st

    # Pipelined end-to-end arm: three streams and two device input sets, so the H2D copy of step i+1 and the
    # D2H read of step i-1 run on the copy engines while the sweep of step i computes.  Every step still moves
    # its own inputs host->device and its own result device->host inside the timed region.
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    p_x = [torch.empty((C, T, d), dtype=dtype, device=device) for _ in range(2)]
    p_aux = [torch.empty((C, T, d), dtype=dtype, device=device) for _ in range(2)]
    p_eps = [torch.empty((C, T, d), dtype=dtype, device=device) for _ in range(2)]
    p_un = [torch.empty((C,), dtype=dtype, device=device) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]      # inputs of set k are on the device
    ev_free = [torch.cuda.Event() for _ in range(2)]    # the sweep that read set k has finished
    ev_done = [torch.cuda.Event() for _ in range(2)]    # the sweep's outputs are complete
    ev_out = [torch.cuda.Event() for _ in range(2)]     # the D2H read of the previous user of the slot finished

    def upload(i):
        k = i % 2
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[k])
            j = i % n_sets
            p_x[k].copy_(h_x, non_blocking=True)
            p_aux[k].copy_(h_aux[j], non_blocking=True)
            p_eps[k].copy_(h_eps[j], non_blocking=True)
            p_un[k].copy_(h_un[j], non_blocking=True)
            ev_in[k].record(s_in)

    def e2e_pipelined(n_steps):
        cur = torch.cuda.current_stream(device)
        keep = [None, None]
        upload(0)
        for i in range(n_steps):
            k = i % 2
            if i + 1 < n_steps:
                upload(i + 1)
            cur.wait_event(ev_in[k])
            st = kernel(ExternalDraws([p_aux[k], p_eps[k]], [p_un[k]]), KalmanSampler(p_x[k], True), delta)
            ev_free[k].record(cur)
            ev_done[k].record(cur)
            keep[k] = st  # outputs stay referenced until their D2H read has been ordered behind the next reuse
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                st.x.record_stream(s_out)
                st.updated.record_stream(s_out)
                h_x_out.copy_(st.x, non_blocking=True)
                h_acc.copy_(st.updated, non_blocking=True)
                ev_out[k].record(s_out)
        cur.wait_event(ev_out[(n_steps - 1) % 2])
        if n_steps > 1:
            cur.wait_event(ev_out[n_steps % 2])
        return keep

    # ---- device-resident timing -------------------------------------------------------------
    state = KalmanSampler(torch.from_numpy(x0).to(device), True)
    for i in range(args.warmup):
        state = resident_step(state, i)
    clocks = ClockSampler(local)
    prof = _lib.Profiler()
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    _lib.set_profiler(prof)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        state = resident_step(state, args.warmup + i)
    ev1.record()
    barrier()
    _lib.set_profiler(None)
    launches = _lib.launch_count() - launches0
    clock_info = clocks.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    units_per_step = T * C * world
    value = units_per_step / (ms_step * 1e-3)
    accept_rate = float(state.updated.float().mean().item())
    finite = bool(torch.isfinite(state.x).all().item())

    # per-ABI-call device time over the timed region (rank-local); the roofline is reported for the call
    # with the largest share of the step (its dominant kernel), the runner-up is listed beside it
    summ = prof.summary()
    units_launch = T * C
    r_elems = d * d if args.dense_r else d
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    step_ms_local = ev0.elapsed_time(ev1)
    cands = {
        "auxssm_kalman_filter_f32": dict(
            kernel=("filter_warp_kernel<float,32>" if not args.dense_r else "filter_recursion_kernel<float,32,H=I>"),
            bytes_unit=(r_elems + d * d + 2 * d) * es,       # R_t, y_t in; P_t, m_t out (F, Q, b, H are stride-0)
            survey_bytes_unit=(5 * d * d + 4 * d) * es,      # SURVEY 8(d) operator-boundary figure for one filter
            flops_unit=2.0 * d ** 3 + 4 * d * d if not args.dense_r else 2 * (2 + 1.0 / 6 + 0.5 + 1) * d ** 3),
        "auxssm_kalman_sample_f32": dict(
            kernel="sample_elements_kernel<float,32,lean> (+ sample_replay_kernel)",
            bytes_unit=(d * d + 3 * d) * es,                 # P_t, m_t, eps_t in; x_t out
            survey_bytes_unit=(3 * d * d + 4 * d) * es,      # read 3d^2+3d, write d
            flops_unit=2 * (1.0 / 6 + 0.5 + 0.5 + 1 + 1.0 / 6) * d ** 3 + 2 * d * d),
    }
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))["per_launch"]
    except Exception:
        pass

    def ncu_traffic(kernel_label):
        for k, v in traffic.items():
            if kernel_label.startswith(k.split("<")[0]):
                return v["dram_read_bytes"] + v["dram_write_bytes"]
        return None

    roofs = []
    for name, cnd in cands.items():
        calls, ms_tot = summ.get(name, (0, 0.0))
        if not calls:
            continue
        per = ms_tot / calls
        # `achieved` uses SURVEY.md section 8(d)'s operator-boundary bytes per (time-step x chain) for this call, as the
        # bench contract asks; the bytes this kernel's own interface moves (F, Q, b, H are stride-0 broadcasts, R_t is
        # diagonal) are reported beside it as `interface_*` -- the fused figure "in addition, never instead".
        ach = cnd["survey_bytes_unit"] * units_launch / (per * 1e-3) / 1e9
        ach_if = cnd["bytes_unit"] * units_launch / (per * 1e-3) / 1e9
        roofs.append({
            "bound": "hbm", "kernel": cnd["kernel"] + " (" + name + ")", "achieved": ach, "peak": peak_hbm,
            "unit": "GB/s", "frac": ach / peak_hbm,
            "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
            "traffic": ncu_traffic(cnd["kernel"]), "traffic_source": "profiles/traffic_r1.json (ncu --set full)",
            "ms_per_launch": per, "launches_timed": calls,
            "alg_bytes_per_unit": cnd["survey_bytes_unit"], "units_per_launch": units_launch,
            "interface_bytes_per_unit": cnd["bytes_unit"], "interface_achieved": ach_if,
            "interface_frac": ach_if / peak_hbm,
            "share_of_step": ms_tot / step_ms_local,
            "fp32_tflops_achieved": cnd["flops_unit"] * units_launch / (per * 1e-3) / 1e12,
            "note": "FP32 issue / dependent-chain latency bound, not HBM bound (DESIGN.md section 4); traffic = ncu DRAM "
                    "bytes of the call's dominant kernel (for the sampler: P_t in + the (G_t, e_t) scratch out, which "
                    "the replay kernel reads back)",
        })
    roofs.sort(key=lambda r: -r["share_of_step"])
    roofline = roofs[0] if roofs else None
    if roofline is not None:
        sweep_bytes = (21 * d * d + 20 * d) * es  # whole sweep, operator-boundary traffic
        roofline["sweep"] = {"alg_bytes_per_unit": sweep_bytes,
                             "achieved": sweep_bytes * units_launch / (step_ms_local / args.steps * 1e-3) / 1e9,
                             "unit": "GB/s"}
        roofline["sweep"]["frac"] = roofline["sweep"]["achieved"] / peak_hbm
    roofline_other = roofs[1:] if len(roofs) > 1 else []
    kernel_ms = {k: {"calls": v[0], "ms_total": v[1]} for k, v in sorted(summ.items())}

    # ---- end-to-end timing (host buffers, H2D/D2H inside the timed region) ---------------------
    for i in range(min(args.warmup, 2)):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e1.record()
    barrier()
    serial_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_pipelined(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kept = e2e_pipelined(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    del kept
    h2d = (3 * C * T * d + C) * es
    d2h = C * T * d * es + C
    e2e = {"value": units_per_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "pipeline": "3 streams, 2 input sets: H2D of step i+1 and D2H of step i-1 overlap the sweep of step i",
           "serial_ms_per_step": serial_ms, "serial_value": units_per_step / (serial_ms * 1e-3)}

    # ---- cross-chain diagnostics all-gather (NCCL; off the timed path) -------------------------
    tracked = torch.stack([state.x[:, 0, 0], state.x[:, -1, 0]], dim=1)  # two tracked scalars per chain
    diag = diagnostics.cross_chain_summary(tracked)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: SV d=30 T=1024, 1024 chains/GPU, second-order aux-Kalman sweep",
                       "d": d, "T": T, "chains_per_gpu": C, "chains_total": C * world, "order": order,
                       "delta": delta, "parallel": True,
                       "R_layout": "dense [C,T,d,d]" if args.dense_r else "diagonal [C,T,d] (Omega_t is diagonal)",
                       "time_chunks": "auto (=1: 1024 chains fill the SMs; see DESIGN.md)",
                       "l2": "every kernel streams more than the 126 MB L2 per launch (P_t history 3.9 GB, states / draws 126 MB each); no flush needed",
                       "draws": "externally supplied normals/uniforms (parity mode), 2 rotating sets"},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_other": roofline_other, "kernel_ms": kernel_ms,
            "accept_rate_last_step": accept_rate, "all_finite": finite, "diagnostics": diag,
        }
        if world == 1 and not args.no_cpu:
            cores = len(os.sched_getaffinity(0))
            v, ms, sample = cpu_reference(1, 0, chains_per_worker=16, workers=cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "ms_per_step_sample": ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU (debug)")
    ap.add_argument("--delta", type=float, default=0.0, help="override the step size delta (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dense-r", action="store_true",
                    help="materialise R_t = Omega_t as dense [C,T,d,d] like the reference (default: diagonal)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

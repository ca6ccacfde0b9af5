#!/usr/bin/env python
"""Benchmark of BASELINE.json's metric: time-steps x chains / second per Gibbs sweep (aux-Kalman, cSMC).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

ONE JSON line.  The headline fields (`value`, `ms_per_step`, `e2e`, `roofline`, `cpu_baseline`) are the
auxiliary-Kalman leg on BASELINE configs[1]: multivariate stochastic volatility, d=30, T=1024, 1024 chains per GPU,
SECOND-ORDER auxiliary Kalman sampler, fp32; one step = one MH sweep `kernel(key, state, delta)` over all chains.
The same line carries, each with its own value / roofline / e2e / cpu_baseline:
  "csmc"     configs[2] (bootstrap cSMC + backward sampling, d=1 SV, N=1024, T=4096, 4096 chains per GPU) and
             configs[3] (guided auxiliary cSMC, killing resampling, d=30, N=256, T=1024, 1024 chains per GPU);
  "config5"  configs[4] (8192 aux-Kalman chains, d=64, T=2048, first-order factory), the batch SPLIT over the ranks
             (strong scaling, sub-batches sized to HBM) followed by split-R-hat + ESS over an NCCL all-gather, timed;
  "few_chain" configs[0]-like regime (d=2, T=128, one chain, fp64): sweeps per second through the CUDA-graph loop.
Chains are independent, so the headline and the cSMC legs scale weakly (N GPUs run N x the chains, no data-path
collective); NCCL carries only the cross-chain diagnostics.

`--impl reference` (and every `cpu_baseline`) times the NumPy PORT of the reference (oracle/, sequential mode, the
mode the authors use on CPU) on the host cores: the reference itself is Python/JAX and JAX is not installable in
this image.  It is a stated baseline, never "JAX".
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

METRIC = "time-steps x chains / second per Gibbs sweep (aux-Kalman; cSMC legs in `csmc`)"
UNIT = "time-steps*chains/s"
SV = dict(nu=0.0, phi=0.9, tau=2.0, rho=0.25)
WORKLOAD = dict(d=30, T=1024, chains_per_gpu=1024, order=2, delta=0.005, dtype="f32")  # delta ~ 40 % acceptance
CSMC_CFG = {
    "configs[2]": dict(d=1, N=1024, T=4096, chains_per_gpu=4096, kind="bootstrap", resampling="multinomial",
                       backward=True, what="bootstrap cSMC + backward sampling, d=1 SV, N=1024, T=4096"),
    "configs[3]": dict(d=30, N=256, T=1024, chains_per_gpu=1024, kind="guided", resampling="killing", backward=True,
                       delta=0.1, what="guided auxiliary cSMC, killing resampling, d=30 SV, N=256, T=1024"),
}
CFG5 = dict(d=64, T=2048, chains_total=8192, order=1, delta=0.05, sub_batch=1024)  # P_t history: 34 GB per sub-batch


def make_data(d, T, seed):
    from oracle import sv as osv
    rng = np.random.default_rng(seed)
    m0, P0, F, Q, b = osv.get_dynamics(SV["nu"], SV["phi"], SV["tau"], SV["rho"], d)
    xs_true, ys = osv.get_data(rng, SV["nu"], SV["phi"], SV["tau"], SV["rho"], d, T)
    return (m0, P0, F, Q, b), xs_true, ys


def load_json(path, default):
    try:
        return json.load(open(os.path.join(ROOT, path)))
    except Exception:
        return default


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baselines: the oracle PORT on host cores (one worker process per core)
# ---------------------------------------------------------------------------------------------

def _cpu_kalman_worker(args):
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


def _cpu_csmc_worker(args):
    name, seed, chains = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import csmc as oc
    cfg = CSMC_CFG[name]
    d, N, T = cfg["d"], cfg["N"], cfg["T"]
    (m0, P0, F, Q, b), xs_true, ys = make_data(d, T, 20250003)
    model = oc.LinearGaussianModel(m0, P0, F, Q, b, "sv", ys=ys)
    rng = np.random.default_rng(seed)
    times = []
    with np.errstate(all="ignore"):
        for _ in range(chains):
            x_star = xs_true + 0.3 * rng.standard_normal((T, d))
            normals = rng.standard_normal((T, N, d))
            u_res, u_kill, u_bwd = rng.uniform(size=(T - 1, N)), rng.uniform(size=(T - 1, N)), rng.uniform(size=T)
            t0 = time.perf_counter()
            if cfg["kind"] == "bootstrap":
                fk = oc.BootstrapFK(model)
            else:
                u, s = oc.aux_u(x_star, cfg["delta"], rng.standard_normal((T, d)))
                fk = oc.GuidedAuxFK(model, u, s)
            oc.csmc_kernel(fk, model, x_star, normals, u_res, u_bwd, cfg["backward"], cfg["resampling"], u_kill)
            times.append(time.perf_counter() - t0)
    return times


def _pool(fn, jobs):
    import concurrent.futures as cf
    if len(jobs) == 1:
        return [fn(jobs[0])]
    with cf.ProcessPoolExecutor(max_workers=len(jobs)) as ex:
        return list(ex.map(fn, jobs))


def cpu_kalman(steps, warmup, chains_per_worker, workers):
    w = WORKLOAD
    t0 = time.perf_counter()
    res = _pool(_cpu_kalman_worker, [(w["d"], w["T"], chains_per_worker, w["order"], w["delta"], 1000 + i, steps, warmup)
                                     for i in range(workers)])
    wall = time.perf_counter() - t0
    per_step = max(sum(r) for r in res) / steps  # slowest worker's mean step time
    value = w["T"] * chains_per_worker * workers / per_step
    sample = ("%d worker processes x %d chains, %d timed sweep(s) after %d warm-up, T=%d, d=%d, second-order sweep, fp64 "
              "NumPy PORT of the reference (oracle/kalman.py), sequential in time (wall %.1fs)"
              % (workers, chains_per_worker, steps, warmup, w["T"], w["d"], wall))
    return value, per_step * 1e3, sample


def cpu_csmc(name, workers, chains_per_worker=1):
    cfg = CSMC_CFG[name]
    t0 = time.perf_counter()
    res = _pool(_cpu_csmc_worker, [(name, 2000 + i, chains_per_worker) for i in range(workers)])
    wall = time.perf_counter() - t0
    per_chain = max(sum(r) for r in res) / chains_per_worker
    value = cfg["T"] * workers / per_chain
    sample = ("%d worker processes x %d chain(s), one sweep each: %s, fp64 NumPy PORT of the reference "
              "(oracle/csmc.py) (wall %.1fs)" % (workers, chains_per_worker, cfg["what"], wall))
    return value, per_chain * 1e3, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    steps, warmup = max(1, min(args.steps, 3)), (1 if args.warmup > 0 else 0)
    value, ms, sample = cpu_kalman(steps, warmup, chains_per_worker=16, workers=cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: SV d=30 T=1024 second-order aux-Kalman sweep (bounded sample of chains)",
                   **{k: WORKLOAD[k] for k in ("d", "T", "order", "delta")}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "NumPy PORT of the reference (the reference is Python/JAX; JAX is not installable in this image)",
    }
    if not args.skip_csmc:
        line["csmc"] = []
        for name in CSMC_CFG:
            v, ms_c, smp = cpu_csmc(name, cores)
            line["csmc"].append({"config": name, "value": v, "unit": UNIT, "ms_per_chain_sweep": ms_c,
                                 "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": smp}})
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


def bind_near_gpu(index):
    """Pin this process to the CPUs of the GPU's own NUMA node BEFORE any pinned host buffer exists, so that the
    end-to-end arm's staging buffers are allocated (first touch) on the node the GPU's PCIe root hangs off.  Eight ranks
    each move 252 MB per step through host memory; with every buffer on one socket they contend for it and for the
    inter-socket link.  Best effort: any failure leaves the affinity as it was."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s" % (dom[-4:].lower(), rest.lower())
        node = int(open(path + "/numa_node").read().strip())
        cpus = set()
        for part in open(path + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        info.update({"pci": bus, "numa_node": node, "local_cpus": len(cpus), "allowed_local_cpus": len(allowed)})
        if node >= 0 and len(allowed) >= 2:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as exc:  # noqa: BLE001
        info["error"] = repr(exc)[:120]
    return info


class Ctx:
    """Per-process distributed / device context and timing helpers."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        self.orig_affinity = os.sched_getaffinity(0)
        self.numa = bind_near_gpu(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, n):
        """barrier, n x fn() between two CUDA events on the current stream, barrier; max over ranks, ms per call."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / n


def kalman_leg(ctx, args, peaks, fp32_peak):
    torch = ctx.torch
    from aux_ssm_samplers_b200 import _lib
    from aux_ssm_samplers_b200 import stochastic_volatility as sv
    from aux_ssm_samplers_b200.draws import ExternalDraws, PhiloxDraws
    from aux_ssm_samplers_b200.kalman import KalmanSampler
    device, world, rank = ctx.device, ctx.world, ctx.rank
    w = WORKLOAD
    d, T, C, order, delta = w["d"], w["T"], w["chains_per_gpu"], w["order"], w["delta"]
    if args.chains:
        C = args.chains
    if args.delta > 0:
        delta = args.delta
    dtype, es = torch.float32, 4
    (m0, P0, F, Q, b), xs_true, ys = make_data(d, T, 20250002)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=device)
    _, kernel = sv.get_kernel(to(ys), to(m0), to(P0), to(F), to(Q), to(b), parallel=True, order=order,
                              dense_R=args.dense_r)
    rng = np.random.default_rng(20250002 + 17 * rank)
    x0 = (xs_true[None] + 0.3 * rng.standard_normal((C, T, d))).astype(np.float32)

    # pinned host buffers of the end-to-end arms
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

    def resident_step(state, i):
        k = i % n_sets
        return kernel(ExternalDraws([d_aux[k], d_eps[k]], [d_un[k]]), state, delta)

    # pipelined end-to-end arm: three streams, two device input sets -- the H2D copy of step i+1 and the D2H read of
    # step i-1 run on the copy engines while the sweep of step i computes.  Every step still moves its own inputs
    # host->device and its own result device->host inside the timed region.
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    p_x = [torch.empty((C, T, d), dtype=dtype, device=device) for _ in range(2)]
    p_aux = [torch.empty((C, T, d), dtype=dtype, device=device) for _ in range(2)]
    p_eps = [torch.empty((C, T, d), dtype=dtype, device=device) for _ in range(2)]
    p_un = [torch.empty((C,), dtype=dtype, device=device) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    philox = PhiloxDraws(20250002 + rank, device)

    def upload(i, with_draws):
        k = i % 2
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[k])
            j = i % n_sets
            p_x[k].copy_(h_x, non_blocking=True)
            if with_draws:
                p_aux[k].copy_(h_aux[j], non_blocking=True)
                p_eps[k].copy_(h_eps[j], non_blocking=True)
                p_un[k].copy_(h_un[j], non_blocking=True)
            ev_in[k].record(s_in)

    def e2e_pipelined(n_steps, with_draws):
        cur = torch.cuda.current_stream(device)
        keep = [None, None]
        upload(0, with_draws)
        for i in range(n_steps):
            k = i % 2
            if i + 1 < n_steps:
                upload(i + 1, with_draws)
            cur.wait_event(ev_in[k])
            key = ExternalDraws([p_aux[k], p_eps[k]], [p_un[k]]) if with_draws else philox
            st = kernel(key, KalmanSampler(p_x[k], True), delta)
            ev_free[k].record(cur)
            ev_done[k].record(cur)
            keep[k] = st
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

    def e2e_serial(i):
        k = i % n_sets
        x = h_x.to(device, non_blocking=True)
        na, ne = h_aux[k].to(device, non_blocking=True), h_eps[k].to(device, non_blocking=True)
        un = h_un[k].to(device, non_blocking=True)
        st = kernel(ExternalDraws([na, ne], [un]), KalmanSampler(x, True), delta)
        h_x_out.copy_(st.x, non_blocking=True)
        h_acc.copy_(st.updated, non_blocking=True)

    # ---- device-resident timing ----
    state = KalmanSampler(torch.from_numpy(x0).to(device), True)
    tracked = []
    for i in range(args.warmup):
        state = resident_step(state, i)
        tracked.append(torch.stack([state.x[:, 0, 0], state.x[:, -1, 0]], dim=1))
    clocks = ClockSampler(ctx.local)
    prof = _lib.Profiler()
    ctx.barrier()
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    _lib.set_profiler(prof)
    _lib.dispatch_log(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        state = resident_step(state, args.warmup + i)
        tracked.append(torch.stack([state.x[:, 0, 0], state.x[:, -1, 0]], dim=1))
    ev1.record()
    ctx.barrier()
    _lib.set_profiler(None)
    launches = _lib.launch_count() - launches0
    kernels_seen = sorted(set(_lib.dispatch_log()))
    clock_info = clocks.stop() if rank == 0 else None
    step_ms_local = ev0.elapsed_time(ev1) / args.steps
    ms_step = ctx.max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    units_per_step = T * C * world
    value = units_per_step / (ms_step * 1e-3)
    accept_rate = float(state.updated.float().mean().item())
    finite = bool(torch.isfinite(state.x).all().item())

    # ---- roofline of the dominant C-ABI call (FP32 issue bound: arithmetic intensity 46 flop/B >> ridge 11) ----
    summ = prof.summary()
    units_launch = T * C
    r_elems = d * d if args.dense_r else d
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_fp32 = float(fp32_peak.get("fp32_tflops", 72.0))
    traffic = load_json("profiles/traffic_r2.json", load_json("profiles/traffic_r1.json", {})).get("per_launch", {})

    def ncu_entry(label):
        for k, v in traffic.items():
            if label.startswith(k.split("<")[0]):
                return v
        return None

    def ncu_traffic(label):
        v = ncu_entry(label)
        return None if v is None else v["dram_read_bytes"] + v["dram_write_bytes"]

    def ncu_smem(label):
        """Shared-memory data-pipe load of the kernel from the committed ncu capture (not measurable with CUDA events):
        the sampler elements kernel is bound by it, not by FP32 issue or HBM."""
        v = ncu_entry(label)
        if v is None or "lsu_wavefronts_pct_of_peak" not in v:
            return None
        return {"bound": "shared-memory data pipe (l1tex__data_pipe_lsu_wavefronts, 1 wavefront / clock / SM)",
                "frac_of_peak": v["lsu_wavefronts_pct_of_peak"] / 100.0,
                "shared_wavefronts_per_unit": v.get("shared_wavefronts", 0) / float(units_launch),
                "source": "profiles/traffic_r2.json (ncu counters of the same kernel at the same shape)"}

    cands = {
        "auxssm_kalman_filter_f32": dict(
            kernel=("filter_warp_kernel<float,32>" if not args.dense_r else "filter_recursion_kernel<float,32,H=I>"),
            bytes_unit=(r_elems + d * d + 2 * d) * es, survey_bytes_unit=(5 * d * d + 4 * d) * es,
            # the structured recursion: predict d^2, d/4 blocks of (4x4 factorisation, 4 x d substitution, rank-4 update)
            flops_unit=2.0 * d ** 3 + 4 * d * d if not args.dense_r else 2 * (2 + 1.0 / 6 + 0.5 + 1) * d ** 3,
            survey_flops_unit=14.0 * d ** 3),
        "auxssm_kalman_sample_f32": dict(
            kernel="sample_elements_kernel<float,32,lean> (+ sample_replay_kernel)",
            bytes_unit=(d * d + 3 * d) * es, survey_bytes_unit=(3 * d * d + 4 * d) * es,
            # chol(S) d^3/3, L^-1 (F P) d^3, P - Z^T Z d^3 (lower half suffices: counted in full), L^-T Z d^3, chol(Sigma) d^3/3
            flops_unit=2 * (1.0 / 6 + 0.5 + 0.5 + 1 + 1.0 / 6) * d ** 3 + 2 * d * d,
            survey_flops_unit=12.7 * d ** 3),
    }
    roofs = []
    for name, cnd in cands.items():
        calls, ms_tot = summ.get(name, (0, 0.0))
        if not calls:
            continue
        per = ms_tot / calls
        tfl = cnd["flops_unit"] * units_launch / (per * 1e-3) / 1e12
        roofs.append({
            "bound": "fp32", "kernel": cnd["kernel"] + " (" + name + ")", "achieved": tfl, "peak": peak_fp32,
            "unit": "TFLOP/s", "frac": tfl / peak_fp32,
            "peak_source": fp32_peak.get("how", "fallback 72 TFLOP/s (profiles/fp32_peak.json missing)"),
            "alg_flops_per_unit": cnd["flops_unit"], "units_per_launch": units_launch,
            "survey_flops_per_unit": cnd["survey_flops_unit"],
            "survey_achieved_tflops": cnd["survey_flops_unit"] * units_launch / (per * 1e-3) / 1e12,
            "traffic": ncu_traffic(cnd["kernel"]), "smem": ncu_smem(cnd["kernel"]),
            "traffic_source": "profiles/traffic_r*.json (ncu --set full, dram__bytes_read + write of the call's dominant kernel)",
            "hbm": {"alg_bytes_per_unit": cnd["survey_bytes_unit"], "interface_bytes_per_unit": cnd["bytes_unit"],
                    "interface_achieved_gbs": cnd["bytes_unit"] * units_launch / (per * 1e-3) / 1e9,
                    "interface_frac_of_hbm_peak": cnd["bytes_unit"] * units_launch / (per * 1e-3) / 1e9 / peak_hbm,
                    "peak_gbs": peak_hbm},
            "ms_per_launch": per, "launches_timed": calls, "share_of_step": ms_tot / (step_ms_local * args.steps),
            "note": "arithmetic intensity of this call is ~%.0f flop/B (ridge ~11): FP32-issue / dependent-chain latency "
                    "bound, so the roofline is the measured FFMA peak; `achieved` counts the flops of the algorithm the "
                    "kernel runs (diagonal F, H = I), `survey_*` SURVEY 8(d)'s dense figure" % (cnd["flops_unit"] / cnd["bytes_unit"]),
        })
    roofs.sort(key=lambda r: -r["share_of_step"])
    roofline = roofs[0] if roofs else None
    kernel_ms = {k: {"calls": v[0], "ms_total": v[1]} for k, v in sorted(summ.items())}

    # ---- end-to-end: host buffers, H2D / D2H inside the timed region ----
    for i in range(min(args.warmup, 2)):
        e2e_serial(i)
    serial_ms = ctx.timed(e2e_serial, args.steps)
    out = {}
    for name, with_draws in (("philox", False), ("external", True)):
        e2e_pipelined(2, with_draws)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kept = e2e_pipelined(args.steps, with_draws)
        e1.record()
        ctx.barrier()
        out[name] = ctx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
        del kept
    e2e = {"value": units_per_step / (out["philox"] * 1e-3), "unit": UNIT, "ms_per_step": out["philox"],
           "h2d_bytes_per_step": C * T * d * es, "d2h_bytes_per_step": C * T * d * es + C,
           "what": "kernel(key, state, delta) with the state in pinned host memory: x H2D, sweep, new x + accept flags D2H "
                   "every step; the sweep's normals / uniforms are generated on the device (PhiloxDraws); 3 streams overlap "
                   "the copies of steps i+1 / i-1 with the sweep of step i",
           "external_draws": {"value": units_per_step / (out["external"] * 1e-3), "ms_per_step": out["external"],
                              "h2d_bytes_per_step": (3 * C * T * d + C) * es,
                              "what": "same, with this step's 2 x [C,T,d] normals + [C] uniforms also shipped from the host "
                                      "(parity mode: externally supplied draws)"},
           "serial_external_draws_ms_per_step": serial_ms}
    hist = torch.stack(tracked, dim=2)  # [C, 2, n]
    res = dict(value=value, ms_step=ms_step, e2e=e2e, roofline=roofline, roofline_other=roofs[1:], kernel_ms=kernel_ms,
               launches=int(launches), kernels_seen=kernels_seen, clock_info=clock_info, accept_rate=accept_rate,
               finite=finite, C=C, d=d, T=T, order=order, delta=delta, hist=hist)
    return res


def csmc_leg(ctx, name, args, peaks):
    torch = ctx.torch
    from aux_ssm_samplers_b200 import _lib, csmc
    from aux_ssm_samplers_b200.draws import PhiloxDraws
    cfg = CSMC_CFG[name]
    device, world, rank = ctx.device, ctx.world, ctx.rank
    d, N, T, C = cfg["d"], cfg["N"], cfg["T"], args.csmc_chains or cfg["chains_per_gpu"]
    dtype, es = torch.float32, 4
    (m0, P0, F, Q, b), xs_true, ys = make_data(d, T, 20250003)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=device)
    M0 = csmc.GaussianDistribution(to(m0), to(P0))
    Mt = csmc.LinearGaussianDynamics(to(F), to(Q), to(b))
    G = csmc.StochasticVolatilityPotential(to(ys))
    rng = np.random.default_rng(20250003 + 31 * rank)
    x0 = (xs_true[None] + 0.3 * rng.standard_normal((C, T, d))).astype(np.float32)
    key = PhiloxDraws(20250003 + rank, device)
    if cfg["kind"] == "bootstrap":
        _, kern = csmc.get_kernel(M0, G, Mt, G, N, cfg["backward"], resampling=cfg["resampling"])
        step = lambda st: kern(key, st)
    else:
        _, kern = csmc.get_generic_kernel(csmc.guided_factory(M0, G, Mt, G), N, cfg["backward"], Pt=Mt,
                                          resampling=cfg["resampling"])
        step = lambda st: kern(key, st, cfg["delta"])
    h_x = torch.empty((C, T, d), dtype=dtype).pin_memory()
    h_x.copy_(torch.from_numpy(x0))
    h_out = torch.empty((C, T, d), dtype=dtype).pin_memory()
    h_upd = torch.empty((C, T), dtype=torch.bool).pin_memory()
    box = {"st": csmc.CSMCState(torch.from_numpy(x0).to(device), None)}

    def resident(i):
        box["st"] = step(box["st"])

    def e2e(i):
        st = step(csmc.CSMCState(h_x.to(device, non_blocking=True), None))
        h_out.copy_(st.x, non_blocking=True)
        h_upd.copy_(st.updated, non_blocking=True)

    steps = max(2, min(args.steps, 3))
    resident(0)
    prof = _lib.Profiler()
    n0 = _lib.launch_count()
    _lib.dispatch_log(reset=True)
    _lib.set_profiler(prof)
    ms = ctx.timed(resident, steps)
    _lib.set_profiler(None)
    launches = _lib.launch_count() - n0
    seen = sorted(set(e for e in _lib.dispatch_log() if "csmc_sweep" in e))
    e2e(0)
    ms_e2e = ctx.timed(e2e, steps)
    units = T * C * world
    calls, ms_tot = prof.summary().get("auxssm_csmc_sweep_f32", (0, 0.0))
    per = ms_tot / max(calls, 1)
    # N[(3d+3)s+4] + (d+1)s + 4 bytes per (time step x chain) with external draws; the draws are
    # generated in the kernel here, which removes N(d+1)s of it (+ Ns of killing uniforms)
    survey_bytes = N * ((3 * d + 3) * es + 4) + (d + 1) * es + 4
    alg_bytes = survey_bytes - N * (d + 1) * es
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    traffic = load_json("profiles/traffic_r2.json", {}).get("per_launch", {})
    tr = None
    for k, v in traffic.items():
        if name in k:
            tr = (v["dram_read_bytes"] + v["dram_write_bytes"]) * (C / float(v.get("chains", C)))
    ach = alg_bytes * T * C / (per * 1e-3) / 1e9 if per else None
    x = box["st"].x
    leg = {
        "config": name, "what": cfg["what"], "chains_per_gpu": C, "chains_total": C * world, "N": N, "T": T, "d": d,
        "value": units / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": 1, "dtype": "f32",
        "draws": "generated in the kernel (Philox4x32-10 keyed on (seed, call, element)): no draw array in HBM",
        "gpu_launches": int(launches), "kernels": seen,
        "roofline": {"bound": "hbm", "kernel": (seen[0] if seen else "auxssm_csmc_sweep_f32"), "achieved": ach,
                     "peak": peak_hbm, "unit": "GB/s", "frac": (ach / peak_hbm) if ach else None, "traffic": tr,
                     "alg_bytes_per_unit": alg_bytes, "survey_bytes_per_unit_external_draws": survey_bytes,
                     "units_per_launch": T * C, "ms_per_launch": per,
                     "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"},
        "e2e": {"value": units / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": C * T * d * es, "d2h_bytes_per_step": C * T * d * es + C * T,
                "what": "kernel(key, state[, delta]) with x* in pinned host memory: H2D, sweep, x + updated flags D2H"},
        "updated_fraction": float(box["st"].updated.float().mean().item()),
        "all_finite": bool(torch.isfinite(x).all().item()),
    }
    del box, x
    torch.cuda.empty_cache()
    return leg


def config5_leg(ctx, args):
    """configs[4]: 8192 independent aux-Kalman chains, d=64, T=2048, sharded over the ranks (STRONG scaling), each rank
    sweeping its share in sub-batches sized to HBM, then split-R-hat + ESS of two tracked scalars over NCCL."""
    torch = ctx.torch
    from aux_ssm_samplers_b200 import diagnostics
    from aux_ssm_samplers_b200 import stochastic_volatility as sv
    from aux_ssm_samplers_b200.draws import PhiloxDraws
    from aux_ssm_samplers_b200.kalman import KalmanSampler
    device, world, rank = ctx.device, ctx.world, ctx.rank
    c5 = CFG5
    d, T, total, order, delta = c5["d"], c5["T"], args.cfg5_chains or c5["chains_total"], c5["order"], c5["delta"]
    mine = total // world + (1 if rank < total % world else 0)
    sub = min(c5["sub_batch"], max(mine, 1))
    dtype = torch.float32
    (m0, P0, F, Q, b), xs_true, ys = make_data(d, T, 20250005)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=device)
    _, kernel = sv.get_kernel(to(ys), to(m0), to(P0), to(F), to(Q), to(b), parallel=True, order=order)
    key = PhiloxDraws(20250005 + rank, device)
    base = to(xs_true)
    # chain states live on the device as one [mine, T, d] array (4.3 GB for 8192 chains); sweeps go sub-batch by sub-batch
    x = base[None] + 0.3 * key.normal((mine, T, d), dtype, device)
    n_sweeps = 3
    tracked = torch.empty((mine, 2, n_sweeps + 1), dtype=dtype, device=device)
    tracked[:, :, 0] = torch.stack([x[:, 0, 0], x[:, -1, 0]], dim=1)

    def sweep(i):
        for lo in range(0, mine, sub):
            st = kernel(key, KalmanSampler(x[lo:lo + sub], True), delta)
            x[lo:lo + sub].copy_(st.x)
        tracked[:, :, i + 1] = torch.stack([x[:, 0, 0], x[:, -1, 0]], dim=1)

    st0 = kernel(key, KalmanSampler(x[:sub], True), delta)  # warm-up on one sub-batch
    del st0
    ms = ctx.timed(sweep, n_sweeps)
    # diagnostics over NCCL: per-chain statistics all-gathered / all-reduced (diagnostics.py)
    diagnostics.split_rhat(tracked[:, 0])
    ctx.barrier()
    t0 = time.perf_counter()
    rhat = [diagnostics.split_rhat(tracked[:, k]) for k in range(2)]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ess = [diagnostics.effective_sample_size(tracked[:, k]) for k in range(2)]
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    leg = {"config": "configs[4]", "what": "aux-Kalman (first order), d=64, T=2048, %d chains sharded over %d GPU(s)" % (total, world),
           "scaling": "strong", "chains_total": total, "chains_this_rank": mine, "sub_batch": sub,
           "value": T * total / (ms * 1e-3), "unit": UNIT, "ms_per_sweep_of_all_chains": ms, "sweeps_timed": n_sweeps,
           "dtype": "f32", "draws": "PhiloxDraws (device)",
           "diagnostics": {"backend": "nccl" if world > 1 else "single process", "tracked_scalars": 2,
                           "samples_per_chain": n_sweeps + 1, "split_rhat": rhat, "ess": ess,
                           "split_rhat_us_per_scalar": (t1 - t0) * 1e6 / 2, "ess_us_per_scalar": (t2 - t1) * 1e6 / 2,
                           "note": "host-timed, includes the all-gather / all-reduce of the per-chain statistics and the "
                                   "device->host read of the result"},
           "all_finite": bool(torch.isfinite(x).all().item())}
    del x, tracked
    torch.cuda.empty_cache()
    return leg


def few_chain_leg(ctx, args):
    """BASELINE configs[0] / SURVEY 8(d) cfg1: linear-Gaussian SSM, d = 2, T = 128, ONE chain, fp64, time-varying dense
    F_t, auxiliary Kalman sampler with the parallel-in-time scans.  Three ways to issue the same sweep:
    composed from the per-primitive entry points (~40 launches), the fused one-launch sweep issued from the host,
    and n_inner fused sweeps + statistics + step-size adaptation replayed as ONE CUDA graph."""
    torch = ctx.torch
    from aux_ssm_samplers_b200 import _lib, latent_gaussian as lgm, mcmc
    from aux_ssm_samplers_b200 import stochastic_volatility as sv
    from aux_ssm_samplers_b200.draws import PhiloxDraws
    from aux_ssm_samplers_b200.kalman import KalmanSampler
    device = ctx.device
    d, T, delta = 2, 128, 0.5
    dtype = torch.float64
    rng = np.random.default_rng(20250001)
    Fs = 0.9 * np.eye(d) + 0.1 * rng.standard_normal((T - 1, d, d)) / np.sqrt(d)
    A = rng.standard_normal((T - 1, d, 5 * d))
    Qs = A @ np.swapaxes(A, -1, -2) / (5 * d)
    bs = rng.standard_normal((T - 1, d))
    m0, P0 = np.zeros(d), np.eye(d)
    x = np.zeros((1, T, d))
    x[0, 0] = rng.standard_normal(d)
    for t in range(1, T):
        x[0, t] = Fs[t - 1] @ x[0, t - 1] + bs[t - 1] + np.linalg.cholesky(Qs[t - 1]) @ rng.standard_normal(d)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=device)
    _, kernel = lgm.get_kernel(None, to(m0), to(P0), to(Fs), to(Qs), to(bs), True, 1, "flat")
    key = PhiloxDraws(1, device)
    box = {"st": KalmanSampler(to(x), True)}
    n0 = _lib.launch_count()
    box["st"] = kernel(key, box["st"], delta)
    per_sweep = _lib.launch_count() - n0

    def eager(i):
        box["st"] = kernel(key, box["st"], delta)

    for i in range(3):
        eager(i)
    ms_eager = ctx.timed(eager, 50)
    n_inner = 32
    loop = mcmc.DeviceLoop(kernel, PhiloxDraws(2, device), KalmanSampler(to(x), True), delta, n_inner=n_inner, adapt=True)
    loop.run(2)
    ms_graph = ctx.timed(lambda i: loop.run(1), 10) / n_inner
    # the same regime composed from the per-primitive entry points (round-1 path): SV d=2 model, parallel-in-time
    (m0s, P0s, Fsv, Qsv, bsv), xs_true, ys = make_data(d, T, 20250001)
    _, kern_c = sv.get_kernel(to(ys), to(m0s), to(P0s), to(Fsv), to(Qsv), to(bsv), parallel=True, order=2)
    boxc = {"st": KalmanSampler(to(xs_true[None]), True)}
    n0 = _lib.launch_count()
    boxc["st"] = kern_c(key, boxc["st"], 0.5)
    per_sweep_c = _lib.launch_count() - n0

    def composed(i):
        boxc["st"] = kern_c(key, boxc["st"], 0.5)

    for i in range(3):
        composed(i)
    ms_comp = ctx.timed(composed, 50)
    return {"config": "configs[0]: LG-SSM d=2, T=128, 1 chain, fp64, time-varying dense F_t, aux-Kalman sweep, parallel scans",
            "fused": {"kernel_launches_per_sweep": int(per_sweep), "ms_per_sweep_host_issued": ms_eager,
                      "ms_per_sweep_cuda_graph": ms_graph, "sweeps_per_graph_replay": n_inner,
                      "kernel_launches_per_replay": int(loop.launches_per_replay),
                      "note": "per sweep: 3 Philox fills (+ counter bumps) and ONE auxssm_kalman_sweep launch (warp-shuffle "
                              "scans); the graph adds running statistics + delta adaptation per sweep"},
            "composed": {"kernel_launches_per_sweep": int(per_sweep_c), "ms_per_sweep_host_issued": ms_comp,
                         "note": "SV d=2 second-order sweep through filter / sampler / log-density entry points (chunked scan)"},
            "value": T / (ms_graph * 1e-3), "unit": UNIT}


def pit_leg(ctx, args):
    """SURVEY 8(f)-3 (secondary, rank 0): parallel-in-time cSMC of csmc/independent.py:78-118 on the d = 1 SV model,
    N = 64 particles, T = 1024, 512 chains: log2(T) + 1 = 11 stitch launches per sweep, N^2 weights per boundary."""
    torch = ctx.torch
    from aux_ssm_samplers_b200 import _lib, csmc
    from aux_ssm_samplers_b200.draws import PhiloxDraws
    d, N, T, C = 1, 64, 1024, 512
    dtype, es = torch.float32, 4
    (m0, P0, F, Q, b), xs_true, ys = make_data(d, T, 20250006)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=ctx.device)
    M0 = csmc.GaussianDistribution(to(m0), to(P0))
    Mt = csmc.LinearGaussianDynamics(to(F), to(Q), to(b))
    G = csmc.StochasticVolatilityPotential(to(ys))
    _, kern = csmc.get_independent_kernel(M0, G, Mt, G, N, gradient=True, parallel=True)
    rng = np.random.default_rng(20250006)
    box = {"st": csmc.CSMCState(to(xs_true[None] + 0.3 * rng.standard_normal((C, T, d))), None)}
    key = PhiloxDraws(20250006, ctx.device)

    def step(i):
        box["st"] = kern(key, box["st"], 0.5)

    step(0)
    n0 = _lib.launch_count()
    ms = ctx.timed(step, 3)
    launches = (_lib.launch_count() - n0) // 3
    # per (time step x chain): the N^2 weights are never in HBM; particles N d s written once and gathered twice,
    # log-weights / origins N (s + 4) read and rewritten once per level that touches the row (log2 T levels)
    levels = int(np.ceil(np.log2(T)))
    alg_bytes = N * d * es * 3 + levels * 2 * N * (es + 4)
    return {"config": "extra: parallel-in-time cSMC (independent.py parallel=True), d=1 SV, N=%d, T=%d, %d chains" % (N, T, C),
            "value": T * C / (ms * 1e-3), "unit": UNIT, "ms_per_sweep": ms, "launches_per_sweep": int(launches),
            "stitch_weights_per_s": (T - 1) * C * N * N / (ms * 1e-3), "alg_bytes_per_unit": alg_bytes,
            "updated_fraction": float(box["st"].updated.float().mean().item()),
            "all_finite": bool(torch.isfinite(box["st"].x).all().item()), "dtype": "f32",
            "draws": "generated in the kernels (Philox)"}


def run_gpu(args):
    ctx = Ctx()
    torch = ctx.torch
    peaks = load_json("MEASURED_PEAKS.json", {})
    fp32_peak = load_json("profiles/fp32_peak.json", {})
    k = kalman_leg(ctx, args, peaks, fp32_peak)
    from aux_ssm_samplers_b200 import diagnostics
    # cross-chain diagnostics of the headline run: split-R-hat / ESS of two tracked scalars over every chain of
    # every rank (NCCL all-gather / all-reduce of per-chain statistics), timed
    hist = k.pop("hist")
    diagnostics.split_rhat(hist[:, 0])
    ctx.barrier()
    t0 = time.perf_counter()
    rhat = [diagnostics.split_rhat(hist[:, j]) for j in range(2)]
    ess = [diagnostics.effective_sample_size(hist[:, j]) for j in range(2)]
    torch.cuda.synchronize()
    diag_us = (time.perf_counter() - t0) * 1e6
    diag = {"chains": int(hist.shape[0]) * ctx.world, "samples_per_chain": int(hist.shape[2]), "split_rhat": rhat,
            "ess": ess, "us_total": diag_us, "backend": "nccl" if ctx.world > 1 else "single process"}
    del hist
    torch.cuda.empty_cache()
    csmc_legs = [] if args.skip_csmc else [csmc_leg(ctx, name, args, peaks) for name in CSMC_CFG]
    cfg5 = None if args.skip_cfg5 else config5_leg(ctx, args)
    few = pit = None
    if ctx.rank == 0 and not args.skip_few:
        few = few_chain_leg(Ctx0(ctx), args)
        pit = pit_leg(Ctx0(ctx), args)
    ctx.barrier()
    if ctx.rank == 0:
        d, T, C = k["d"], k["T"], k["C"]
        line = {
            "metric": METRIC, "value": k["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": k["ms_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: SV d=30 T=1024, 1024 chains/GPU, second-order aux-Kalman sweep",
                       "d": d, "T": T, "chains_per_gpu": C, "chains_total": C * ctx.world, "order": k["order"],
                       "delta": k["delta"],
                       "time_parallelism": "time_chunks auto = 1: 1024 independent chains already fill the SMs, so the "
                                           "sequential-in-time recursion runs (parallel=True is requested; see DESIGN.md)",
                       "R_layout": "dense [C,T,d,d]" if args.dense_r else "diagonal [C,T,d] (Omega_t is diagonal)",
                       "l2": "every kernel streams more than the 126 MB L2 per launch (P_t history 3.9 GB, states / draws "
                             "126 MB each); no flush needed",
                       "draws": "`value`: externally supplied normals / uniforms resident in HBM (parity mode), 2 rotating "
                                "sets; `e2e`: generated on the device (Philox)"},
            "clocks": k["clock_info"], "host_numa": ctx.numa, "e2e": k["e2e"], "gpu_launches": k["launches"], "kernels": k["kernels_seen"],
            "roofline": k["roofline"], "roofline_other": k["roofline_other"], "kernel_ms": k["kernel_ms"],
            "accept_rate_last_step": k["accept_rate"], "all_finite": k["finite"], "diagnostics": diag,
            "csmc": csmc_legs, "config5": cfg5, "few_chain": few, "pit_csmc": pit,
        }
        # a leg whose states went non-finite is not a measurement: say so loudly (stderr + a key in the line)
        legs = [("headline", line)] + [(l["config"], l) for l in csmc_legs] + [("configs[4]", cfg5), ("pit_csmc", pit)]
        bad = [name for name, l in legs if l is not None and l.get("all_finite") is False]
        line["non_finite_legs"] = bad
        if bad:
            print("bench.py: NON-FINITE STATES in %s -- those numbers are invalid" % bad, file=sys.stderr, flush=True)
        if ctx.world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, ctx.orig_affinity)  # the CPU baseline uses every core the box gives us
            cores = len(os.sched_getaffinity(0))
            v, ms, sample = cpu_kalman(1, 0, chains_per_worker=16, workers=cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "ms_per_step_sample": ms}
            for leg in csmc_legs:
                v, ms_c, smp = cpu_csmc(leg["config"], cores)
                leg["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": smp}
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


class Ctx0:
    """Single-rank view of a context (a leg that only rank 0 runs must not call collectives)."""

    def __init__(self, ctx):
        self.torch, self.device, self.world, self.rank, self.local = ctx.torch, ctx.device, 1, 0, ctx.local

    def barrier(self):
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        return v

    timed = Ctx.timed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU of the headline leg (debug)")
    ap.add_argument("--csmc-chains", type=int, default=0, help="override chains per GPU of the cSMC legs (debug)")
    ap.add_argument("--cfg5-chains", type=int, default=0, help="override the total number of configs[4] chains (debug)")
    ap.add_argument("--delta", type=float, default=0.0, help="override the step size delta (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--skip-csmc", action="store_true")
    ap.add_argument("--skip-cfg5", action="store_true")
    ap.add_argument("--skip-few", action="store_true")
    ap.add_argument("--dense-r", action="store_true",
                    help="materialise R_t = Omega_t as dense [C,T,d,d] like the reference (default: diagonal)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

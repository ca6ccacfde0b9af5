#!/usr/bin/env python
"""Secondary benchmark: conditional-SMC particle Gibbs sweeps (BASELINE.json configs[2] and configs[3]).

  python profiles/bench_csmc.py --config 3 [--chains C] [--steps K]
    config 3: bootstrap cSMC + backward sampling, d=1 SV, N=1024, T=4096   (configs[2])
    config 4: guided auxiliary cSMC, killing resampling, d=30, N=256, T=1024 (configs[3])
Draws are generated once on the device (torch.Generator) and replayed (parity-mode inputs: the kernel
reads normals/uniforms from HBM exactly as with externally supplied draws).  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aux_ssm_samplers_b200 import _lib, csmc  # noqa: E402
from aux_ssm_samplers_b200.draws import ExternalDraws  # noqa: E402
from oracle import sv as osv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dtype, es = torch.float32, 4
    if args.config == 3:
        d, N, T, C = 1, 1024, 4096, args.chains or 592
        backward, resampling, kind = True, "multinomial", "bootstrap"
    else:
        d, N, T, C = 30, 256, 1024, args.chains or 296
        backward, resampling, kind = True, "killing", "guided"
    rng = np.random.default_rng(20250000 + args.config)
    m0, P0, F, Q, b = osv.get_dynamics(0.0, 0.9, 2.0, 0.25, d)
    xs_true, ys = osv.get_data(rng, 0.0, 0.9, 2.0, 0.25, d, T)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=dev)
    M0 = csmc.GaussianDistribution(to(m0), to(P0))
    Mt = csmc.LinearGaussianDynamics(to(F), to(Q), to(b))
    G = csmc.StochasticVolatilityPotential(to(ys))
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    normals = torch.randn((C, T, N, d), dtype=dtype, device=dev, generator=gen)
    u_res = torch.rand((C, T - 1, N), dtype=dtype, device=dev, generator=gen)
    u_kill = torch.rand((C, T - 1, N), dtype=dtype, device=dev, generator=gen) if resampling == "killing" else None
    u_bwd = torch.rand((C, T), dtype=dtype, device=dev, generator=gen)
    n_aux = torch.randn((C, T, d), dtype=dtype, device=dev, generator=gen)
    x0 = to(xs_true[None] + 0.3 * rng.standard_normal((C, T, d)))
    if kind == "bootstrap":
        _, kernel = csmc.get_kernel(M0, G, Mt, G, N, backward, resampling=resampling)
        step = lambda st: kernel(ExternalDraws([normals], [u_res, u_bwd]), st)
    else:
        _, kernel = csmc.get_generic_kernel(csmc.guided_factory(M0, G, Mt, G), N, backward, Pt=Mt,
                                            resampling=resampling)
        unis = [u_res, u_kill, u_bwd]
        step = lambda st: kernel(ExternalDraws([n_aux, normals], unis), st, 0.1)
    state = csmc.CSMCState(x0, None)
    for _ in range(args.warmup):
        state = step(state)
    torch.cuda.synchronize()
    prof = _lib.Profiler()
    _lib.set_profiler(prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        state = step(state)
    e1.record()
    torch.cuda.synchronize()
    _lib.set_profiler(None)
    ms = e0.elapsed_time(e1) / args.steps
    summ = prof.summary()
    k_calls, k_ms = summ.get("auxssm_csmc_sweep_f32", (1, 0.0))
    per_launch = k_ms / max(k_calls, 1)
    n_u = 2 if resampling == "killing" else 1
    # algorithmic bytes per (time step x chain): draws in, history out, history back in, outputs
    bytes_unit = N * (d + n_u) * es + (N * (d + 1) * es) * 2 + (d + 1) * es + 4
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    ach = bytes_unit * T * C / (per_launch * 1e-3) / 1e9
    print(json.dumps({
        "metric": "time-steps x chains / second per Gibbs sweep (cSMC)", "value": T * C / (ms * 1e-3),
        "unit": "time-steps*chains/s", "ms_per_step": ms, "dtype": "f32",
        "config": {"workload": "configs[%d]" % (args.config - 1), "kind": kind, "d": d, "N": N, "T": T, "chains": C,
                   "backward_sampling": backward, "resampling": resampling},
        "roofline": {"bound": "hbm", "kernel": "csmc_sweep_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "ms_per_launch": per_launch, "alg_bytes_per_unit": bytes_unit},
        "updated_frac": float(state.updated.float().mean().item()),
        "finite": bool(torch.isfinite(state.x).all().item()),
    }))


if __name__ == "__main__":
    main()

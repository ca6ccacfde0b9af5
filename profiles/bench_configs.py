#!/usr/bin/env python
"""Secondary measurements of the aux-Kalman sweep on the other BASELINE.json shapes (not bench lines):
  configs[0]: linear-Gaussian SSM d=2, T=128, single chain, parallel-in-time filter (the reference's CPU-runnable case)
  configs[4]: d=64, T=2048 chains (per-GPU share of the 8192-chain batch: 1024; --chains to override)
python profiles/bench_configs.py --config 1|5 [--chains C] [--steps K]   -> one JSON line"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aux_ssm_samplers_b200 import _lib, kalman  # noqa: E402
from aux_ssm_samplers_b200 import stochastic_volatility as sv  # noqa: E402
from aux_ssm_samplers_b200.draws import TorchDraws  # noqa: E402
from aux_ssm_samplers_b200.kalman import KalmanSampler  # noqa: E402
from oracle import sv as osv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(7)
    if args.config == 1:
        d, T, C, dtype = 2, 128, args.chains or 1, torch.float64
    else:
        d, T, C, dtype = 64, 2048, args.chains or 1024, torch.float32
    m0, P0, F, Q, b = osv.get_dynamics(0.0, 0.9, 2.0, 0.25, d)
    xs_true, ys = osv.get_data(rng, 0.0, 0.9, 2.0, 0.25, d, T)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=dev)
    _, kernel = sv.get_kernel(to(ys), to(m0), to(P0), to(F), to(Q), to(b), parallel=True, order=2)
    state = KalmanSampler(to(xs_true[None] + 0.3 * rng.standard_normal((C, T, d))), True)
    key = TorchDraws(11)
    delta = 0.005 if d > 2 else 0.2
    for _ in range(args.warmup):
        state = kernel(key, state, delta)
    torch.cuda.synchronize()
    prof = _lib.Profiler()
    _lib.set_profiler(prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        state = kernel(key, state, delta)
    e1.record()
    torch.cuda.synchronize()
    _lib.set_profiler(None)
    ms = e0.elapsed_time(e1) / args.steps
    calls = {k: round(v[1] / v[0], 4) for k, v in prof.summary().items()}
    print(json.dumps({"config": "configs[%d]" % (args.config - 1), "d": d, "T": T, "chains": C,
                      "dtype": str(dtype).split(".")[-1], "ms_per_sweep": ms,
                      "value": C * T / (ms * 1e-3), "unit": "time-steps*chains/s",
                      "accept_rate": float(state.updated.float().mean()), "finite": bool(torch.isfinite(state.x).all()),
                      "ms_per_call": calls, "draws": "device Philox via torch.Generator (TorchDraws), inside the timed region"}))


if __name__ == "__main__":
    main()

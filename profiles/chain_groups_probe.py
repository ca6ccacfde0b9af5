"""Probe: the headline sweep (configs[1]) with the chains split into G independent groups, one CUDA stream each.

The filter is latency-bound (one warp per chain, ~7 warps per SM at 1024 chains) and the sampler is bound by the
shared-memory pipe; independent chain groups on separate streams let one group's filter run under another group's
sampler.  Prints ms per sweep of all chains for G = 1, 2, 4, 8 (K sweeps per group issued back to back, joined once).
Run under gpurun: python profiles/chain_groups_probe.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from aux_ssm_samplers_b200 import stochastic_volatility as sv  # noqa: E402
from aux_ssm_samplers_b200.draws import PhiloxDraws  # noqa: E402
from aux_ssm_samplers_b200.kalman import KalmanSampler  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    w = bench.WORKLOAD
    d, T, C, order, delta = w["d"], w["T"], w["chains_per_gpu"], w["order"], w["delta"]
    (m0, P0, F, Q, b), xs_true, ys = bench.make_data(d, T, 20250002)
    to = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
    _, kernel = sv.get_kernel(to(ys), to(m0), to(P0), to(F), to(Q), to(b), parallel=True, order=order)
    rng = np.random.default_rng(1)
    x0 = torch.from_numpy((xs_true[None] + 0.3 * rng.standard_normal((C, T, d))).astype(np.float32)).to(dev)
    K, W = 10, 3
    for G, prio in [(1, False), (2, False), (4, False), (8, False), (2, True), (4, True), (8, True), (16, True)]:
        lo_p, hi_p = -5, 0
        streams = [torch.cuda.Stream(dev, priority=(-(g % 6) if prio else 0)) for g in range(G)]
        keys = [PhiloxDraws(100 + g, dev) for g in range(G)]
        sl = [slice(g * C // G, (g + 1) * C // G) for g in range(G)]
        states = [KalmanSampler(x0[s].clone(), True) for s in sl]
        cur = torch.cuda.current_stream(dev)

        def run(n):
            ev = torch.cuda.Event()
            ev.record(cur)
            for s in streams:
                s.wait_event(ev)
            for _ in range(n):
                for g in range(G):
                    with torch.cuda.stream(streams[g]):
                        states[g] = kernel(keys[g], states[g], delta)
            for s in streams:
                e = torch.cuda.Event()
                e.record(s)
                cur.wait_event(e)

        run(W)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        run(K)
        e1.record(cur)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        acc = float(torch.cat([s.updated.float() for s in states]).mean())
        print(json.dumps({"groups": G, "stream_priorities": prio, "ms_per_sweep": ms, "value": T * C / ms * 1e3,
                          "accept": acc}), flush=True)


if __name__ == "__main__":
    main()

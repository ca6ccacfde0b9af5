#!/usr/bin/env python
"""Timeline of one device-resident sweep of the bench workload: duration of every C-ABI call and the gap (torch glue,
allocator, host latency) before it.  python profiles/gap_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aux_ssm_samplers_b200 import _lib, stochastic_volatility as sv  # noqa: E402
from aux_ssm_samplers_b200.draws import ExternalDraws  # noqa: E402
from aux_ssm_samplers_b200.kalman import KalmanSampler  # noqa: E402

dev = torch.device("cuda", 0)
w = bench.WORKLOAD
d, T, C, delta = w["d"], w["T"], w["chains_per_gpu"], w["delta"]
(m0, P0, F, Q, b), xs_true, ys = bench.make_data(d, T, 1)
to = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
_, kernel = sv.get_kernel(to(ys), to(m0), to(P0), to(F), to(Q), to(b), parallel=True, order=w["order"])
g = torch.Generator(device=dev); g.manual_seed(0)
na = torch.randn((C, T, d), device=dev, generator=g); ne = torch.randn((C, T, d), device=dev, generator=g)
un = torch.rand((C,), device=dev, generator=g)
state = KalmanSampler(to(xs_true[None]).repeat(C, 1, 1).contiguous(), True)
for _ in range(3):
    state = kernel(ExternalDraws([na, ne], [un]), state, delta)
torch.cuda.synchronize()
prof = _lib.Profiler(); _lib.set_profiler(prof)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
state = kernel(ExternalDraws([na, ne], [un]), state, delta)
e1.record(); torch.cuda.synchronize(); _lib.set_profiler(None)
prev = e0
print("sweep %.3f ms" % e0.elapsed_time(e1))
for name, a, bb in prof.records:
    print("  gap %7.3f ms | %-28s %7.3f ms" % (prev.elapsed_time(a), name, a.elapsed_time(bb)))
    prev = bb
print("  gap %7.3f ms | end" % prev.elapsed_time(e1))

"""Probe: host time of one kernel() call of the headline sweep (no synchronisation), and a cProfile of 20 calls."""
import cProfile, pstats, io, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aux_ssm_samplers_b200 import stochastic_volatility as sv
from aux_ssm_samplers_b200.draws import PhiloxDraws
from aux_ssm_samplers_b200.kalman import KalmanSampler

dev = torch.device("cuda", 0)
w = bench.WORKLOAD
d, T, order, delta = w["d"], w["T"], w["order"], w["delta"]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
(m0, P0, F, Q, b), xs_true, ys = bench.make_data(d, T, 20250002)
to = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
_, kernel = sv.get_kernel(to(ys), to(m0), to(P0), to(F), to(Q), to(b), parallel=True, order=order)
x0 = torch.from_numpy((xs_true[None] + 0.3 * np.random.default_rng(1).standard_normal((C, T, d))).astype(np.float32)).to(dev)
key = PhiloxDraws(1, dev)
st = KalmanSampler(x0, True)
for _ in range(5):
    st = kernel(key, st, delta)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    st = kernel(key, st, delta)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(json.dumps({"chains": C, "host_ms_per_call": (t1 - t0) / 20 * 1e3, "drain_ms": (t2 - t1) * 1e3}))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    st = kernel(key, st, delta)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue())

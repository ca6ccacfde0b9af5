import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tests.test_round2_gpu import _latent_problem, dev
from aux_ssm_samplers_b200 import latent_gaussian as lgm
from aux_ssm_samplers_b200.draws import PhiloxDraws
from aux_ssm_samplers_b200.kalman import KalmanSampler
rng = np.random.default_rng(0)
d, T = 2, 128
Fs, Qs, bs, m0, P0, ys, x = _latent_problem(rng, d, T, 1, "sv")
t = lambda a: dev(a, torch.float64)
for delta in (0.2, 0.05, 0.01, 0.002):
    key = PhiloxDraws(4)
    xs = t(x)
    accs=[]
    for it in range(20):
        na, ne, un = key.normal(xs.shape, xs.dtype, "cuda"), key.normal(xs.shape, xs.dtype, "cuda"), key.uniform((1,), xs.dtype, "cuda")
        xn, acc, alpha = lgm.fused_sweep(xs, t(m0), t(P0), t(Fs), t(Qs), t(bs), t(ys), na, ne, un, delta, 2, "sv")
        accs.append((float(alpha), bool(acc)))
        xs = xn
    print(delta, accs[:6], sum(a for _, a in accs))

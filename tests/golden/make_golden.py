#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

These are REGRESSION vectors produced by the NumPy oracle (oracle/), not outputs of the reference itself: the reference
needs JAX, which cannot be installed in this environment, so its own test pins are re-derived in tests/test_oracle_*.py
instead.  The fixtures freeze the oracle's outputs for fixed seeds so that (a) the oracle cannot
drift silently and (b) the CUDA path is also checked against stored numbers, not only against a live oracle run.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import csmc as oc  # noqa: E402
from oracle import kalman as ok  # noqa: E402
from oracle import sv as osv  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def kalman_sweep_case(seed=2025, T=24, d=5, C=3, delta=0.2, order=2):
    rng = np.random.default_rng(seed)
    m0, P0, F, Q, b = osv.get_dynamics(0.0, 0.9, 2.0, 0.25, d)
    xs_true, ys = osv.get_data(rng, 0.0, 0.9, 2.0, 0.25, d, T)
    x = xs_true[None] + 0.3 * rng.standard_normal((C, T, d))
    n_aux, n_eps, un = rng.standard_normal((C, T, d)), rng.standard_normal((C, T, d)), rng.uniform(size=C)
    ex, eacc, aux = ok.kalman_sweep(x, delta, n_aux, n_eps, un, *osv.make_factories(ys, m0, P0, F, Q, b, order), True)
    return dict(ys=ys, m0=m0, P0=P0, F=F, Q=Q, b=b, x=x, n_aux=n_aux, n_eps=n_eps, un=un, delta=delta, order=order,
                x_out=ex, accepted=eacc, alpha=aux["alpha"])


def csmc_case(seed=77, T=10, N=48, d=2, C=2):
    rng = np.random.default_rng(seed)
    model = oc.LinearGaussianModel(np.zeros(d), np.eye(d), 0.9 * np.eye(d), 0.19 * np.eye(d) + 0.02, np.zeros(d), "sv",
                                   ys=rng.standard_normal((T, d)))
    x_star = rng.standard_normal((C, T, d))
    normals = rng.standard_normal((C, T, N, d))
    u_res, u_bwd = rng.uniform(size=(C, T - 1, N)), rng.uniform(size=(C, T))
    xs, Bs, As = [], [], []
    for c in range(C):
        x, upd, B, (w_T, hist_x, lws, A) = oc.csmc_kernel(oc.BootstrapFK(model), model, x_star[c], normals[c], u_res[c],
                                                          u_bwd[c], True)
        xs.append(x); Bs.append(B); As.append(A)
    return dict(m0=model.m0, P0=model.P0, F=model.F, Q=model.Q, b=model.b, ys=model.ys, x_star=x_star, normals=normals,
                u_res=u_res, u_bwd=u_bwd, x_out=np.stack(xs), ancestors=np.stack(Bs), As=np.stack(As))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "kalman_sweep_sv_d5.npz"), **kalman_sweep_case())
    np.savez_compressed(os.path.join(HERE, "csmc_bootstrap_sv_d2.npz"), **csmc_case())
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))

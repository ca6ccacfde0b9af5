"""Oracle factories for the latent-Gaussian auxiliary Kalman sampler (oracle only; test infrastructure).

Generalises ``oracle/sv.py:make_factories`` (examples/stochastic_volatility/auxiliary_kalman.py:10-57) to time-varying
dynamics and the closed set of separable potentials of the fused sweep kernel: flat, Gaussian observation
(sum_i log N(y_ti; x_ti, sig)), stochastic volatility (sum_i log N(y_ti; 0, exp(x_ti / 2)))."""
import numpy as np

from .linalg import chol, mv, mvn_logpdf, norm_logpdf


def potential_terms(kind, x, ys, sig):
    if kind == "flat":
        return np.zeros_like(x)
    with np.errstate(all="ignore"):
        if kind == "gaussian":
            return np.nan_to_num(norm_logpdf(ys, x, sig))
        return np.nan_to_num(norm_logpdf(ys, 0.0, np.exp(0.5 * x)))


def grad_hess(kind, x, ys, sig):
    if kind == "flat":
        return np.zeros_like(x), np.zeros_like(x)
    if kind == "gaussian":
        return (ys - x) / sig ** 2, np.full_like(x, -1.0 / sig ** 2)
    w = 0.5 * ys * ys * np.exp(-x)
    return w - 0.5, -w


def make_factories(ys, m0, P0, Fs, Qs, bs, order, potential="sv", sig=1.0):
    """-> (dynamics_factory, observations_factory, log_likelihood_fn) for ``oracle.kalman.kalman_sweep``;
    ``Fs, Qs [T-1, d, d]``, ``bs [T-1, d]``."""
    Tm1, d = bs.shape
    T = Tm1 + 1
    eye = np.eye(d)
    cP0, cQ = chol(P0), chol(Qs)

    def dynamics_factory(_x):
        return m0, P0, Fs, Qs, bs

    def observations_factory(x, u, delta):
        grad, hess = grad_hess(potential, x, ys, sig)
        lead = x.shape[:-2]
        Hs = np.broadcast_to(eye, lead + (T, d, d))
        cs = np.zeros(lead + (T, d))
        if order == 1:
            aux = u + 0.5 * delta * np.nan_to_num(grad)
            return aux, Hs, np.broadcast_to(0.5 * delta * eye, lead + (T, d, d)), cs
        omega = 1.0 / (-hess + 2.0 / delta)
        return omega * (2.0 * u / delta + grad - hess * x), Hs, omega[..., :, None] * eye, cs

    def log_likelihood_fn(x):
        out = mvn_logpdf(x[..., 0, :], m0, cP0)
        pred = mv(Fs, x[..., :-1, :]) + bs
        out = out + np.sum(mvn_logpdf(x[..., 1:, :], pred, cQ), axis=-1)
        return out + np.sum(potential_terms(potential, x, ys, sig), axis=(-2, -1))

    return dynamics_factory, observations_factory, log_likelihood_fn

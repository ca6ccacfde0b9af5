"""Multivariate stochastic-volatility model adapters.

Mirrors the reference's ``aux_samplers/examples/stochastic_volatility``: ``get_dynamics`` /
``stationary_covariance`` (model.py:35-53) and the auxiliary-Kalman kernel getter
``get_kernel(ys, m0, P0, F, Q, b, parallel, order)`` (auxiliary_kalman.py:10-57), with the per-sweep
proposal construction (gradient / Hessian of the observation potential, first / second order
observations factory, potential part of ``log_likelihood_fn``) running as CUDA kernels of libauxssm.so.
"""
import torch

from . import _lib
from . import kalman


def stationary_covariance(phi, tau, rho, dim, dtype=torch.float64, device="cpu"):
    """model.py:46-53."""
    U = tau * rho * torch.ones((dim, dim), dtype=dtype, device=device)
    U.fill_diagonal_(tau)
    return U / (1.0 - phi ** 2)


def get_dynamics(nu, phi, tau, rho, dim, dtype=torch.float64, device="cpu"):
    """model.py:35-43 (including its ``b = mu + F @ mu``)."""
    F = phi * torch.eye(dim, dtype=dtype, device=device)
    Q = stationary_covariance(phi, tau, rho, dim, dtype, device)
    mu = nu * torch.ones((dim,), dtype=dtype, device=device)
    b = mu + F @ mu
    return mu, Q.clone(), F, Q, b


def observations_factory(ys, x, u, delta, order, dense_R=False):
    """auxiliary_kalman.py:28-48.  ``x, u``: [C, T, d]; returns (aux_ys, Hs=None (identity), Rs, cs=None).

    Both factories produce DIAGONAL observation covariances ((delta/2) I, resp. Omega_t with a diagonal
    Hessian, model.py:66-69).  By default they are returned as ``kalman.Diagonal([C, T or 1, d])``;
    ``dense_R=True`` materialises ``[C, T or 1, d, d]`` exactly like the reference does."""
    dev = _lib.require_cuda(ys, x, u)
    sfx = _lib.suffix(x.dtype)
    C, T, d = x.shape
    x, u, ys = x.contiguous(), u.contiguous(), ys.contiguous()
    dl = kalman._as_delta(delta, C, x.dtype, dev)
    aux_ys = torch.empty_like(x)
    Tr = 1 if order == 1 else T
    Rs = torch.empty((C, Tr, d, d) if dense_R else (C, Tr, d), dtype=x.dtype, device=dev)
    ys_stride = 0 if ys.dim() == 2 else T * d
    _lib.call("auxssm_sv_obs_factory_" + sfx,
        _lib.stream_ptr(dev), order, C, T, d, _lib.ptr(ys), ys_stride, _lib.ptr(x), _lib.ptr(u), _lib.ptr(dl),
        1 if dl.numel() > 1 else 0, _lib.ptr(aux_ys), _lib.ptr(Rs), 1 if dense_R else 0)
    return aux_ys, None, (Rs if dense_R else kalman.Diagonal(Rs)), None


def log_potential(ys, x, out=None):
    """sum_t sum_i nan_to_num(log N(y_ti; 0, exp(x_ti / 2)))  (model.py:56-76); x: [C, T, d] -> [C].
    With ``out`` given, accumulates onto it in place."""
    dev = _lib.require_cuda(ys, x)
    sfx = _lib.suffix(x.dtype)
    C, T, d = x.shape
    x, ys = x.contiguous(), ys.contiguous()
    acc = out is not None
    if out is None:
        out = torch.empty((C,), dtype=x.dtype, device=dev)
    ys_stride = 0 if ys.dim() == 2 else T * d
    _lib.call("auxssm_sv_log_potential_" + sfx,
        _lib.stream_ptr(dev), C, T * d, _lib.ptr(ys), ys_stride, _lib.ptr(x), _lib.ptr(out), 1 if acc else 0)
    return out


def get_kernel(ys, m0, P0, F, Q, b, parallel, order=1, *, time_chunks=0, dense_R=False, prior_cancels=True):
    """auxiliary_kalman.py:10-57: auxiliary Kalman sampler for the SV model, first or second order.

    The SV target is the proposal's own latent dynamics times the observation potential, so the ``log p(x)`` terms of
    the acceptance ratio cancel (``kalman.get_kernel(prior_cancels=True)``): by default neither the two
    ``prior_logpdf`` calls of ``log_likelihood_fn`` (:50-54) nor the prior part of the two ``posterior_logpdf`` calls are
    evaluated.  ``prior_cancels=False`` evaluates all four like the reference does."""
    d = m0.shape[0]
    # the SV dynamics are F = phi I (model.py:35-43): hint the structured kernels.  The hint is verified on the device
    # on every call (include/auxssm.h), so a dense F passed here still gives the general kernels' results.
    f_diagonal = True
    Fs, Qs, bs = F.reshape(1, d, d), Q.reshape(1, d, d), b.reshape(1, d)  # time-invariant: stride 0, never tiled

    def dynamics_factory(_x):
        return m0, P0, Fs, Qs, bs

    def obs_factory(x, u, delta):
        return observations_factory(ys, x, u, delta, order, dense_R)

    prior_lgssm = kalman.LGSSM(m0, P0, Fs, Qs, bs, None, None, None)

    def log_likelihood_fn(x):
        if prior_cancels:
            return log_potential(ys, x)
        out = kalman.prior_logpdf(x, prior_lgssm)  # mvn.logpdf(x0; m0, P0) + sum_t mvn.logpdf(x_t; F x_{t-1} + b, Q)
        return log_potential(ys, x, out=out)

    return kalman.get_kernel(dynamics_factory, obs_factory, log_likelihood_fn, parallel, time_chunks=time_chunks,
                             f_diagonal=f_diagonal, prior_cancels=prior_cancels)

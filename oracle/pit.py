"""NumPy restatement of the reference's parallel-in-time conditional SMC (oracle only; test infrastructure).

Follows, for the closed model family of ``oracle/csmc.py:LinearGaussianModel``:

* ``aux_samplers/_primitives/csmc/pit/csmc.py:68-114`` (``_csmc``: all proposals at once, particle 0 pinned to the
  reference trajectory, initial weights ``Qt.logpdf - Mt.logpdf`` (+ ``G0`` at t = 0) normalised per time step),
* ``aux_samplers/_primitives/csmc/pit/operator.py:39-149`` (stitching operator: the N x N weights
  ``Gt(x_t^j, x_{t-1}^i) + log w_{t-1}^i + log w_t^j``, conditional multinomial resampling of N pairs out of N^2
  -- a single pair at the last level --, gather of both halves, weights reset to 1/N),
* ``aux_samplers/_primitives/csmc/pit/dc_map.py:73-123`` (divide and conquer over T padded to a power of two;
  pairs that reach into the padding are passed through unchanged),
* ``aux_samplers/csmc/independent.py:78-118,212-231,239-248`` (the auxiliary factory: proposals
  ``N(u_t + s_t^2 grad_t, s_t^2 I)``, ``Qt = N(u_t, s_t^2 I)`` when ``gradient``, ``G0' = G0 + log M0``,
  ``Gt' = log Mt(x_t | x_{t-1}) + Gt``).

Draws are explicit: ``normals [T, N, d]`` (``Mt.sample`` per time step, pit/csmc.py:75) and ``uniforms [T, N]``:
row t serves the stitch whose right-hand block starts at time t (``keys_b[0]``, operator.py:78,81); the last level
draws a single pair and uses ``uniforms[t, 0]``.  Trajectories are never copied: like the reference's ``origins``
(operator.py:97-99) the state of a block is the index of the ORIGINAL particle each of its N trajectories visits
at every time step, and ``x_t^n = xs[t, origins[t, n]]``.

Pinned by tests/golden/ref_pit.npz (the reference's own source executed on oracle/refshim).
"""
import numpy as np

from .csmc import canonical_max, canonical_sum, categorical, multinomial
from .linalg import norm_logpdf


def _lse(lw):
    """logsumexp over the last axis with the canonical sum."""
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        amax = canonical_max(lw)
        amax = np.where(np.isfinite(amax), amax, 0).astype(lw.dtype)
        return (np.log(canonical_sum(np.exp(lw - amax[..., None]))) + amax).astype(lw.dtype)


def pit_init(model, x_star, u, s, grad, gradient, normals):
    """Proposals and per-step normalised initial log-weights (pit/csmc.py:75-92)."""
    dtype = normals.dtype
    s = np.asarray(s, dtype=dtype)
    mean = (u + (s ** 2)[:, None] * grad).astype(dtype) if gradient else u.astype(dtype)
    xs = (mean[:, None, :] + s[:, None, None] * normals).astype(dtype)
    xs[:, 0] = x_star
    if gradient:
        lw = (np.sum(norm_logpdf(xs, u[:, None, :], s[:, None, None]), axis=-1)
              - np.sum(norm_logpdf(xs, mean[:, None, :], s[:, None, None]), axis=-1)).astype(dtype)
    else:
        lw = np.zeros(xs.shape[:2], dtype=dtype)
    lw[0] = lw[0] + model.pot(0, xs[0]) + model.m0_logpdf(xs[0])
    lw = lw - _lse(lw)[:, None]
    return xs, lw.astype(dtype)


def stitch_log_weights(model, t, xa, lwa, xb, lwb):
    """[N, N] stitching log-weights (operator.py:118-149): entry (i, j) pairs left trajectory i (ending in xa[i] at
    time t - 1) with right trajectory j (starting in xb[j] at time t)."""
    N = xa.shape[0]
    inc = np.empty((N, N), dtype=xa.dtype)
    pot = model.pot(t, xb)
    for i in range(N):
        inc[i] = model.mt_logpdf(xb, xa[i]) + pot
    return (inc + lwa[:, None] + lwb[None, :]).astype(xa.dtype)


def pit_csmc(model, x_star, u, s, grad, gradient, normals, uniforms, return_levels=False):
    """One parallel-in-time cSMC sweep for ONE chain.  Returns (x [T, d], ancestors [T]); ``updated`` of
    pit/csmc.py:59 is ``ancestors != 0``."""
    T, N, d = normals.shape
    dtype = normals.dtype
    xs, lw = pit_init(model, x_star, u, s, grad, gradient, normals)
    origins = np.tile(np.arange(N), (T, 1))
    K = 0
    while (1 << K) < T:
        K += 1
    levels = []
    for k in range(K):
        size = 1 << k
        last = k == K - 1
        picks = {}
        for a0 in range(0, 1 << K, 2 * size):
            t = a0 + size                       # first time step of the right-hand block
            if t >= T:                          # dc_map.py:101: both blocks must hold real time steps
                continue
            b1 = min(t + size, T)
            xa, xb = xs[t - 1][origins[t - 1]], xs[t][origins[t]]
            lww = stitch_log_weights(model, t, xa, lw[t - 1], xb, lw[t]).reshape(-1)
            with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
                w = np.exp(lww - _lse(lww)).astype(dtype)
            if last:
                idx = np.asarray([categorical(w, np.asarray(uniforms[t, 0], dtype=dtype))])
            else:
                idx = multinomial(w, np.asarray(uniforms[t], dtype=dtype))
            l, r = idx // N, idx % N
            origins_a, origins_b = origins[a0:t][:, l], origins[t:b1][:, r]
            if last:
                origins = np.concatenate([origins_a, origins_b])[:, 0]
            else:
                origins[a0:t], origins[t:b1] = origins_a, origins_b
                lw[a0:b1] = dtype.type(-np.log(N))
            picks[t] = (l, r)
        levels.append(picks)
    x = xs[np.arange(T), origins]
    if return_levels:
        return x, origins, levels, xs
    return x, origins

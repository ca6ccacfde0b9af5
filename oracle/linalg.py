"""Small batched dense linear algebra in NumPy with XLA-like NaN semantics (oracle only).

``numpy.linalg.cholesky`` raises on non-PD input whereas ``jnp.linalg.cholesky``
returns NaNs, and the reference relies on the latter (``filtering.py:113-116``,
``sampling.py:102-104``).  These helpers are vectorised over leading axes.
"""
import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


def chol(a):
    """Lower Cholesky factor, column by column; non-PD pivots give NaN (not an exception).

    Mirrors ``jnp.linalg.cholesky`` on well-posed inputs.  Infinite diagonal entries
    (the reference's missing-observation trick, ``filtering.py:93-98``) propagate as
    inf on the diagonal and 0 off it, like LAPACK ``potrf`` does.
    """
    a = np.array(a, copy=True)
    n = a.shape[-1]
    L = np.zeros_like(a)
    with np.errstate(invalid="ignore", divide="ignore"):
        for k in range(n):
            s = a[..., k, k] - np.sum(L[..., k, :k] ** 2, axis=-1)
            lkk = np.sqrt(s)  # NaN when s < 0
            L[..., k, k] = lkk
            if k + 1 < n:
                r = a[..., k + 1:, k] - np.einsum("...ij,...j->...i", L[..., k + 1:, :k], L[..., k, :k])
                L[..., k + 1:, k] = r / lkk[..., None]
    return L


def solve_lower(L, b, vec=None):
    """Solve L x = b for lower-triangular L; b is (..., n) [vec=True] or (..., n, m)."""
    if vec is None:
        vec = b.ndim == L.ndim - 1
    if vec:
        b = b[..., None]
    n = L.shape[-1]
    x = np.zeros(np.broadcast_shapes(L.shape[:-2], b.shape[:-2]) + b.shape[-2:], dtype=np.result_type(L, b))
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n):
            acc = b[..., i, :] - np.einsum("...k,...km->...m", L[..., i, :i], x[..., :i, :])
            x[..., i, :] = acc / L[..., i, i, None]
    return x[..., 0] if vec else x


def solve_upper_from_lower(L, b):
    """Solve L^T x = b for lower-triangular L."""
    vec = b.ndim == L.ndim - 1
    if vec:
        b = b[..., None]
    n = L.shape[-1]
    x = np.zeros(np.broadcast_shapes(L.shape[:-2], b.shape[:-2]) + b.shape[-2:], dtype=np.result_type(L, b))
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n - 1, -1, -1):
            acc = b[..., i, :] - np.einsum("...k,...km->...m", L[..., i + 1:, i], x[..., i + 1:, :])
            x[..., i, :] = acc / L[..., i, i, None]
    return x[..., 0] if vec else x


def cho_solve(L, b):
    """Solve (L L^T) x = b, like ``jax.scipy.linalg.cho_solve((L, True), b)``."""
    return solve_upper_from_lower(L, solve_lower(L, b))


def mT(a):
    return np.swapaxes(a, -1, -2)


def sym(a):
    return 0.5 * (a + mT(a))


def mv(A, x):
    return np.einsum("...ij,...j->...i", A, x)


def tril_log_det(chol_):
    """Reference ``math/mvn/base.py:108-128``: non-finite diagonal entries count as 1."""
    d = np.diagonal(chol_, axis1=-2, axis2=-1)
    d = np.where(np.isfinite(d), d, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.nansum(np.log(np.abs(d)), axis=-1)


def mvn_logpdf(x, m, chol_):
    """Reference ``math/mvn/base.py:16-58`` (Cholesky-parametrised, NaN/inf tolerant).

    ``_INF = 1e500`` there is the Python float ``inf``; so is the clip range.
    """
    diag = np.diagonal(chol_, axis1=-2, axis2=-1)
    dim = np.sum(np.isfinite(diag), axis=-1)
    clip = np.where(np.isfinite(chol_), chol_, np.inf)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        y = solve_lower(clip, x - m, vec=True)
        const = tril_log_det(chol_) + 0.5 * dim * LOG_2PI
        return -0.5 * np.sum(y * y, axis=-1) - const


def norm_logpdf(x, loc, scale):
    """``jax.scipy.stats.norm.logpdf``."""
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        z = (x - loc) / scale
        return -0.5 * z * z - np.log(scale) - 0.5 * LOG_2PI


def logsumexp(a, axis=-1):
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        amax = np.max(a, axis=axis, keepdims=True)
        amax = np.where(np.isfinite(amax), amax, 0.0)
        out = np.log(np.sum(np.exp(a - amax), axis=axis, keepdims=True)) + amax
    return np.squeeze(out, axis=axis)


def normalize(log_w, axis=-1):
    """Reference ``math/utils.py:24-39``: exp(lw - logsumexp(lw))."""
    with np.errstate(invalid="ignore", over="ignore"):
        return np.exp(log_w - np.expand_dims(logsumexp(log_w, axis=axis), axis))

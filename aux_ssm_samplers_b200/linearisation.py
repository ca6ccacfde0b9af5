"""Linearisation of conditional moments: host mirror of ``aux_samplers/_primitives/linearisation.py``.

``cubature(mean, cov, params, x_star, P_star)`` (:78-107) and ``gauss_hermite(..., order=3)`` (:48-75) return
``(F, Q, b)`` like the reference.  Differences forced by the platform: ``x_star [..., dx]`` / ``P_star [..., dx, dx]``
carry any number of leading batch axes (the reference is ``jax.vmap``-ed over time by its callers), and ``mean`` /
``cov`` must be BATCHED callables ``mean(points [..., n_pts, dx], params) -> [..., n_pts, dy]``,
``cov(points, params) -> [..., n_pts, dy, dy]`` (or ``[dy, dy]`` for a constant covariance) -- what
``jax.vmap(mean, in_axes=[0, None])`` (:116) does to the reference's per-point callables.  The sigma points and the
moment fit run in ``auxssm_sigma_points_*`` / ``auxssm_sigma_fit_*``; the two user callables are the only code
evaluated in between.  ``extended`` (:12-45) needs a Jacobian of the user's callable: ``torch.func.jacfwd / jacrev``.
"""
import math

import numpy as np
import torch

from . import _lib


def _cubature_points(n_dim):
    """Spherical cubature rule, linearisation.py:226-247: (w [2n], xi [n, 2n])."""
    w = np.ones((2 * n_dim,)) / (2 * n_dim)
    xi = np.concatenate([np.eye(n_dim), -np.eye(n_dim)], axis=0) * np.sqrt(n_dim)
    return w, xi.T


def _hermite_coeff(order):
    H = [np.array([1]), np.array([2, 0])]
    for i in range(2, order + 1):
        H.append(2 * np.append(H[i - 1], 0) - 2 * (i - 1) * np.pad(H[i - 2], (2, 0), "constant", constant_values=0))
    return H


def _gauss_hermite_points(n_dim, order=3):
    """Gauss--Hermite product rule, linearisation.py:142-196: (w [p^n], xi [n, p^n]).  Host NumPy, computed once."""
    n, p = n_dim, order
    coeff = _hermite_coeff(p)
    roots = np.flip(np.roots(coeff[-1]))
    w_1d = np.array([2 ** (p - 1) * math.factorial(p) * np.sqrt(np.pi) / (p ** 2 * np.polyval(coeff[p - 1], roots[i]) ** 2)
                     for i in range(p)])
    table = np.zeros((n, p ** n))
    for i in range(n):
        base = np.ones((1, p ** (n - i - 1)))
        for j in range(1, p):
            base = np.concatenate([base, (j + 1) * np.ones((1, p ** (n - i - 1)))], axis=1)
        table[n - i - 1, :] = np.tile(base, (1, int(p ** i)))
    table = table.astype("int64") - 1
    return np.prod(w_1d[table], axis=0) / (np.sqrt(np.pi) ** n), math.sqrt(2) * roots[table]


def _generic_sigma_points(mean, cov, params, x_star, P_star, rule):
    dev = _lib.require_cuda(x_star, P_star)
    dtype = x_star.dtype
    sfx = _lib.suffix(dtype)
    batch, dx = x_star.shape[:-1], x_star.shape[-1]
    n = int(np.prod(batch)) if batch else 1
    w_np, xi_np = rule(dx)
    n_pts = w_np.shape[0]
    w = torch.as_tensor(w_np, dtype=dtype, device=dev).contiguous()
    xi = torch.as_tensor(np.ascontiguousarray(xi_np), dtype=dtype, device=dev)
    xs = x_star.reshape(n, dx).contiguous()
    Ps = P_star.reshape(n, dx, dx).contiguous()
    points = torch.empty((n, n_pts, dx), dtype=dtype, device=dev)
    chol = torch.empty((n, dx, dx), dtype=dtype, device=dev)
    st = _lib.stream_ptr(dev)
    _lib.call("auxssm_sigma_points_" + sfx, st, n, dx, n_pts, _lib.ptr(xs), _lib.ptr(Ps), _lib.ptr(xi), _lib.ptr(points),
              _lib.ptr(chol))
    pts_b = points.reshape(batch + (n_pts, dx))
    f_pts = mean(pts_b, params)
    dy = f_pts.shape[-1]
    f_pts = f_pts.to(dtype).reshape(n, n_pts, dy).contiguous()
    v = cov(pts_b, params)
    _lib.require_cuda(xs, f_pts, v)
    v = v.to(dtype)
    if v.dim() == 2:                       # constant covariance [dy, dy]
        v, vis, vps = v.contiguous(), 0, 0
    else:
        v = v.reshape(n, n_pts, dy, dy).contiguous()
        vis, vps = n_pts * dy * dy, dy * dy
    F = torch.empty((n, dy, dx), dtype=dtype, device=dev)
    Q = torch.empty((n, dy, dy), dtype=dtype, device=dev)
    b = torch.empty((n, dy), dtype=dtype, device=dev)
    _lib.call("auxssm_sigma_fit_" + sfx, st, n, dx, dy, n_pts, _lib.ptr(w), _lib.ptr(points), _lib.ptr(xs), _lib.ptr(chol),
              _lib.ptr(f_pts), _lib.ptr(v), vis, vps, _lib.ptr(F), _lib.ptr(Q), _lib.ptr(b))
    return F.reshape(batch + (dy, dx)), Q.reshape(batch + (dy, dy)), b.reshape(batch + (dy,))


def cubature(mean, cov, params, x_star, P_star):
    """Spherical cubature linearisation (linearisation.py:78-107)."""
    return _generic_sigma_points(mean, cov, params, x_star, P_star, _cubature_points)


def gauss_hermite(mean, cov, params, x_star, P_star, order=3):
    """Gauss--Hermite linearisation (linearisation.py:48-75); ``order ** dx`` sigma points."""
    return _generic_sigma_points(mean, cov, params, x_star, P_star, lambda dim: _gauss_hermite_points(dim, order))


def extended(mean, cov, params, x_star, _P_star=None):
    """First-order Taylor linearisation (linearisation.py:12-45) of an UNBATCHED ``mean(x [dx], params) -> [dy]``;
    leading batch axes of ``x_star`` are mapped with ``torch.func.vmap``.  The Jacobian of a user callable is the
    user's autodiff (``jax.jacfwd / jacrev`` in the reference, ``torch.func`` here), not a library kernel."""
    from torch.func import jacfwd, jacrev, vmap
    dx = x_star.shape[-1]
    flat = x_star.reshape(-1, dx)
    b0 = vmap(lambda x: mean(x, params))(flat)
    dy = b0.shape[-1]
    jac = jacrev if dy < dx else jacfwd
    F = vmap(jac(lambda x: mean(x, params)))(flat)
    Q = vmap(lambda x: cov(x, params))(flat)
    b = b0 - torch.einsum("nij,nj->ni", F, flat)
    batch = x_star.shape[:-1]
    return F.reshape(batch + (dy, dx)), Q.reshape(batch + (dy, dy)), b.reshape(batch + (dy,))

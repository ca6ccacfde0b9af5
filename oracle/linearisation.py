"""NumPy restatement of the reference's sigma-point linearisation (oracle only; test infrastructure).

Follows ``aux_samplers/_primitives/linearisation.py``: ``_generic_sigma_points`` (:110-133), ``_cov`` (:136-139),
``_cubature_points`` (:226-247), ``_gauss_hermite_points`` (:142-196) / ``_hermite_coeff`` (:199-223) and the
first-order ``extended`` rule (:12-45, with the Jacobian supplied by the caller).  Pinned by the reference's own
test (``test_linearisation.py:17-51``: a linear model is recovered exactly by all three rules) and by
tests/golden/ref_linearisation.npz (the reference's source executed on oracle/refshim for a non-linear model).
"""
import math

import numpy as np


def cubature_points(n_dim):
    """(w [2n], xi [n, 2n]) of the spherical cubature rule (:226-247)."""
    w = np.ones((2 * n_dim,)) / (2 * n_dim)
    xi = np.concatenate([np.eye(n_dim), -np.eye(n_dim)], axis=0) * np.sqrt(n_dim)
    return w, xi.T


def _hermite_coeff(order):
    H = [np.array([1]), np.array([2, 0])]
    for i in range(2, order + 1):
        H.append(2 * np.append(H[i - 1], 0) - 2 * (i - 1) * np.pad(H[i - 2], (2, 0), "constant", constant_values=0))
    return H


def gauss_hermite_points(n_dim, order=3):
    """(w [p^n], xi [n, p^n]) of the Gauss--Hermite product rule (:142-196)."""
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
    w = np.prod(w_1d[table], axis=0) / (np.sqrt(np.pi) ** n)
    xi = math.sqrt(2) * roots[table]
    return w, xi


def sigma_points(x_star, P_star, xi):
    """points [n_pts, dx] = x* + (chol(P*) xi)^T (:111-114)."""
    chol = np.linalg.cholesky(P_star)
    return x_star[None, :] + (chol @ xi).T, chol


def sigma_fit(w, points, x_star, chol, f_pts, v_pts):
    """(F, Q, b) from the function values at the sigma points (:116-133); v_pts is [n_pts, dy, dy]."""
    m_f = w @ f_pts
    Psi_x = ((points - x_star[None, :]).T * w[None, :]) @ (f_pts - m_f[None, :])
    y = np.linalg.solve(chol, Psi_x)
    F_x = np.linalg.solve(chol.T, y).T
    v_f = np.sum(w[:, None, None] * v_pts, 0)
    Phi = ((f_pts - m_f[None, :]).T * w[None, :]) @ (f_pts - m_f[None, :])
    temp = F_x @ chol
    return F_x, Phi - temp @ temp.T + v_f, m_f - F_x @ x_star


def linearise(mean, cov, params, x_star, P_star, rule="cubature", order=3):
    dim = x_star.shape[0]
    w, xi = cubature_points(dim) if rule == "cubature" else gauss_hermite_points(dim, order)
    points, chol = sigma_points(x_star, P_star, xi)
    f_pts = np.stack([mean(p, params) for p in points])
    v_pts = np.stack([cov(p, params) for p in points])
    return sigma_fit(w, points, x_star, chol, f_pts, v_pts)


def extended(mean, jac, cov, params, x_star):
    """First-order rule (:12-45) with the Jacobian ``jac(x, params)`` supplied (the reference uses jax.jacfwd/jacrev)."""
    F = jac(x_star, params)
    return F, cov(x_star, params), mean(x_star, params) - F @ x_star

"""NumPy restatement of the reference's stochastic-volatility example pieces (oracle only).

Follows ``aux_samplers/examples/stochastic_volatility/model.py:12-82`` and
``auxiliary_kalman.py:10-57``.  Parity unpinned: the reference has no test for these.
"""
import numpy as np

from .linalg import chol, mv, mvn_logpdf, norm_logpdf


def stationary_covariance(phi, tau, rho, dim):
    """model.py:46-53."""
    U = tau * rho * np.ones((dim, dim))
    U[np.diag_indices(dim)] = tau
    return U / (1.0 - phi ** 2)


def get_dynamics(nu, phi, tau, rho, dim):
    """model.py:35-43."""
    F = phi * np.eye(dim)
    Q = stationary_covariance(phi, tau, rho, dim)
    mu = nu * np.ones((dim,))
    b = mu + F @ mu
    return mu, Q.copy(), F, Q, b


def get_data(rng, nu, phi, tau, rho, dim, T):
    """model.py:12-31 with explicit normals from ``rng``."""
    m0, P0, F, Q, b = get_dynamics(nu, phi, tau, rho, dim)
    cP0, cQ = np.linalg.cholesky(P0), np.linalg.cholesky(Q)
    x = m0 + cP0 @ rng.standard_normal(dim)
    xs, ys = np.empty((T, dim)), np.empty((T, dim))
    for t in range(T):
        xs[t] = x
        ys[t] = np.exp(0.5 * x) * rng.standard_normal(dim)
        x = F @ x + b + cQ @ rng.standard_normal(dim)
    return xs, ys


def log_potential_terms(x, y):
    """model.py:72-76: nan_to_num(log N(y; 0, exp(x/2)))."""
    return np.nan_to_num(norm_logpdf(y, 0.0, np.exp(0.5 * x)))


def log_potential(xs, ys):
    """model.py:56-59, summed over time and dimensions (leading chain axes kept)."""
    return np.sum(log_potential_terms(xs, ys), axis=(-2, -1))


def grad_hess(x, y):
    """d/dx and d2/dx2 of log N(y; 0, exp(x/2)) = -y^2 e^{-x}/2 - x/2 - const."""
    w = 0.5 * y * y * np.exp(-x)
    return w - 0.5, -w


def make_factories(ys, m0, P0, F, Q, b, order):
    """auxiliary_kalman.py:10-57 -> (dynamics_factory, observations_factory, log_likelihood_fn)."""
    T, d = ys.shape
    eye = np.eye(d)
    cP0, cQ = chol(P0), chol(Q)

    def dynamics_factory(_x):
        return (m0, P0, np.broadcast_to(F, (T - 1, d, d)), np.broadcast_to(Q, (T - 1, d, d)),
                np.broadcast_to(b, (T - 1, d)))

    def observations_factory(x, u, delta):
        grad, hess = grad_hess(x, ys)
        lead = x.shape[:-2]
        Hs = np.broadcast_to(eye, lead + (T, d, d))
        cs = np.zeros(lead + (T, d))
        if order == 1:
            aux = u + 0.5 * delta * np.nan_to_num(grad)
            Rs = np.broadcast_to(0.5 * delta * eye, lead + (T, d, d))
            return aux, Hs, Rs, cs
        omega = 1.0 / (-hess + 2.0 / delta)  # Hessian is diagonal (model.py:66-69)
        aux = omega * (2.0 * u / delta + grad - hess * x)
        Rs = omega[..., :, None] * eye
        return aux, Hs, Rs, cs

    def log_likelihood_fn(x):
        out = mvn_logpdf(x[..., 0, :], m0, cP0)
        pred = mv(F, x[..., :-1, :]) + b
        out = out + np.sum(mvn_logpdf(x[..., 1:, :], pred, cQ), axis=-1)
        return out + log_potential(x, ys)

    return dynamics_factory, observations_factory, log_likelihood_fn

"""Pins the Kalman oracle (oracle/kalman.py) the way the reference's own tests pin the
reference: against an explicit masked-row Kalman filter and an RTS smoother
(reference aux_samplers/_primitives/test_kalman/test_filtering.py:20-107,
test_sampling.py:23-127)."""
import numpy as np
import numpy.testing as npt
import pytest
from scipy.linalg import block_diag
from scipy.stats import multivariate_normal

from oracle import kalman as ok
from oracle.explicit import kalman_filter_masked, rts_smoother
from oracle.linalg import mvn_logpdf


def random_lgssm(rng, T, dx, dy, lead=()):
    m0 = rng.standard_normal(lead + (dx,))
    P0 = rng.standard_normal(lead + (dx, 5 * dx))
    P0 = P0 @ np.swapaxes(P0, -1, -2)
    Fs = rng.standard_normal(lead + (T - 1, dx, dx))
    Qs = rng.standard_normal(lead + (T - 1, dx, 5 * dx))
    Qs = Qs @ np.swapaxes(Qs, -1, -2)
    bs = rng.standard_normal(lead + (T - 1, dx))
    Hs = rng.standard_normal(lead + (T, dy, dx))
    Rs = rng.standard_normal(lead + (T, dy, 5 * dy))
    Rs = Rs @ np.swapaxes(Rs, -1, -2)
    cs = rng.standard_normal(lead + (T, dy))
    ys = rng.standard_normal(lead + (T, dy))
    return ys, (m0, P0, Fs, Qs, bs, Hs, Rs, cs)


@pytest.mark.parametrize("seed", [0, 1234])
@pytest.mark.parametrize("T", [5, 7])
@pytest.mark.parametrize("dx", [1, 2])
@pytest.mark.parametrize("dy", [1, 3])
@pytest.mark.parametrize("parallel", [False, True])
def test_vs_explicit_with_nans(seed, T, dx, dy, parallel):
    # reference test_filtering.py:20-55 (which only runs parallel=False; we pin both modes)
    rng = np.random.default_rng(seed)
    ys, lg = random_lgssm(rng, T, dx, dy)
    ys[1, :] = np.nan
    ys[3, 0] = np.nan
    lg[5][3, 0, :] = np.nan
    ms, Ps, ell = ok.filtering(ys, lg, parallel)
    m0, P0, Fs, Qs, bs, Hs, Rs, cs = lg
    ems, ePs, eell = kalman_filter_masked(ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs)
    npt.assert_allclose(ms, ems, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(Ps, ePs, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(ell, eell, rtol=1e-7)


@pytest.mark.parametrize("seed", [0, 1234])
@pytest.mark.parametrize("T", [3, 5])
@pytest.mark.parametrize("dx", [1, 2])
@pytest.mark.parametrize("dy", [1, 3])
@pytest.mark.parametrize("parallel", [True, False])
def test_batched_model(seed, T, dx, dy, parallel):
    # reference test_filtering.py:58-107: B independent sub-models == block-diagonal model
    B = 3
    rng = np.random.default_rng(seed)
    bys, blg = random_lgssm(rng, T, dx, dy, lead=(B,))
    bms, bPs, bell = ok.filtering(bys, blg, parallel)
    bd = lambda a: np.stack([block_diag(*a[:, t]) for t in range(a.shape[1])])
    m0 = blg[0].reshape(B * dx)
    P0 = block_diag(*blg[1])
    Fs, Qs, Hs, Rs = bd(blg[2]), bd(blg[3]), bd(blg[5]), bd(blg[6])
    bs = np.swapaxes(blg[4], 0, 1).reshape(T - 1, B * dx)
    cs = np.swapaxes(blg[7], 0, 1).reshape(T, B * dy)
    ys = np.swapaxes(bys, 0, 1).reshape(T, B * dy)
    ms, Ps, ell = ok.filtering(ys, (m0, P0, Fs, Qs, bs, Hs, Rs, cs), parallel)
    ems, ePs, eell = kalman_filter_masked(ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs)
    npt.assert_allclose(ms, ems, rtol=1e-7, atol=1e-10)
    npt.assert_allclose(Ps, ePs, rtol=1e-7, atol=1e-10)
    npt.assert_allclose(ell, eell, rtol=1e-7)
    npt.assert_allclose(np.swapaxes(bms, 0, 1).reshape(T, B * dx), ems, rtol=1e-7, atol=1e-10)
    npt.assert_allclose(np.sum(bell), eell, rtol=1e-7)
    for b in range(B):
        npt.assert_allclose(bPs[b], ePs[:, b * dx:(b + 1) * dx, b * dx:(b + 1) * dx], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("dx,dy", [(1, 1), (2, 3), (3, 2)])
def test_parallel_equals_sequential_sampling_exactly(dx, dy):
    # same eps => same trajectory in both modes (reference sampling.py:28-39)
    rng = np.random.default_rng(3)
    T = 9
    ys, lg = random_lgssm(rng, T, dx, dy, lead=(4,))
    ms, Ps, _ = ok.filtering(ys, lg, False)
    eps = rng.standard_normal((4, T, dx))
    xs_seq = ok.sampling(eps, ms, Ps, lg, False)
    xs_par = ok.sampling(eps, ms, Ps, lg, True)
    npt.assert_allclose(xs_seq, xs_par, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("parallel", [True, False])
def test_sampling_moments_vs_rts(parallel):
    # reference test_sampling.py:23-68 (500k draws, 1e-2) -- 200k draws, slightly looser
    rng = np.random.default_rng(0)
    T, dx, dy, n = 5, 2, 2, 200_000
    ys, lg = random_lgssm(rng, T, dx, dy)
    lg = list(lg)
    lg[2] = 0.5 * lg[2]
    ms, Ps, _ = ok.filtering(ys, lg, False)
    sm, sP = rts_smoother(ms, Ps, lg[2], lg[3], lg[4])
    eps = rng.standard_normal((n, T, dx))
    xs = ok.sampling(eps, np.broadcast_to(ms, (n,) + ms.shape), np.broadcast_to(Ps, (n,) + Ps.shape), lg, parallel)
    npt.assert_allclose(xs.mean(0), sm, atol=3e-2, rtol=3e-2)
    for t in range(T):
        npt.assert_allclose(np.cov(xs[:, t].T), sP[t], atol=5e-2, rtol=5e-2)


def test_mvn_logpdf_doctest():
    # reference math/mvn/base.py:37-46
    z, mu = np.array([1.0, 2.0, 3.0]), np.array([2.0, 3.0, 4.0])
    L = np.array([[1, 0, 0], [0.2, 1.3, 0], [0.123, -0.5, 1.7]])
    npt.assert_allclose(mvn_logpdf(z, mu, L), multivariate_normal.logpdf(z, mu, L @ L.T))


def test_posterior_logpdf_is_a_density_ratio():
    # unpinned by the reference; check against scipy on the joint Gaussian
    rng = np.random.default_rng(5)
    T, dx, dy = 4, 2, 2
    ys, lg = random_lgssm(rng, T, dx, dy)
    m0, P0, Fs, Qs, bs, Hs, Rs, cs = lg
    xs = rng.standard_normal((T, dx))
    _, _, ell = ok.filtering(ys, lg, False)
    lp = ok.posterior_logpdf(ys, xs, ell, lg)
    expect = multivariate_normal.logpdf(xs[0], m0, P0)
    for t in range(1, T):
        expect += multivariate_normal.logpdf(xs[t], Fs[t - 1] @ xs[t - 1] + bs[t - 1], Qs[t - 1])
    for t in range(T):
        expect += multivariate_normal.logpdf(ys[t], Hs[t] @ xs[t] + cs[t], Rs[t])
    npt.assert_allclose(lp, expect - ell, rtol=1e-9)

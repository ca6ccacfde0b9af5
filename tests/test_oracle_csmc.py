"""Pins the cSMC oracle (oracle/csmc.py) the way the reference's own tests pin the reference
(aux_samplers/_primitives/test_csmc/test_resamplings.py:11-66, test_csmc.py:18-69)."""
import numpy as np
import numpy.testing as npt
import pytest

from oracle import csmc as oc
from oracle.explicit import conditional_systematic_alg4


def test_canonical_reductions_match_numpy_closely():
    rng = np.random.default_rng(0)
    for n in (1, 7, 32, 33, 100, 1024):
        w = rng.uniform(size=(3, n))
        npt.assert_allclose(oc.canonical_cdf(w), np.cumsum(w, -1), rtol=1e-13)
        npt.assert_allclose(oc.canonical_sum(w), w.sum(-1), rtol=1e-13)
        w32 = w.astype(np.float32)
        assert oc.canonical_cdf(w32).dtype == np.float32
        npt.assert_allclose(oc.canonical_cdf(w32), np.cumsum(w, -1), rtol=2e-6)


def test_multinomial_resampling():
    # reference test_resamplings.py:11-24 -- frequencies match the weights, index 0 pinned
    rng = np.random.default_rng(42)
    w = rng.uniform(size=10)
    w /= w.sum()
    u = rng.uniform(size=(100_000, 10))
    idx = oc.multinomial(np.broadcast_to(w, u.shape), u)
    counts = np.bincount(idx[:, 1:].ravel(), minlength=10)
    npt.assert_allclose(counts / counts.sum(), w, atol=3e-3)
    assert (idx[:, 0] == 0).all()


@pytest.mark.parametrize("seed", [42, 43, 44])
@pytest.mark.parametrize("M", [100, 1000])
def test_systematic_resampling_vs_alg4(seed, M):
    # reference test_resamplings.py:27-66
    rng = np.random.default_rng(seed)
    N = 100
    w = rng.uniform(size=M) ** 2
    w /= w.sum()
    for _ in range(100):
        u, v, ww = rng.uniform(size=3)
        got = oc.systematic(w, u, v, ww, N)
        want = conditional_systematic_alg4(w, u, v, ww, N)
        npt.assert_array_equal(got, want)
        assert got[0] == 0


def test_killing_keeps_heavy_particles_and_pins_zero():
    rng = np.random.default_rng(1)
    w = rng.uniform(size=(2000, 16))
    w /= w.sum(-1, keepdims=True)
    A = oc.killing(w, rng.uniform(size=w.shape), rng.uniform(size=w.shape))
    assert (A[:, 0] == 0).all()
    heaviest = w.argmax(-1)
    rows = np.arange(w.shape[0])
    keep = heaviest != 0
    assert (A[rows[keep], heaviest[keep]] == heaviest[keep]).all()  # w_i / max w = 1 => always survives
    # unbiasedness of the offspring numbers: E[#offspring of j] = N w_j for j >= 1 ... checked loosely
    w1 = np.full((20000, 8), 1 / 8)
    A1 = oc.killing(w1, rng.uniform(size=w1.shape), rng.uniform(size=w1.shape))
    assert (A1 == np.arange(8)).all()  # equal weights: nobody is killed


@pytest.mark.parametrize("backward", [True, False])
def test_flat_potential_recovers_ar1_prior(backward):
    # reference test_csmc.py:18-69 (50 000 iterations, atol 0.05) -- 6 000 iterations, atol 0.1
    rng = np.random.default_rng(3)
    T, N, rho, n_iter = 5, 32, 0.9, 6000
    model = oc.LinearGaussianModel([0.0], [[1.0]], [[rho]], [[1 - rho ** 2]], [0.0], "flat")
    fk = oc.BootstrapFK(model)
    x = np.zeros((T, 1))
    out = np.empty((n_iter, T))
    for i in range(n_iter):
        x, _, _, _ = oc.csmc_kernel(fk, model, x, rng.standard_normal((T, N, 1)), rng.uniform(size=(T - 1, N)),
                                    rng.uniform(size=T), backward)
        out[i] = x[:, 0]
    out = out[n_iter // 10:]
    npt.assert_allclose(out.mean(0), 0.0, atol=0.1)
    npt.assert_allclose(out.var(0), 1.0, atol=0.12)
    lag1 = np.mean(out[:, 1:] * out[:, :-1], axis=0)
    npt.assert_allclose(lag1, rho, atol=0.12)


def test_star_trajectory_is_particle_zero_and_ancestor_zero():
    rng = np.random.default_rng(5)
    T, N, d = 7, 16, 2
    model = oc.LinearGaussianModel(np.zeros(d), np.eye(d), 0.8 * np.eye(d), 0.36 * np.eye(d), np.zeros(d), "sv",
                                   ys=rng.standard_normal((T, d)))
    x_star = rng.standard_normal((T, d))
    w_T, xs, log_ws, As = oc.csmc_forward(oc.BootstrapFK(model), x_star, rng.standard_normal((T, N, d)),
                                          rng.uniform(size=(T - 1, N)))
    npt.assert_array_equal(xs[:, 0], x_star)
    assert (As[:, 0] == 0).all()
    npt.assert_allclose(w_T.sum(), 1.0, rtol=1e-12)


def test_guided_and_independent_weights_are_importance_ratios():
    """For any proposal the (proposal density x weight) must equal prior transition x potential."""
    rng = np.random.default_rng(6)
    T, d = 4, 3
    Q = 0.3 * np.eye(d) + 0.1
    model = oc.LinearGaussianModel(np.zeros(d), Q / 0.19, 0.9 * np.eye(d), Q, np.zeros(d), "sv",
                                   ys=rng.standard_normal((T, d)))
    u = rng.standard_normal((T, d))
    s = np.full(T, 0.4)
    x_prev, x = rng.standard_normal((5, d)), rng.standard_normal((5, d))
    target = model.mt_logpdf(x, x_prev) + model.pot(2, x)
    g = oc.GuidedAuxFK(model, u, s)
    mu = g._mu(2, model.mt_mean(x_prev))
    aux = np.sum(oc.norm_logpdf(x, u[2], s[2]), -1)
    npt.assert_allclose(g.gt(2, x, x_prev) + oc.mvn_logpdf(x, mu, g.cL[2]), target + aux, rtol=1e-10)
    ind = oc.IndependentAuxFK(model, u, s)
    npt.assert_allclose(ind.gt(2, x, x_prev), target, rtol=1e-12)

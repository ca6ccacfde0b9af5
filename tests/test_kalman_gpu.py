"""Parity of the CUDA Kalman path (through the C ABI) against the oracle on identical inputs.

Tolerances (BASELINE.json north_star): relative 1e-5 for fp64, 1e-3 for fp32.
"""
import numpy as np
import numpy.testing as npt
import pytest
import torch

from oracle import kalman as ok
from oracle import sv as osv

pytestmark = pytest.mark.gpu

TOL = {torch.float64: dict(rtol=1e-5, atol=1e-7), torch.float32: dict(rtol=1e-3, atol=2e-3)}


def dev(a, dtype):
    return None if a is None else torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype)


def host(t):
    return t.detach().cpu().double().numpy()


def random_lgssm(rng, T, dx, dy, lead=(), stable=False):
    m0 = rng.standard_normal(lead + (dx,))
    P0 = rng.standard_normal(lead + (dx, 5 * dx))
    P0 = P0 @ np.swapaxes(P0, -1, -2) / (5 * dx if stable else 1)
    Fs = rng.standard_normal(lead + (T - 1, dx, dx))
    if stable:
        Fs = 0.9 * np.eye(dx) + 0.1 * Fs / np.sqrt(dx)
    Qs = rng.standard_normal(lead + (T - 1, dx, 5 * dx))
    Qs = Qs @ np.swapaxes(Qs, -1, -2) / (5 * dx if stable else 1)
    bs = rng.standard_normal(lead + (T - 1, dx))
    Hs = rng.standard_normal(lead + (T, dy, dx))
    Rs = rng.standard_normal(lead + (T, dy, 5 * dy))
    Rs = Rs @ np.swapaxes(Rs, -1, -2) / (5 * dy if stable else 1)
    cs = rng.standard_normal(lead + (T, dy))
    ys = rng.standard_normal(lead + (T, dy))
    return ys, (m0, P0, Fs, Qs, bs, Hs, Rs, cs)


def run_filter(ys, lg, dtype, parallel, **kw):
    from aux_ssm_samplers_b200 import kalman
    ms, Ps, ell = kalman.filtering(dev(ys, dtype), [dev(a, dtype) for a in lg], parallel, **kw)
    torch.cuda.synchronize()
    return host(ms), host(Ps), host(ell)


@pytest.mark.parametrize("seed", [0, 1234])
@pytest.mark.parametrize("T", [5, 7])
@pytest.mark.parametrize("dx", [1, 2])
@pytest.mark.parametrize("dy", [1, 3])
@pytest.mark.parametrize("mode", ["seq", "par2", "par3", "parmax"])
def test_filter_small_with_nans(seed, T, dx, dy, mode):
    # the reference's own test shapes and NaN pattern (test_filtering.py:20-55)
    rng = np.random.default_rng(seed)
    ys, lg = random_lgssm(rng, T, dx, dy)
    ys[1, :] = np.nan
    ys[3, 0] = np.nan
    lg[5][3, 0, :] = np.nan
    ems, ePs, eell = ok.filtering(ys, lg, False)
    kw = dict(seq=dict(parallel=False), par2=dict(parallel=True, time_chunks=2),
              par3=dict(parallel=True, time_chunks=3), parmax=dict(parallel=True, time_chunks=T))[mode]
    ms, Ps, ell = run_filter(ys, lg, torch.float64, **kw)
    npt.assert_allclose(ms, ems, **TOL[torch.float64])
    npt.assert_allclose(Ps, ePs, **TOL[torch.float64])
    npt.assert_allclose(ell, eell, **TOL[torch.float64])


@pytest.mark.parametrize("parallel", [False, True])
@pytest.mark.parametrize("dx,dy", [(1, 1), (2, 3)])
def test_filter_inner_batch_layout(parallel, dx, dy):
    # reference test_filtering.py:58-107: batched [T, B, ...] LGSSM, ell summed over B
    rng = np.random.default_rng(7)
    T, B = 5, 3
    bys, blg = random_lgssm(rng, T, dx, dy, lead=(B,))
    ems, ePs, eell = ok.filtering(bys, blg, False)
    sw = lambda a, time: np.swapaxes(a, 0, 1) if time else a
    lg_tb = [blg[0], blg[1]] + [sw(a, True) for a in blg[2:]]
    ms, Ps, ell = run_filter(sw(bys, True), lg_tb, torch.float64, parallel, inner_batch=True,
                             time_chunks=2 if parallel else 0)
    npt.assert_allclose(np.swapaxes(ms, 0, 1), ems, **TOL[torch.float64])
    npt.assert_allclose(np.swapaxes(Ps, 0, 1), ePs, **TOL[torch.float64])
    npt.assert_allclose(ell, eell.sum(), **TOL[torch.float64])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,T,C", [(2, 128, 1), (8, 40, 3), (16, 33, 2), (30, 64, 4), (64, 12, 2)])
@pytest.mark.parametrize("chunks", [1, 4])
@pytest.mark.parametrize("dense_h", [False, True])
def test_filter_mid_sizes(dtype, d, T, C, chunks, dense_h):
    if d == 64 and dtype == torch.float64 and (dense_h or chunks > 1):
        pytest.skip("d=64 fp64 with dense H / parallel exceeds shared memory (documented)")
    rng = np.random.default_rng(d * 1000 + T)
    ys, lg = random_lgssm(rng, T, d, d, lead=(C,), stable=True)
    lg = list(lg)
    if not dense_h:
        lg[5] = np.broadcast_to(np.eye(d), (C, T, d, d)).copy()
        lg[7] = np.zeros((C, T, d))
    ys[0, T // 2, :] = np.nan
    ys[-1, 3, 0] = np.nan
    ems, ePs, eell = ok.filtering(ys, lg, False)
    glg = list(lg)
    if not dense_h:
        glg[5] = None
        glg[7] = None
    ms, Ps, ell = run_filter(ys, glg, dtype, chunks > 1, time_chunks=chunks)
    scale = np.abs(ePs).max()
    npt.assert_allclose(ms, ems, rtol=TOL[dtype]["rtol"], atol=TOL[dtype]["rtol"] * max(1.0, np.abs(ems).max()))
    npt.assert_allclose(Ps, ePs, rtol=TOL[dtype]["rtol"], atol=TOL[dtype]["rtol"] * scale)
    npt.assert_allclose(ell, eell, rtol=TOL[dtype]["rtol"])


def test_filter_time_invariant_broadcast():
    # stride-0 F, Q, b, H, R shared over time and chains
    rng = np.random.default_rng(11)
    T, d, C = 20, 5, 6
    ys, lg = random_lgssm(rng, T, d, d, stable=True)
    lg = [lg[0], lg[1]] + [a[:1] for a in lg[2:]]
    ys = rng.standard_normal((C, T, d))
    full = [lg[0], lg[1]] + [np.broadcast_to(a, (T - 1,) + a.shape[1:]) for a in lg[2:5]] + \
           [np.broadcast_to(a, (T,) + a.shape[1:]) for a in lg[5:]]
    ems, ePs, eell = ok.filtering(ys, full, False)
    ms, Ps, ell = run_filter(ys, lg, torch.float64, False)
    npt.assert_allclose(ms, ems, **TOL[torch.float64])
    npt.assert_allclose(Ps, ePs, **TOL[torch.float64])
    npt.assert_allclose(ell, eell, **TOL[torch.float64])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,T,C", [(1, 9, 3), (2, 128, 1), (5, 17, 4), (30, 48, 3), (64, 10, 2)])
@pytest.mark.parametrize("chunks", [1, 3])
def test_sampling_vs_oracle(dtype, d, T, C, chunks):
    if d == 64 and dtype == torch.float64 and chunks > 1:
        pytest.skip("d=64 fp64 parallel sampling exceeds shared memory (documented)")
    from aux_ssm_samplers_b200 import kalman
    rng = np.random.default_rng(d + T)
    ys, lg = random_lgssm(rng, T, d, d, lead=(C,), stable=True)
    ms, Ps, _ = ok.filtering(ys, lg, False)
    eps = rng.standard_normal((C, T, d))
    exs = ok.sampling(eps, ms, Ps, lg, False)
    xs = kalman.sampling(dev(eps, dtype), dev(ms, dtype), dev(Ps, dtype), [dev(a, dtype) for a in lg],
                         chunks > 1, time_chunks=chunks)
    torch.cuda.synchronize()
    r = TOL[dtype]["rtol"]
    npt.assert_allclose(host(xs), exs, rtol=r, atol=r * max(1.0, np.abs(exs).max()))


def test_sampling_zero_covariance_last_step():
    # P_T = 0: Cholesky fails, nan_to_num -> 0 (sampling.py:102-104,121)
    from aux_ssm_samplers_b200 import kalman
    rng = np.random.default_rng(3)
    T, d = 6, 3
    ys, lg = random_lgssm(rng, T, d, d, stable=True)
    ms, Ps, _ = ok.filtering(ys, lg, False)
    Ps[-1] = 0.0
    eps = rng.standard_normal((T, d))
    exs = ok.sampling(eps, ms, Ps, lg, False)
    xs = kalman.sampling(dev(eps, torch.float64), dev(ms, torch.float64), dev(Ps, torch.float64),
                         [dev(a, torch.float64) for a in lg], False)
    npt.assert_allclose(host(xs), exs, **TOL[torch.float64])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("dx,dy,T,C", [(1, 1, 7, 2), (2, 3, 9, 3), (30, 30, 40, 2)])
def test_logpdfs_vs_oracle(dtype, dx, dy, T, C):
    from aux_ssm_samplers_b200 import kalman
    rng = np.random.default_rng(dx * 10 + dy)
    ys, lg = random_lgssm(rng, T, dx, dy, lead=(C,), stable=True)
    ys[0, 2, :] = np.nan  # a missing observation: its likelihood term is dropped by nansum
    xs = rng.standard_normal((C, T, dx))
    _, _, ell = ok.filtering(ys, lg, False)
    glg = [dev(a, dtype) for a in lg]
    r = TOL[dtype]["rtol"]
    for name, got, want in [
        ("prior", kalman.prior_logpdf(dev(xs, dtype), glg), ok.prior_logpdf(xs, lg)),
        ("loglik", kalman.log_likelihood(dev(ys, dtype), dev(xs, dtype), glg), ok.log_likelihood(ys, xs, lg)),
        ("post", kalman.posterior_logpdf(dev(ys, dtype), dev(xs, dtype), dev(ell, dtype), glg),
         ok.posterior_logpdf(ys, xs, ell, lg)),
    ]:
        npt.assert_allclose(host(got), want, rtol=r, err_msg=name)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("order,seed", [(1, 43), (2, 44)])
@pytest.mark.parametrize("parallel", [False, True])
@pytest.mark.parametrize("dense_R", [False, True])
def test_sv_sweep_vs_oracle(dtype, order, seed, parallel, dense_R):
    """Full auxiliary-Kalman MH sweep on the SV model with identical externally supplied draws."""
    from aux_ssm_samplers_b200 import stochastic_volatility as sv
    from aux_ssm_samplers_b200.draws import ExternalDraws
    rng = np.random.default_rng(seed)
    T, d, C, delta = 24, 4, 6, 0.3
    m0, P0, F, Q, b = osv.get_dynamics(0.0, 0.9, 2.0, 0.25, d)
    xs_true, ys = osv.get_data(rng, 0.0, 0.9, 2.0, 0.25, d, T)
    x = xs_true[None] + 0.3 * rng.standard_normal((C, T, d))  # chains near the posterior bulk
    n_aux, n_eps, un = rng.standard_normal((C, T, d)), rng.standard_normal((C, T, d)), rng.uniform(size=C)
    facs = osv.make_factories(ys, m0, P0, F, Q, b, order)
    ex, eacc, aux = ok.kalman_sweep(x, delta, n_aux, n_eps, un, *facs, parallel)
    _, kernel = sv.get_kernel(dev(ys, dtype), dev(m0, dtype), dev(P0, dtype), dev(F, dtype), dev(Q, dtype),
                              dev(b, dtype), parallel, order, time_chunks=3 if parallel else 0, dense_R=dense_R)
    from aux_ssm_samplers_b200.kalman import KalmanSampler
    state = kernel(ExternalDraws(normals=[n_aux, n_eps], uniforms=[un]), KalmanSampler(dev(x, dtype), True), delta)
    torch.cuda.synchronize()
    # alpha is bounded away from the uniform draws with overwhelming probability for these seeds
    margin = np.abs(aux["alpha"] - un).min()
    assert margin > 1e-3, "degenerate test draw"
    npt.assert_array_equal(host(state.updated).astype(bool), eacc)
    r = TOL[dtype]["rtol"]
    npt.assert_allclose(host(state.x), ex, rtol=r, atol=r * np.abs(ex).max())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,T,C", [(3, 11, 2), (30, 40, 3), (64, 9, 2)])
@pytest.mark.parametrize("dense_h", [False, True])
@pytest.mark.parametrize("chunks", [1, 3])
@pytest.mark.parametrize("fkind", ["dense", "diag", "diag-shared"])
def test_filter_and_logpdf_diagonal_R(dtype, d, T, C, dense_h, chunks, fkind):
    """AUXSSM_FLAG_R_DIAGONAL: diag(R_t) passed as [C, T, d] must equal the dense-R oracle."""
    from aux_ssm_samplers_b200 import kalman
    if d == 64 and dtype == torch.float64 and (dense_h or chunks > 1):
        pytest.skip("d=64 fp64 with dense H / parallel exceeds shared memory (documented)")
    rng = np.random.default_rng(d + T + chunks)
    ys, lg = random_lgssm(rng, T, d, d, lead=(C,), stable=True)
    lg = list(lg)
    rdiag = rng.uniform(0.05, 1.0, size=(C, T, d))
    lg[6] = rdiag[..., :, None] * np.eye(d)
    if fkind == "diag":          # time-varying diagonal F: dense predict path
        lg[2] = rng.uniform(0.5, 0.95, size=(C, T - 1, d))[..., :, None] * np.eye(d)
    elif fkind == "diag-shared":  # time-invariant diagonal F: O(d^2) predict on the register tile
        lg[2] = np.broadcast_to(np.diag(rng.uniform(0.5, 0.95, size=d)), (C, T - 1, d, d)).copy()
    if not dense_h:
        lg[5] = np.broadcast_to(np.eye(d), (C, T, d, d)).copy()
        lg[7] = np.zeros((C, T, d))
    ys[0, 2, 1] = np.nan
    ys[-1, 4, :] = np.nan
    ems, ePs, eell = ok.filtering(ys, lg, False)
    glg = [dev(a, dtype) for a in lg]
    glg[6] = kalman.Diagonal(dev(rdiag, dtype))
    if fkind == "diag-shared":
        glg[2] = dev(lg[2][:1, :1], dtype)  # [1, 1, d, d]: stride 0 over chains and time
    if not dense_h:
        glg[5] = glg[7] = None
    ms, Ps, ell = kalman.filtering(dev(ys, dtype), glg, chunks > 1, time_chunks=chunks,
                                   f_diagonal=(fkind == "diag-shared"))
    r = TOL[dtype]["rtol"]
    npt.assert_allclose(host(ms), ems, rtol=r, atol=r * max(1.0, np.abs(ems).max()))
    npt.assert_allclose(host(Ps), ePs, rtol=r, atol=r * np.abs(ePs).max())
    npt.assert_allclose(host(ell), eell, rtol=r)
    xs = rng.standard_normal((C, T, d))
    got = kalman.posterior_logpdf(dev(ys, dtype), dev(xs, dtype), dev(eell, dtype), glg)
    npt.assert_allclose(host(got), ok.posterior_logpdf(ys, xs, eell, lg), rtol=r)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,T,C", [(1, 5, 2), (4, 300, 3), (30, 130, 2), (64, 140, 2)])
def test_logpdf_fast_path_time_invariant(dtype, d, T, C):
    """Time-invariant (stride-0) dynamics + H = I + diagonal R take the thread-per-time-step kernel (prior / posterior)
    and the row-per-warp kernel (likelihood alone)."""
    from aux_ssm_samplers_b200 import kalman
    rng = np.random.default_rng(100 + d)
    _, lg = random_lgssm(rng, T, d, d, stable=True)
    m0, P0, F, Q, b = lg[0], lg[1], lg[2][:1], lg[3][:1], lg[4][:1]
    rdiag = rng.uniform(0.05, 1.0, size=(C, T, d))
    ys = rng.standard_normal((C, T, d))
    ys[0, 1, 0] = np.nan
    xs = rng.standard_normal((C, T, d))
    ell = rng.standard_normal(C)
    full = (m0, P0, np.broadcast_to(F, (T - 1, d, d)), np.broadcast_to(Q, (T - 1, d, d)),
            np.broadcast_to(b, (T - 1, d)), np.broadcast_to(np.eye(d), (T, d, d)),
            rdiag[..., :, None] * np.eye(d), np.zeros((T, d)))
    glg = kalman.LGSSM(dev(m0, dtype), dev(P0, dtype), dev(F, dtype), dev(Q, dtype), dev(b, dtype), None,
                       kalman.Diagonal(dev(rdiag, dtype)), None)
    r = TOL[dtype]["rtol"]
    npt.assert_allclose(host(kalman.prior_logpdf(dev(xs, dtype), glg)), ok.prior_logpdf(xs, full), rtol=r)
    npt.assert_allclose(host(kalman.log_likelihood(dev(ys, dtype), dev(xs, dtype), glg)),
                        ok.log_likelihood(ys, xs, full), rtol=r)
    npt.assert_allclose(host(kalman.posterior_logpdf(dev(ys, dtype), dev(xs, dtype), dev(ell, dtype), glg)),
                        ok.posterior_logpdf(ys, xs, ell, full), rtol=r)
    # likelihood-only call with a time-invariant offset c (stride 0) and an infinite variance entry (dropped from the
    # dimension count, mvn/base.py:49-57; -inf through norm.logpdf when d = 1, kalman/base.py:160-162): the
    # row-per-warp kernel
    c0 = rng.standard_normal(d)
    rdiag2 = rdiag.copy()
    rdiag2[C - 1, 2, d - 1] = np.inf
    Rd = np.zeros((C, T, d, d))
    Rd[..., np.arange(d), np.arange(d)] = rdiag2  # (inf * eye would put NaN off the diagonal)
    full_c = full[:6] + (Rd, np.broadcast_to(c0, (T, d)))
    glg_c = glg._replace(Rs=kalman.Diagonal(dev(rdiag2, dtype)), cs=dev(c0[None, None], dtype))
    with np.errstate(all="ignore"):
        want = ok.log_likelihood(ys, xs, full_c)
    npt.assert_allclose(host(kalman.log_likelihood(dev(ys, dtype), dev(xs, dtype), glg_c)), want, rtol=r)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,T,C", [(3, 9, 2), (30, 33, 3)])
def test_sampling_time_invariant_diagonal_F(dtype, d, T, C):
    from aux_ssm_samplers_b200 import kalman
    rng = np.random.default_rng(d)
    ys, lg = random_lgssm(rng, T, d, d, lead=(C,), stable=True)
    lg = list(lg)
    F = np.diag(rng.uniform(0.5, 0.95, size=d))
    lg[2] = np.broadcast_to(F, (C, T - 1, d, d)).copy()
    ms, Ps, _ = ok.filtering(ys, lg, False)
    eps = rng.standard_normal((C, T, d))
    exs = ok.sampling(eps, ms, Ps, lg, False)
    glg = [dev(a, dtype) for a in lg]
    glg[2] = dev(F[None], dtype)
    for hint in (False, True):
        xs = kalman.sampling(dev(eps, dtype), dev(ms, dtype), dev(Ps, dtype), glg, False, f_diagonal=hint)
        r = TOL[dtype]["rtol"]
        npt.assert_allclose(host(xs), exs, rtol=r, atol=r * max(1.0, np.abs(exs).max()))

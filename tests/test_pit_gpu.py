"""Parallel-in-time conditional SMC on the GPU (auxssm_pit_csmc_sweep_*, csmc.get_independent_kernel(parallel=True))
against (1) fixtures produced by the reference's own source (tests/golden/ref_pit.npz: csmc/independent.py:78-118 ->
_primitives/csmc/pit/{csmc,operator,dc_map}.py executed on oracle/refshim) and (2) the NumPy oracle (oracle/pit.py,
itself pinned by the same fixtures) with identical draws.  fp64: ancestors bit-exact, trajectories 1e-5 relative;
fp32: 1e-3 relative on the chains whose index sequence agrees -- the N^2-entry CDF passes through exp / log whose last
bit differs between NumPy and CUDA, so an index may flip when a uniform lies within a few ulp of one of the N^2 knots;
a majority of chains must agree end to end (small N) and every chain must be a valid gather of its own particles."""
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

from oracle import csmc as oc
from oracle import pit as opit

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = {torch.float64: 1e-5, torch.float32: 1e-3}
NP = {torch.float64: np.float64, torch.float32: np.float32}


def dev(a, dtype):
    return None if a is None else torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype)


def host(t):
    return t.detach().cpu().numpy()


def family(model, dtype):
    from aux_ssm_samplers_b200 import csmc
    M0 = csmc.GaussianDistribution(dev(model.m0, dtype), dev(model.P0, dtype))
    Mt = csmc.LinearGaussianDynamics(dev(model.F, dtype), dev(model.Q, dtype), dev(model.b, dtype))
    if model.potential == "flat":
        G = csmc.FlatPotential()
    elif model.potential == "gaussian":
        G = csmc.GaussianObservationPotential(dev(model.ys, dtype), model.sig)
    else:
        G = csmc.StochasticVolatilityPotential(dev(model.ys, dtype))
    return M0, G, Mt, G


def check(dtype, x_gpu, upd_gpu, x_ref, upd_ref, need_all):
    C = x_ref.shape[0]
    r = RTOL[dtype]
    ok = 0
    for c in range(C):
        same = np.array_equal(upd_gpu[c], upd_ref[c])
        if same:
            npt.assert_allclose(x_gpu[c], x_ref[c], rtol=r, atol=r * np.abs(x_ref[c]).max())
            ok += 1
        else:
            assert not need_all, "chain %d: updated flags differ" % c
    assert ok == C if need_all else 2 * ok > C, "only %d of %d chains agree" % (ok, C)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("case", ["a", "b", "c"])
@pytest.mark.parametrize("gradient", [False, True])
def test_pit_kernel_vs_reference_fixture(dtype, case, gradient):
    from aux_ssm_samplers_b200 import csmc
    from aux_ssm_samplers_b200.draws import ExternalDraws
    z = np.load(os.path.join(GOLD, "ref_pit.npz"))
    model = oc.LinearGaussianModel(*[z["%s_%s" % (case, k)] for k in ("m0", "P0", "F", "Q", "b")], "sv", ys=z[case + "_ys"])
    tag = "%s_g%d_" % (case, gradient)
    N = int(z[case + "_N"])
    M0, G0, Mt, Gt = family(model, dtype)
    init, kernel = csmc.get_independent_kernel(M0, G0, Mt, Gt, N, gradient=gradient, parallel=True)
    key = ExternalDraws([z[tag + "normals_aux"], z[tag + "normals"]], [z[tag + "uniforms"]])
    st = kernel(key, init(dev(z[case + "_x_star"], dtype)), dev(z[case + "_delta"], dtype))
    check(dtype, host(st.x), host(st.updated), z[tag + "x"], z[tag + "updated"], need_all=dtype == torch.float64)


def make_model(rng, d, T, potential, dtype):
    A = rng.standard_normal((d, d))
    Q = (A @ A.T / d + 0.3 * np.eye(d)) * 0.4
    F = 0.8 * np.eye(d) + 0.1 * rng.standard_normal((d, d)) / np.sqrt(d)
    b = 0.1 * rng.standard_normal(d)
    m0 = 0.2 * rng.standard_normal(d)
    ys = rng.standard_normal((T, d)) if potential != "flat" else None
    return oc.LinearGaussianModel(m0, Q * 2.0, F, Q, b, potential, ys=ys, sig=0.7, dtype=NP[dtype])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,N,T,potential,gradient", [
    (1, 2, 2, "gaussian", 0), (1, 32, 33, "sv", 1), (2, 64, 16, "sv", 0), (4, 100, 9, "gaussian", 1),
    (30, 128, 7, "sv", 0), (32, 33, 6, "flat", 1), (5, 128, 64, "sv", 1)])
def test_pit_sweep_vs_oracle(dtype, d, N, T, potential, gradient):
    """Direct C-ABI sweep with supplied draws vs oracle/pit.py: padded T, T a power of two, N^2 up to 16384 (three
    CDF levels), N not a multiple of 32, d = 32 (largest), every potential."""
    from aux_ssm_samplers_b200 import csmc
    from aux_ssm_samplers_b200.draws import ExternalDraws
    rng = np.random.default_rng(1000 * d + N + T)
    C = 3
    npd = NP[dtype]
    model = make_model(rng, d, T, potential, dtype)
    x_star = rng.standard_normal((C, T, d)).astype(npd)
    s = np.sqrt(0.5 * (0.2 + 0.3 * rng.uniform(size=T))).astype(npd)
    u = (x_star + s[None, :, None] * rng.standard_normal((C, T, d))).astype(npd)
    normals = rng.standard_normal((C, T, N, d)).astype(npd)
    uniforms = rng.uniform(size=(C, T, N)).astype(npd)
    grad = np.stack([oc.grad_log_target(model, u[c]) for c in range(C)]).astype(npd) if gradient else None
    M0, G0, Mt, Gt = family(model, dtype)
    x_gpu, anc_gpu = csmc.pit_sweep(dev(x_star, dtype), M0, Mt, Gt, N, ExternalDraws([normals], [uniforms]), dev(u, dtype),
                                    dev(np.broadcast_to(s, (C, T)), dtype), dev(grad, dtype))
    x_gpu, anc_gpu = host(x_gpu), host(anc_gpu)
    assert np.isfinite(x_gpu).all() and (anc_gpu >= 0).all() and (anc_gpu < N).all()
    agree = 0
    for c in range(C):
        g = grad[c] if gradient else np.zeros_like(u[c])
        x, anc, levels, xs = opit.pit_csmc(model, x_star[c], u[c], s, g, gradient, normals[c], uniforms[c], return_levels=True)
        # whatever the index sequence, the output must be a gather of this chain's own particles
        npt.assert_allclose(x_gpu[c], xs[np.arange(T), anc_gpu[c]], rtol=RTOL[dtype], atol=RTOL[dtype])
        if np.array_equal(anc_gpu[c], anc):
            agree += 1
            npt.assert_allclose(x_gpu[c], x, rtol=RTOL[dtype], atol=RTOL[dtype] * np.abs(x).max())
        else:
            assert dtype == torch.float32, "fp64 ancestors differ from the oracle (chain %d)" % c
    if dtype == torch.float64:
        assert agree == C
    elif N <= 33:
        assert 2 * agree > C


def test_in_kernel_philox_draws_equal_the_filled_arrays():
    """The sweep with draws generated inside the kernels equals the sweep fed with the arrays PhiloxDraws fills."""
    from aux_ssm_samplers_b200 import csmc
    from aux_ssm_samplers_b200.draws import PhiloxDraws
    rng = np.random.default_rng(5)
    dtype, d, N, T, C = torch.float64, 3, 40, 13, 4
    model = make_model(rng, d, T, "sv", dtype)
    M0, G0, Mt, Gt = family(model, dtype)
    x_star = dev(rng.standard_normal((C, T, d)), dtype)
    init, kernel = csmc.get_independent_kernel(M0, G0, Mt, Gt, N, gradient=True, parallel=True)
    outs = []
    for in_kernel in (True, False):
        key = PhiloxDraws(99, in_kernel=in_kernel)
        st = init(x_star)
        for _ in range(2):
            st = kernel(key, st, 0.4)
        outs.append((host(st.x), host(st.updated)))
    npt.assert_array_equal(outs[0][1], outs[1][1])
    npt.assert_array_equal(outs[0][0], outs[1][0])


def test_pit_gibbs_recovers_the_smoothing_distribution():
    """Stationarity (the reference's test_pit_csmc.py:18-62 only plots): AR(1) prior + Gaussian observations, many
    independent chains of the parallel-in-time auxiliary kernel; means / variances of x_t against the exact RTS
    smoother."""
    from aux_ssm_samplers_b200 import csmc
    from aux_ssm_samplers_b200.draws import PhiloxDraws
    rho, sig_y, T, N, C = 0.9, 0.5, 12, 32, 4096
    rng = np.random.default_rng(3)
    xs = np.zeros(T)
    xs[0] = rng.standard_normal()
    for t in range(1, T):
        xs[t] = rho * xs[t - 1] + np.sqrt(1 - rho ** 2) * rng.standard_normal()
    ys = xs + sig_y * rng.standard_normal(T)
    # exact smoother (scalar Kalman filter + RTS)
    m, P = np.zeros(T), np.zeros(T)
    mp, Pp = 0.0, 1.0
    for t in range(T):
        if t > 0:
            mp, Pp = rho * m[t - 1], rho ** 2 * P[t - 1] + 1 - rho ** 2
        k = Pp / (Pp + sig_y ** 2)
        m[t], P[t] = mp + k * (ys[t] - mp), (1 - k) * Pp
    ms, Ps = m.copy(), P.copy()
    for t in range(T - 2, -1, -1):
        Pp = rho ** 2 * P[t] + 1 - rho ** 2
        g = P[t] * rho / Pp
        ms[t] = m[t] + g * (ms[t + 1] - rho * m[t])
        Ps[t] = P[t] + g ** 2 * (Ps[t + 1] - Pp)
    f = lambda v: torch.tensor(v, dtype=torch.float32, device="cuda")
    M0 = csmc.GaussianDistribution(f([0.0]), f([[1.0]]))
    Mt = csmc.LinearGaussianDynamics(f([[rho]]), f([[1 - rho ** 2]]), f([0.0]))
    G = csmc.GaussianObservationPotential(f(ys[:, None]), sig_y)
    init, kernel = csmc.get_independent_kernel(M0, G, Mt, G, N, gradient=True, parallel=True)
    state = init(torch.zeros((C, T, 1), dtype=torch.float32, device="cuda"))
    key = PhiloxDraws(7)
    acc = 0.0
    for _ in range(80):
        state = kernel(key, state, 0.8)
        acc += state.updated.float().mean().item()
    x = state.x[..., 0].double().cpu().numpy()
    assert acc / 80 > 0.1  # fraction of time steps whose particle changed, averaged over sweeps
    npt.assert_allclose(x.mean(0), ms, atol=0.05)
    npt.assert_allclose(x.var(0), Ps, atol=0.05)


def test_pit_rejects_unsupported_shapes():
    from aux_ssm_samplers_b200 import csmc
    from aux_ssm_samplers_b200._lib import AuxssmError
    from aux_ssm_samplers_b200.draws import PhiloxDraws
    f = lambda v: torch.tensor(v, dtype=torch.float64, device="cuda")
    M0 = csmc.GaussianDistribution(f([0.0]), f([[1.0]]))
    Mt = csmc.LinearGaussianDynamics(f([[0.5]]), f([[0.75]]), f([0.0]))
    G = csmc.FlatPotential()
    init, kernel = csmc.get_independent_kernel(M0, G, Mt, G, 129, parallel=True)
    with pytest.raises(AuxssmError):
        kernel(PhiloxDraws(1), init(torch.zeros((2, 4, 1), dtype=torch.float64, device="cuda")), 0.3)
    with pytest.raises(NotImplementedError):
        csmc.get_independent_kernel(M0, G, Mt, G, 8, gradient="per_particle", parallel=True)

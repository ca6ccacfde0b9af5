"""Sigma-point / extended linearisation on the GPU (auxssm_sigma_points_*, auxssm_sigma_fit_*, linearisation.py)
against the reference-generated fixture tests/golden/ref_linearisation.npz (reference source on oracle/refshim), the
reference's own pin (test_linearisation.py:17-51: a linear model is recovered exactly) and the NumPy oracle at the
largest supported size.  Tolerances: 1e-5 (fp64) / 1e-3 (fp32), relative to the largest entry."""
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

from oracle import linearisation as ol

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = {torch.float64: 1e-5, torch.float32: 1e-3}


def dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype)


def close(got, want, dtype):
    r = RTOL[dtype]
    npt.assert_allclose(got.detach().cpu().numpy(), want, rtol=r, atol=r * max(np.abs(want).max(), 1e-30))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_linearisation_vs_reference_fixture(dtype, tag):
    from aux_ssm_samplers_b200 import linearisation as lin
    z = np.load(os.path.join(GOLD, "ref_linearisation.npz"))
    W, c, Q0 = dev(z[tag + "_W"], dtype), dev(z[tag + "_c"], dtype), dev(z[tag + "_Q0"], dtype)
    mean = lambda x, _p: torch.sigmoid(x @ W.T) + c * (x ** 2).sum(-1, keepdim=True) * 0.1
    cov = lambda x, _p: Q0 * (1.0 + 0.2 * torch.exp(-x ** 2).sum(-1))[..., None, None]
    xs, Ps = dev(z[tag + "_x"], dtype), dev(z[tag + "_P"], dtype)
    for name, out in (("cub", lin.cubature(mean, cov, None, xs, Ps)), ("gh", lin.gauss_hermite(mean, cov, None, xs, Ps, 3))):
        for got, k in zip(out, ("F", "Q", "b")):
            close(got, z["%s_%s_%s" % (tag, name, k)], dtype)
    mean1 = lambda x, _p: torch.sigmoid(W @ x) + c * (x ** 2).sum() * 0.1
    cov1 = lambda x, _p: Q0 * (1.0 + 0.2 * torch.exp(-x ** 2).sum())
    for got, k in zip(lin.extended(mean1, cov1, None, xs, Ps), ("F", "Q", "b")):
        close(got, z["%s_ext_%s" % (tag, k)], dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_all_rules_recover_a_linear_model(dtype):
    """test_linearisation.py:17-51 of the reference, batched over 7 expansion points."""
    from aux_ssm_samplers_b200 import cubature, extended, gauss_hermite
    rng = np.random.default_rng(1)
    A, b = rng.standard_normal((2, 4)), rng.standard_normal(2)
    q = rng.standard_normal((2, 5))
    Q = q @ q.T
    xs = rng.standard_normal((7, 4))
    p = rng.standard_normal((7, 4, 10))
    Ps = p @ np.swapaxes(p, -1, -2)
    At, bt, Qt = dev(A, dtype), dev(b, dtype), dev(Q, dtype)
    mean, cov = (lambda x, _p: x @ At.T + bt), (lambda x, _p: Qt)
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    outs = [cubature(mean, cov, None, dev(xs, dtype), dev(Ps, dtype)), gauss_hermite(mean, cov, None, dev(xs, dtype), dev(Ps, dtype)),
            extended(lambda x, _p: At @ x + bt, lambda x, _p: Qt, None, dev(xs, dtype), None)]
    for F_, Q_, b_ in outs:
        npt.assert_allclose(F_.cpu().numpy(), np.broadcast_to(A, (7, 2, 4)), atol=tol * 10)
        npt.assert_allclose(Q_.cpu().numpy(), np.broadcast_to(Q, (7, 2, 2)), atol=tol * 200)
        npt.assert_allclose(b_.cpu().numpy(), np.broadcast_to(b, (7, 2)), atol=tol * 30)


def test_largest_size_and_batch_axes_vs_oracle():
    from aux_ssm_samplers_b200 import linearisation as lin
    rng = np.random.default_rng(2)
    dtype, dx, dy = torch.float64, 32, 32
    W = rng.standard_normal((dy, dx)) / np.sqrt(dx)
    Q0 = np.eye(dy) * 0.3
    xs = rng.standard_normal((2, 3, dx))
    p = rng.standard_normal((2, 3, dx, 2 * dx))
    Ps = p @ np.swapaxes(p, -1, -2) / (2 * dx)
    Wt, Qt = dev(W, dtype), dev(Q0, dtype)
    F, Q, b = lin.cubature(lambda x, _p: torch.tanh(x @ Wt.T), lambda x, _p: Qt, None, dev(xs, dtype), dev(Ps, dtype))
    assert F.shape == (2, 3, dy, dx) and Q.shape == (2, 3, dy, dy) and b.shape == (2, 3, dy)
    for i in range(2):
        for j in range(3):
            Fo, Qo, bo = ol.linearise(lambda x, _p: np.tanh(W @ x), lambda x, _p: Q0, None, xs[i, j], Ps[i, j], "cubature")
            close(F[i, j], Fo, dtype); close(Q[i, j], Qo, dtype); close(b[i, j], bo, dtype)


def test_unsupported_dimension_is_an_error():
    from aux_ssm_samplers_b200 import AuxssmError, cubature
    x = torch.zeros((33,), dtype=torch.float64, device="cuda")
    with pytest.raises(AuxssmError):
        cubature(lambda p, _: p, lambda p, _: torch.eye(33, dtype=torch.float64, device="cuda"), None, x,
                 torch.eye(33, dtype=torch.float64, device="cuda"))

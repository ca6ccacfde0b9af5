"""Pins the oracle against the reference's OWN runnable code, read from /root/reference at test time.

These tests only run where the reference tree exists (the build container); on the GPU box, which has no
/root/reference, they skip -- the fixtures they correspond to are committed under tests/golden/ (ref_*.npz) and
checked by tests/test_reference_golden.py without the tree.

1. `_primitives/test_kalman/common.py` needs only numpy + scipy and is loaded by file path: the oracle's filter
   (sequential AND parallel, with the NaN pattern of test_filtering.py:45-47, and the batched model of :58-107) and
   sampler moments (test_sampling.py:23-68) are checked against `explicit_kalman_filter` / `explicit_kalman_smoothing`.
2. The reference's own pytest modules are executed, unmodified, on the NumPy stand-in for jax (oracle/refshim) in a
   subprocess: this validates the stand-in that produced the committed fixtures.
3. The committed fixtures are regenerated in a temporary directory and compared with the committed ones.
"""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import numpy.testing as npt
import pytest
from scipy.linalg import block_diag

from oracle import kalman as ok

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = os.path.join(REF, "aux_samplers", "_primitives", "test_kalman", "common.py")
pytestmark = pytest.mark.skipif(not os.path.exists(COMMON), reason="reference tree not present on this machine")


def ref_common():
    spec = importlib.util.spec_from_file_location("ref_test_kalman_common", COMMON)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_style_model(T, dx, dy, lead=()):
    """The construction of test_filtering.py:29-43 (np.random.seed is set by the caller)."""
    m0 = np.random.randn(*lead, dx)
    P0 = np.random.randn(*lead, dx, 5 * dx)
    P0 = P0 @ np.swapaxes(P0, -1, -2)
    Fs = np.random.randn(T - 1, *lead, dx, dx)
    Qs = np.random.randn(T - 1, *lead, dx, 5 * dx)
    Qs = Qs @ np.swapaxes(Qs, -1, -2)
    bs = np.random.randn(T - 1, *lead, dx)
    Hs = np.random.randn(T, *lead, dy, dx)
    Rs = np.random.randn(T, *lead, dy, 5 * dy)
    Rs = Rs @ np.swapaxes(Rs, -1, -2)
    cs = np.random.randn(T, *lead, dy)
    ys = np.random.randn(T, *lead, dy)
    return ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs


@pytest.mark.parametrize("seed", [0, 1234])
@pytest.mark.parametrize("T", [5, 7])
@pytest.mark.parametrize("dx", [1, 2])
@pytest.mark.parametrize("dy", [1, 3])
@pytest.mark.parametrize("parallel", [False, True])
def test_filter_vs_reference_explicit_filter_with_nans(seed, T, dx, dy, parallel):
    # test_filtering.py:20-55 (the reference itself only runs parallel=False here; the oracle's parallel path is
    # held to the same explicit filter)
    common = ref_common()
    np.random.seed(seed)
    ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs = reference_style_model(T, dx, dy)
    ys[1, :] = np.nan
    ys[3, 0] = np.nan
    Hs[3, 0, :] = np.nan
    with np.errstate(all="ignore"):
        ms, Ps, ell = ok.filtering(ys, (m0, P0, Fs, Qs, bs, Hs, Rs, cs), parallel)
    ems, ePs, eell = common.explicit_kalman_filter(ys, m0, P0, Hs, Rs, cs, Fs, Qs, bs)
    npt.assert_allclose(ms, ems, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(Ps, ePs, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(ell, eell, rtol=1e-7)


@pytest.mark.parametrize("seed", [0, 1234])
@pytest.mark.parametrize("T", [3, 5])
@pytest.mark.parametrize("dx", [1, 2])
@pytest.mark.parametrize("dy", [1, 3])
@pytest.mark.parametrize("parallel", [True, False])
def test_batched_filter_vs_reference_explicit_filter(seed, T, dx, dy, parallel):
    # test_filtering.py:58-107
    common = ref_common()
    B = 3
    np.random.seed(seed)
    bys, bm0, bP0, bFs, bQs, bbs, bHs, bRs, bcs = reference_style_model(T, dx, dy, lead=(B,))
    bd = lambda a: np.stack([block_diag(*a[t]) for t in range(a.shape[0])])
    m0, P0 = bm0.reshape(B * dx), block_diag(*bP0)
    Fs, Qs, Hs, Rs = bd(bFs), bd(bQs), bd(bHs), bd(bRs)
    bs, cs, ys = bbs.reshape(T - 1, B * dx), bcs.reshape(T, B * dy), bys.reshape(T, B * dy)
    ems, ePs, eell = common.explicit_kalman_filter(ys, m0, P0, Hs, Rs, cs, Fs, Qs, bs)
    ms, Ps, ell = ok.filtering(ys, (m0, P0, Fs, Qs, bs, Hs, Rs, cs), parallel)
    npt.assert_allclose(ms, ems, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(Ps, ePs, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(ell, eell, rtol=1e-7)
    # the inner-batch form: the oracle keeps batch axes leading, [B, T, ...]
    sw = lambda a: np.swapaxes(a, 0, 1)
    bms, bPs, bell = ok.filtering(sw(bys), (bm0, bP0, sw(bFs), sw(bQs), sw(bbs), sw(bHs), sw(bRs), sw(bcs)), parallel)
    npt.assert_allclose(sw(bms).reshape(T, B * dx), ems, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(bd(sw(bPs)), ePs, rtol=1e-7, atol=1e-9)
    npt.assert_allclose(np.sum(bell), eell, rtol=1e-7)  # filtering.py:43-45


@pytest.mark.parametrize("seed", [42, 666])
@pytest.mark.parametrize("T,dx,dy", [(3, 1, 1), (5, 2, 3), (5, 2, 1)])
@pytest.mark.parametrize("parallel", [True, False])
def test_sampler_moments_vs_reference_explicit_smoother(seed, T, dx, dy, parallel):
    # test_sampling.py:23-68 (500 000 draws, atol = rtol = 1e-2); 200 000 vectorised draws here
    common = ref_common()
    np.random.seed(seed)
    ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs = reference_style_model(T, dx, dy)
    lg = (m0, P0, Fs, Qs, bs, Hs, Rs, cs)
    ms, Ps, _ = ok.filtering(ys, lg, False)
    n = 200_000
    eps = np.random.default_rng(seed).standard_normal((n, T, dx))
    xs = ok.sampling(eps, np.broadcast_to(ms, (n,) + ms.shape), np.broadcast_to(Ps, (n,) + Ps.shape),
                     tuple(np.broadcast_to(a, (n,) + a.shape) for a in lg), parallel)
    ems, ePs = common.explicit_kalman_smoothing(ms, Ps, Fs, Qs, bs)
    scale = np.sqrt(np.abs(ePs).max())
    npt.assert_allclose(xs.mean(0), ems, atol=2e-2 * scale, rtol=2e-2)
    for t in range(T):
        npt.assert_allclose(np.cov(xs[:, t].T).reshape(dx, dx), ePs[t], atol=2e-2 * scale ** 2, rtol=2e-2)


def _shim_env():
    env = dict(os.environ)
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "oracle", "refshim"), REF])
    return env


def test_reference_test_modules_pass_on_the_jax_stand_in():
    """The reference's own test_filtering.py (48 cases incl. the NaN pattern and the batched / parallel model), the
    exact batched-sampler test and the multinomial frequency test run green on oracle/refshim."""
    t = os.path.join(REF, "aux_samplers", "_primitives")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider",
           os.path.join(t, "test_kalman", "test_filtering.py"),
           os.path.join(t, "test_kalman", "test_sampling.py") + "::test_batched_model",
           os.path.join(t, "test_csmc", "test_resamplings.py") + "::test_multinomial_resampling"]
    r = subprocess.run(cmd, cwd="/tmp", env=_shim_env(), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_committed_reference_fixtures_regenerate_identically(tmp_path):
    """tests/golden/ref_*.npz are what make_reference_golden.py produces from the reference tree today."""
    script = os.path.join(ROOT, "tests", "golden", "make_reference_golden.py")
    src = open(script).read().replace('HERE = os.path.dirname(os.path.abspath(__file__))', 'HERE = %r' % str(tmp_path))
    src = src.replace("ROOT = os.path.dirname(os.path.dirname(HERE))", "ROOT = %r" % ROOT)
    gen = tmp_path / "gen.py"
    gen.write_text(src)
    r = subprocess.run([sys.executable, str(gen), "kalman", "misc"], cwd="/tmp", capture_output=True, text=True,
                       timeout=1500, env={**os.environ, "PYTHONDONTWRITEBYTECODE": "1"})
    assert r.returncode == 0, r.stderr[-3000:]
    for name in ("ref_kalman_primitives.npz", "ref_resampling_misc.npz"):
        new, old = np.load(tmp_path / name), np.load(os.path.join(ROOT, "tests", "golden", name))
        assert sorted(new.files) == sorted(old.files)
        for k in new.files:
            if k != "provenance":
                npt.assert_array_equal(new[k], old[k])

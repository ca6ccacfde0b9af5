"""The library is usable with NO PyTorch in the process: `aux_ssm_samplers_b200.raw` drives the C ABI with cuda-python
device memory (the boundary of SURVEY 8(b): plain pointers and sizes, nothing torch-typed).  Runs in a child process
where `import torch` is made to fail, filters and samples a time-varying LGSSM with dense H / R (the reference's test
shapes, test_filtering.py:20-55) and compares with the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
sys.modules["torch"] = None                      # any `import torch` now raises ImportError
import numpy as np
sys.path.insert(0, %r)
from aux_ssm_samplers_b200 import raw
assert "torch" not in [m for m, v in sys.modules.items() if v is not None], "torch was imported"
from oracle import kalman as ok
rng = np.random.default_rng(5)
for dtype, tol in ((np.float64, 1e-9), (np.float32, 2e-4)):
    T, dx, dy = 24, 3, 2
    spd = lambda n, k: (lambda a: a @ np.swapaxes(a, -1, -2) / (5 * n))(rng.standard_normal((k, n, 5 * n)))
    m0, P0 = rng.standard_normal(dx), spd(dx, 1)[0]
    Fs, Qs, bs = 0.7 * rng.standard_normal((T - 1, dx, dx)) / np.sqrt(dx), spd(dx, T - 1), rng.standard_normal((T - 1, dx))
    Hs, Rs, cs = rng.standard_normal((T, dy, dx)), spd(dy, T), rng.standard_normal((T, dy))
    ys, eps = rng.standard_normal((T, dy)), rng.standard_normal((T, dx))
    lg = [a.astype(dtype) for a in (m0, P0, Fs, Qs, bs, Hs, Rs, cs)]
    for parallel in (False, True):
        ms, Ps, ell = raw.filtering(ys.astype(dtype), lg, parallel)
        xs = raw.sampling(eps.astype(dtype), ms, Ps, lg, parallel)
        ms_o, Ps_o, ell_o = ok.filtering(ys, (m0, P0, Fs, Qs, bs, Hs, Rs, cs), parallel)
        xs_o = ok.sampling(eps, ms_o, Ps_o, (m0, P0, Fs, Qs, bs, Hs, Rs, cs), parallel)
        for got, want in ((ms, ms_o), (Ps, Ps_o), (np.asarray(ell), np.sum(ell_o)), (xs, xs_o)):
            np.testing.assert_allclose(got, want, rtol=tol * 50, atol=tol * 50 * max(np.abs(want).max(), 1.0))
print("NO_TORCH_OK")
'''


def test_c_abi_runs_without_torch_in_the_process():
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NO_TORCH_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, validates
descriptors before any launch, and the Python mirror refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from aux_ssm_samplers_b200 import _lib
    from aux_ssm_samplers_b200.build import build_library
    build_library()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    from aux_ssm_samplers_b200 import _lib
    header = open(os.path.join(ROOT, "include", "auxssm.h")).read()
    declared = set(re.findall(r"\b(auxssm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.auxssm_version() == 200


def test_descriptor_validation_without_gpu(lib):
    from aux_ssm_samplers_b200._lib import LgssmDesc
    d = LgssmDesc()
    d.T, d.dx, d.dy, d.n_chains, d.n_batch = 8, 65, 2, 1, 1
    assert lib.auxssm_kalman_filter_scratch_bytes(d, 4) == 0
    assert b"AUXSSM_MAX_DIM" in lib.auxssm_last_error()
    d.dx = 0
    rc = lib.auxssm_kalman_filter_f32(None, d, *([None] * 12), None, 0)
    assert rc == -1
    d.dx, d.parallel, d.time_chunks = 3, 1, 4
    assert lib.auxssm_kalman_filter_scratch_bytes(d, 8) > 0
    assert lib.auxssm_kalman_sample_scratch_bytes(d, 8) > 0
    d.parallel = 0
    assert lib.auxssm_kalman_sample_scratch_bytes(d, 8) >= 8 * (3 * 3 + 3) * 8  # (G_t, e_t) elements
    # null data pointers are rejected before any launch
    rc = lib.auxssm_kalman_filter_f64(None, d, *([None] * 12), None, 0)
    assert rc == -1 and b"null" in lib.auxssm_last_error()


def test_python_mirror_has_no_cpu_path():
    from aux_ssm_samplers_b200 import AuxssmError, kalman
    T, d = 4, 2
    lg = (torch.zeros(d), torch.eye(d), torch.eye(d).repeat(T - 1, 1, 1), torch.eye(d).repeat(T - 1, 1, 1),
          torch.zeros(T - 1, d), None, torch.eye(d).repeat(T, 1, 1), None)
    with pytest.raises(AuxssmError, match="CUDA tensors only"):
        kalman.filtering(torch.zeros(T, d), lg, False)


def test_reference_api_surface():
    import aux_ssm_samplers_b200 as pkg
    init, kernel = pkg.get_kernel(lambda x: None, lambda x, u, d: None, lambda x: None, True)
    st = init(torch.zeros(3, 2))
    assert st.updated is True and st.x.shape == (3, 2)
    assert pkg.LGSSM._fields == ("m0", "P0", "Fs", "Qs", "bs", "Hs", "Rs", "cs")


def test_package_import_is_lazy_and_the_abi_module_is_torch_free():
    """`import aux_ssm_samplers_b200._abi` / `.raw` must not import torch (checked in a child process)."""
    import subprocess
    import sys
    code = ("import sys; sys.modules['torch'] = None; sys.path.insert(0, %r); "
            "import aux_ssm_samplers_b200, aux_ssm_samplers_b200._abi as a, aux_ssm_samplers_b200.raw; "
            "print(len(a.exported_symbols()))" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and int(r.stdout.strip()) >= 60, r.stderr[-2000:]

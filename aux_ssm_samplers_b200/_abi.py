"""The C ABI of libauxssm.so as ctypes declarations (descriptor structs, enums, one signature per entry point of
include/auxssm.h).  No torch, no CUDA import: shared by `_lib` (torch plumbing: tensors, streams) and `raw`
(cuda-python plumbing: no PyTorch anywhere in the process)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libauxssm.so")

AUXSSM_MAX_DIM = 64
FLAG_H_IDENTITY = 1
FLAG_R_DIAGONAL = 2
FLAG_F_IS_DIAGONAL = 4
LOGPDF_PRIOR, LOGPDF_LOGLIK, LOGPDF_MINUS_ELL, LOGPDF_POSTERIOR = 1, 2, 4, 7


class AuxssmError(RuntimeError):
    pass


class Stride(ctypes.Structure):
    _fields_ = [("chain", ctypes.c_int64), ("batch", ctypes.c_int64), ("time", ctypes.c_int64)]


class LgssmDesc(ctypes.Structure):
    _fields_ = [
        ("T", ctypes.c_int32), ("dx", ctypes.c_int32), ("dy", ctypes.c_int32),
        ("n_chains", ctypes.c_int32), ("n_batch", ctypes.c_int32),
        ("parallel", ctypes.c_int32), ("time_chunks", ctypes.c_int32), ("flags", ctypes.c_uint32),
        ("ys", Stride), ("m0", Stride), ("P0", Stride), ("Fs", Stride), ("Qs", Stride), ("bs", Stride),
        ("Hs", Stride), ("Rs", Stride), ("cs", Stride), ("ms", Stride), ("Ps", Stride),
        ("eps", Stride), ("xs", Stride),
    ]


class CsmcDesc(ctypes.Structure):
    _fields_ = [
        ("T", ctypes.c_int32), ("N", ctypes.c_int32), ("d", ctypes.c_int32), ("n_chains", ctypes.c_int32),
        ("proposal", ctypes.c_int32), ("potential", ctypes.c_int32), ("resampling", ctypes.c_int32),
        ("backward", ctypes.c_int32), ("gradient", ctypes.c_int32),
        ("ys_chain_stride", ctypes.c_int64), ("gains_time_stride", ctypes.c_int64), ("sig", ctypes.c_double),
        ("rng_seed", ctypes.c_uint64), ("rng_counter", ctypes.c_uint64), ("rng_counter_ptr", ctypes.c_void_p),
        ("dyn_time_varying", ctypes.c_int32), ("pt_F", ctypes.c_void_p), ("pt_b", ctypes.c_void_p),
        ("pt_chol_Q", ctypes.c_void_p),
    ]


class SweepDesc(ctypes.Structure):
    _fields_ = [("lg", LgssmDesc), ("order", ctypes.c_int32), ("potential", ctypes.c_int32),
                ("delta_stride", ctypes.c_int32), ("sig", ctypes.c_double), ("ys_chain_stride", ctypes.c_int64)]


PROP_BOOTSTRAP, PROP_INDEPENDENT, PROP_GUIDED = 0, 1, 2
POT_FLAT, POT_GAUSSIAN, POT_SV = 0, 1, 2
RESAMPLE_MULTINOMIAL, RESAMPLE_KILLING, RESAMPLE_SYSTEMATIC = 0, 1, 2

_P = ctypes.c_void_p
_I32, _I64, _U32, _SZ = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_size_t
_U64, _F64 = ctypes.c_uint64, ctypes.c_double
_DESC = ctypes.POINTER(LgssmDesc)

# name -> (restype, argtypes); `{s}` expands to f32 / f64
_SIGNATURES = {
    "auxssm_last_error": (ctypes.c_char_p, []),
    "auxssm_version": (_I32, []),
    "auxssm_launch_count": (ctypes.c_int64, []),
    "auxssm_dispatch_log": (ctypes.c_char_p, [_I32]),
    "auxssm_kalman_filter_scratch_bytes": (_SZ, [_DESC, _I32]),
    "auxssm_kalman_filter_{s}": (_I32, [_P, _DESC] + [_P] * 12 + [_P, _SZ]),
    "auxssm_kalman_sample_scratch_bytes": (_SZ, [_DESC, _I32]),
    "auxssm_kalman_sample_{s}": (_I32, [_P, _DESC] + [_P] * 7 + [_P, _SZ]),
    "auxssm_lgssm_logpdf_scratch_bytes": (_SZ, [_DESC, _I32]),
    "auxssm_lgssm_logpdf_{s}": (_I32, [_P, _DESC, _U32] + [_P] * 12 + [_P, _SZ]),
    "auxssm_kalman_sweep_scratch_bytes": (_SZ, [ctypes.POINTER(SweepDesc), _I32]),
    "auxssm_kalman_sweep_{s}": (_I32, [_P, ctypes.POINTER(SweepDesc)] + [_P] * 14 + [_P, _SZ]),
    "auxssm_auxiliary_draw_{s}": (_I32, [_P, _I64, _I64, _P, _P, _P, _I32, _P]),
    "auxssm_kalman_mh_{s}": (_I32, [_P, _I64, _I64, _P, _I32] + [_P] * 10),
    "auxssm_sv_obs_factory_{s}": (_I32, [_P, _I32, _I64, _I64, _I32, _P, _I64, _P, _P, _P, _I32, _P, _P, _I32]),
    "auxssm_sv_log_potential_{s}": (_I32, [_P, _I64, _I64, _P, _I64, _P, _P, _I32]),
    "auxssm_csmc_scratch_bytes": (_SZ, [ctypes.POINTER(CsmcDesc), _I32]),
    "auxssm_csmc_sweep_{s}": (_I32, [_P, ctypes.POINTER(CsmcDesc)] + [_P] * 21 + [_P, _SZ]),
    "auxssm_pit_csmc_scratch_bytes": (_SZ, [ctypes.POINTER(CsmcDesc), _I32]),
    "auxssm_pit_csmc_sweep_{s}": (_I32, [_P, ctypes.POINTER(CsmcDesc)] + [_P] * 14 + [_P, _SZ]),
    "auxssm_guided_gains_{s}": (_I32, [_P, _I32, _I32, _P, _P, _P, _P, _P]),
    "auxssm_normalize_{s}": (_I32, [_P, _I64, _I32, _P, _P]),
    "auxssm_resample_multinomial_{s}": (_I32, [_P, _I64, _I32, _P, _P, _P]),
    "auxssm_resample_killing_{s}": (_I32, [_P, _I64, _I32, _P, _P, _P, _P]),
    "auxssm_categorical_{s}": (_I32, [_P, _I64, _I32, _P, _P, _P]),
    "auxssm_resample_systematic_{s}": (_I32, [_P, _I64, _I32, _I32, _P, _P, _P]),
    "auxssm_csmc_grad_log_target_{s}": (_I32, [_P, ctypes.POINTER(CsmcDesc)] + [_P] * 8),
    "auxssm_cholesky_{s}": (_I32, [_P, _I64, _I32, _P, _P]),
    "auxssm_delta_adaptation_{s}": (_I32, [_P, _I64, _P, _P, _I32, _F64, _F64, _F64, _F64, _P]),
    "auxssm_mcmc_stats_{s}": (_I32, [_P, _I32, _I64, _I32, _P, _P, _P, _F64, _P]),
    "auxssm_autocov_{s}": (_I32, [_P, _I32, _I64, _P, _P, _P]),
    "auxssm_sigma_points_{s}": (_I32, [_P, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    "auxssm_sigma_fit_{s}": (_I32, [_P, _I64, _I32, _I32, _I32] + [_P] * 6 + [_I64, _I64, _P, _P, _P]),
    "auxssm_philox_uniform_{s}": (_I32, [_P, _U64, _P, _I64, _P]),
    "auxssm_philox_normal_{s}": (_I32, [_P, _U64, _P, _I64, _P]),
    "auxssm_philox_advance": (_I32, [_P, _P, _U64]),
}


def exported_symbols():
    """Every symbol include/auxssm.h declares (used by the CPU-side ABI test)."""
    out = []
    for name in _SIGNATURES:
        if "{s}" in name:
            out += [name.format(s="f32"), name.format(s="f64")]
        else:
            out.append(name)
    return out


def load():
    """ctypes handle of libauxssm.so with restype / argtypes set on every entry point; raises AuxssmError when the
    library has not been built (there is no CPU fallback)."""
    if not os.path.exists(LIB_PATH):
        raise AuxssmError(
            "libauxssm.so not found at %s -- build it with `python -m aux_ssm_samplers_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        for sfx in (("f32", "f64") if "{s}" in name else ("",)):
            fn = getattr(handle, name.format(s=sfx))
            fn.restype = res
            fn.argtypes = args
    return handle

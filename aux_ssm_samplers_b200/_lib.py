"""ctypes binding of libauxssm.so (the C ABI declared in include/auxssm.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used only for device memory, streams and (elsewhere) torch.distributed.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libauxssm.so")

from ._abi import (AUXSSM_MAX_DIM, FLAG_H_IDENTITY, FLAG_R_DIAGONAL, FLAG_F_IS_DIAGONAL, LOGPDF_PRIOR, LOGPDF_LOGLIK,  # noqa: F401
                   LOGPDF_MINUS_ELL, LOGPDF_POSTERIOR, AuxssmError, Stride, LgssmDesc, CsmcDesc, SweepDesc, PROP_BOOTSTRAP,
                   PROP_INDEPENDENT, PROP_GUIDED, POT_FLAT, POT_GAUSSIAN, POT_SV, RESAMPLE_MULTINOMIAL, RESAMPLE_KILLING,
                   RESAMPLE_SYSTEMATIC, _SIGNATURES, exported_symbols, load)

_lib = None


def lib():
    """Load libauxssm.so (once).  Raises AuxssmError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    handle = load()
    _lib = handle
    return _lib


def suffix(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise AuxssmError("libauxssm supports float32 and float64 only, got %s" % dtype)


def check(rc, what):
    if rc != 0:
        msg = lib().auxssm_last_error().decode("utf-8", "replace")
        raise AuxssmError("%s failed (code %d): %s" % (what, rc, msg))


class Profiler:
    """Brackets every C-ABI call with CUDA events on the current stream (bench.py's per-kernel timing)."""

    def __init__(self):
        self.records = []  # (name, start_event, end_event)
        self._open = None

    def begin(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._open = (name, ev)

    def end(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.records.append((self._open[0], self._open[1], ev))
        self._open = None

    def summary(self):
        """name -> (calls, total_ms); call after torch.cuda.synchronize()."""
        out = {}
        for name, a, b in self.records:
            n, ms = out.get(name, (0, 0.0))
            out[name] = (n + 1, ms + a.elapsed_time(b))
        return out


_profiler = None


def set_profiler(p):
    global _profiler
    _profiler = p


def call(name, *args):
    """Invoke one C-ABI entry point; raises AuxssmError on a non-zero status."""
    fn = getattr(lib(), name)
    p = _profiler
    if p is not None:
        p.begin(name)
    rc = fn(*args)
    if p is not None:
        p.end(name)
    check(rc, name)


def launch_count():
    """Cumulative number of CUDA kernels libauxssm.so has launched in this process."""
    return int(lib().auxssm_launch_count())


def dispatch_log(reset=True):
    """Names of the kernels this host thread has launched since the last reset (which path served a call)."""
    return lib().auxssm_dispatch_log(1 if reset else 0).decode("utf-8", "replace").split(";")


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise AuxssmError("libauxssm operates on CUDA tensors only (got a %s tensor); there is no CPU path"
                              % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise AuxssmError("all tensors must live on the same CUDA device")
    return dev

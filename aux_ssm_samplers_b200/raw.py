"""libauxssm.so without PyTorch: the C ABI driven with cuda-python device memory and streams.

``north_star`` keeps the host side in the reference's own stack (JAX arrays handed to ``jax.ffi`` custom calls); PyTorch
is only the plumbing THIS repository's Python mirror happens to use.  This module shows that nothing in the library
depends on it: device buffers come from ``cuda.bindings.runtime`` (``cudaMalloc`` / ``cudaMemcpy``), arrays are NumPy,
and the calls are the plain ctypes entry points of ``_abi`` -- the route any array library with device pointers takes
(route B).  ``tests/test_no_torch_gpu.py`` runs it in a process where ``import torch`` is blocked.

``filtering`` / ``sampling`` below mirror ``_primitives/kalman/filtering.py:18-46`` and ``sampling.py:11-40`` for one
time-varying LGSSM in the reference's layout (``ys [T,dy]``, ``Fs [T-1,dx,dx]`` ...), fp32 or fp64.
"""
import ctypes

import numpy as np

from . import _abi
from ._abi import AuxssmError, LgssmDesc, Stride

_rt = None
_handle = None


def _runtime():
    global _rt
    if _rt is None:
        from cuda.bindings import runtime as rt  # cuda-python
        _rt = rt
    return _rt


def _ok(res):
    err, *out = res
    if int(err) != 0:
        raise AuxssmError("CUDA runtime error %s" % (err,))
    return out[0] if len(out) == 1 else (out if out else None)


def lib():
    global _handle
    if _handle is None:
        _handle = _abi.load()
    return _handle


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise AuxssmError("%s failed (code %d): %s" % (name, rc, lib().auxssm_last_error().decode("utf-8", "replace")))


class DeviceBuffer:
    """A ``cudaMalloc`` allocation; ``from_numpy`` / ``to_numpy`` copy synchronously."""

    def __init__(self, nbytes):
        self.nbytes = int(max(nbytes, 1))
        self.ptr = int(_ok(_runtime().cudaMalloc(self.nbytes)))

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        buf = cls(a.nbytes)
        if a.nbytes:
            _ok(_runtime().cudaMemcpy(buf.ptr, a.ctypes.data, a.nbytes, _runtime().cudaMemcpyKind.cudaMemcpyHostToDevice))
        return buf

    def to_numpy(self, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        if out.nbytes:
            _ok(_runtime().cudaMemcpy(out.ctypes.data, self.ptr, out.nbytes, _runtime().cudaMemcpyKind.cudaMemcpyDeviceToHost))
        return out

    @property
    def c(self):
        return ctypes.c_void_p(self.ptr)

    def free(self):
        if self.ptr:
            _runtime().cudaFree(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class Stream:
    def __init__(self):
        self.handle = _ok(_runtime().cudaStreamCreate())

    @property
    def c(self):
        return ctypes.c_void_p(int(self.handle))

    def synchronize(self):
        _ok(_runtime().cudaStreamSynchronize(self.handle))


def _sfx(dtype):
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]


def _desc(T, dx, dy, parallel):
    d = LgssmDesc()
    d.T, d.dx, d.dy, d.n_chains, d.n_batch, d.parallel = T, dx, dy, 1, 1, int(bool(parallel))
    tm = lambda n: Stride(0, 0, n)  # noqa: E731  (one chain: only the time stride matters)
    d.ys, d.Fs, d.Qs, d.bs = tm(dy), tm(dx * dx), tm(dx * dx), tm(dx)
    d.Hs, d.Rs, d.cs = tm(dy * dx), tm(dy * dy), tm(dy)
    d.ms, d.Ps, d.eps, d.xs = tm(dx), tm(dx * dx), tm(dx), tm(dx)
    return d


def filtering(ys, lgssm, parallel=False, stream=None):
    """``filtering(ys, lgssm, parallel)`` of filtering.py:18-46 -> ``(ms [T,dx], Ps [T,dx,dx], ell)`` as NumPy arrays."""
    m0, P0, Fs, Qs, bs, Hs, Rs, cs = [np.ascontiguousarray(a) for a in lgssm]
    dtype = ys.dtype
    T, dy = ys.shape
    dx = m0.shape[0]
    st = stream or Stream()
    d = _desc(T, dx, dy, parallel)
    bufs = [DeviceBuffer.from_numpy(np.asarray(a, dtype=dtype)) for a in (ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs)]
    ms, Ps, ell = DeviceBuffer(T * dx * dtype.itemsize), DeviceBuffer(T * dx * dx * dtype.itemsize), DeviceBuffer(dtype.itemsize)
    nb = lib().auxssm_kalman_filter_scratch_bytes(ctypes.byref(d), dtype.itemsize)
    scratch = DeviceBuffer(nb)
    call("auxssm_kalman_filter_" + _sfx(dtype), st.c, ctypes.byref(d), *[b.c for b in bufs], ms.c, Ps.c, ell.c, scratch.c, nb)
    st.synchronize()
    return ms.to_numpy((T, dx), dtype), Ps.to_numpy((T, dx, dx), dtype), ell.to_numpy((1,), dtype)[0]


def sampling(eps, ms, Ps, lgssm, parallel=False, stream=None):
    """``sampling(key, ms, Ps, lgssm, parallel)`` of sampling.py:11-40 with the N(0, I) draws ``eps [T,dx]`` of :128
    supplied by the caller -> ``xs [T,dx]``."""
    m0, P0, Fs, Qs, bs, *_ = [np.ascontiguousarray(a) for a in lgssm]
    dtype = ms.dtype
    T, dx = ms.shape
    st = stream or Stream()
    d = _desc(T, dx, dx, parallel)
    bufs = [DeviceBuffer.from_numpy(np.asarray(a, dtype=dtype)) for a in (ms, Ps, Fs, Qs, bs, eps)]
    xs = DeviceBuffer(T * dx * dtype.itemsize)
    nb = lib().auxssm_kalman_sample_scratch_bytes(ctypes.byref(d), dtype.itemsize)
    scratch = DeviceBuffer(nb)
    call("auxssm_kalman_sample_" + _sfx(dtype), st.c, ctypes.byref(d), *[b.c for b in bufs], xs.c, scratch.c, nb)
    st.synchronize()
    return xs.to_numpy((T, dx), dtype)

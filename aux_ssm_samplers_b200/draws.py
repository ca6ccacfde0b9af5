"""Stand-ins for JAX PRNG keys.

The reference draws inside its kernels with ``jax.random.*(key, ...)``.  Here the key -> draw step is
the caller's: a *key* is any object with ``normal(shape, dtype, device)`` and
``uniform(shape, dtype, device)``.  ``ExternalDraws`` replays caller-supplied arrays in call order
(parity mode: identical draws are fed to the oracle and to the CUDA path); ``PhiloxDraws`` generates them
on the device with libauxssm's counter-based Philox4x32-10 kernels (reproducible from (seed, call counter),
restated in ``oracle/philox.py``, and safe inside a captured CUDA graph because the counter lives in device
memory); ``TorchDraws`` uses a seeded ``torch.Generator`` (plumbing).  None is bit-compatible with JAX's threefry.
"""
import ctypes

import torch

from . import _lib


class ExternalDraws:
    def __init__(self, normals=(), uniforms=()):
        self._normals = list(normals)
        self._uniforms = list(uniforms)

    def _pop(self, queue, kind, shape, dtype, device):
        if not queue:
            raise RuntimeError("ExternalDraws: no %s draws left for shape %s" % (kind, tuple(shape)))
        t = torch.as_tensor(queue.pop(0)).to(device=device, dtype=dtype)
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError("ExternalDraws: %s draw has shape %s, expected %s"
                               % (kind, tuple(t.shape), tuple(shape)))
        return t

    def normal(self, shape, dtype, device):
        return self._pop(self._normals, "normal", shape, dtype, device)

    def uniform(self, shape, dtype, device):
        return self._pop(self._uniforms, "uniform", shape, dtype, device)


class TorchDraws:
    def __init__(self, seed, device="cuda"):
        self._gen = torch.Generator(device=device)
        self._gen.manual_seed(int(seed))

    def normal(self, shape, dtype, device):
        return torch.randn(shape, dtype=dtype, device=device, generator=self._gen)

    def uniform(self, shape, dtype, device):
        return torch.rand(shape, dtype=dtype, device=device, generator=self._gen)


class PhiloxDraws:
    """Draws generated on the device by ``auxssm_philox_{normal,uniform}_*``.

    Draw call number k (0, 1, 2, ... in call order, whatever its kind) fills its array with the Philox4x32-10
    stream keyed by ``seed`` at counter words (block index, k).  ``oracle.philox`` reproduces the same arrays
    on the CPU, which keeps parity tests possible without shipping draws over PCIe.
    """

    def __init__(self, seed, device="cuda", in_kernel=True):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.device = torch.device(device)
        self.counter = torch.zeros((1,), dtype=torch.int64, device=self.device)
        # the fused cSMC sweeps can generate their particle draws themselves (no [C,T,N,d] array in HBM); they then
        # consume the same call numbers, so both settings give bit-identical chains
        self.in_kernel = bool(in_kernel)

    def advance(self, n_calls):
        """Skip ``n_calls`` draw calls (used after a sweep that generated its draws in the kernel)."""
        _lib.call("auxssm_philox_advance", _lib.stream_ptr(self.device), _lib.ptr(self.counter),
                  ctypes.c_uint64(int(n_calls)))

    def _fill(self, kind, shape, dtype, device):
        out = torch.empty(tuple(shape), dtype=dtype, device=device)
        if out.numel() == 0:
            return out
        _lib.require_cuda(out, self.counter)
        _lib.call("auxssm_philox_%s_%s" % (kind, _lib.suffix(dtype)), _lib.stream_ptr(out.device),
                  ctypes.c_uint64(self.seed), _lib.ptr(self.counter), out.numel(), _lib.ptr(out))
        return out

    def normal(self, shape, dtype, device):
        return self._fill("normal", shape, dtype, device)

    def uniform(self, shape, dtype, device):
        return self._fill("uniform", shape, dtype, device)

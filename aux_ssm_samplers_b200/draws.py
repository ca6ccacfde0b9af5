"""Stand-ins for JAX PRNG keys.

The reference draws inside its kernels with ``jax.random.*(key, ...)``.  Here the key -> draw step is
the caller's: a *key* is any object with ``normal(shape, dtype, device)`` and
``uniform(shape, dtype, device)``.  ``ExternalDraws`` replays caller-supplied arrays in call order
(parity mode: identical draws are fed to the oracle and to the CUDA path); ``TorchDraws`` generates
them on the device with a seeded ``torch.Generator`` (plumbing, not bit-compatible with JAX's threefry).
"""
import torch


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

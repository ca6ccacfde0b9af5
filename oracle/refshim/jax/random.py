"""`jax.random` stand-in: recorded, reproducible draws (TEST INFRASTRUCTURE; ../README.md).

A key is a split path.  Every draw is a pure function of (path, kind, shape) -- so a key used twice gives the
same numbers, like in JAX -- and is appended to LOG as (kind, path, array).  The bit streams are NOT JAX's
threefry streams; fixtures store the recorded draws next to the outputs instead."""
import numpy as _np

from ._core import wrap

LOG = []


class Key:
    def __init__(self, path):
        self.path = tuple(int(p) for p in path)

    def __repr__(self):
        return "Key%s" % (self.path,)


def PRNGKey(seed):
    return Key((int(seed),))


key = PRNGKey


def split(k, num=2):
    out = _np.empty((num,), dtype=object)
    for i in range(num):
        out[i] = Key(k.path + (i,))
    return out


def _gen(k, kind):
    return _np.random.default_rng(_np.random.SeedSequence(entropy=list(k.path) + [{"n": 11, "u": 13}[kind]]))


def normal(k, shape=(), dtype=float):
    v = _gen(k, "n").standard_normal(tuple(shape)).astype(dtype)
    LOG.append(("normal", k.path, v))
    return wrap(v)


def uniform(k, shape=(), dtype=float, minval=0.0, maxval=1.0):
    v = _gen(k, "u").uniform(size=tuple(shape)).astype(dtype)
    LOG.append(("uniform", k.path, v))
    return wrap(minval + (maxval - minval) * v)


def bernoulli(k, p=0.5, shape=None):
    p = _np.asarray(p)
    return wrap(_np.asarray(uniform(k, p.shape if shape is None else shape)) < p)


def choice(k, a, shape=(), replace=True, p=None, axis=0):
    """Upstream JAX (jax/_src/random.py, replace=True with p): inverse-CDF search on cumsum(p)."""
    arr = None
    if not _np.isscalar(a) and _np.ndim(a) > 0:
        arr = _np.asarray(a)
        n = arr.shape[axis]
    else:
        n = int(a)
    if p is None:
        raise NotImplementedError
    cdf = _np.cumsum(_np.asarray(p))
    r = cdf[-1] * (1.0 - _np.asarray(uniform(k, shape)))
    idx = _np.searchsorted(cdf, r)
    idx = _np.minimum(idx, n - 1)  # XLA clamps out-of-range gathers
    return wrap(idx if arr is None else _np.take(arr, idx, axis=axis))

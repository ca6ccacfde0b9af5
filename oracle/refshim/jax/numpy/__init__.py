"""`jax.numpy` on NumPy (TEST INFRASTRUCTURE; ../../README.md).  Unknown names fall through to NumPy with their
array results re-wrapped as JArray; the functions the differentiated code paths use are Dual-aware."""
import numpy as _np

from .._core import (Dual, JArray, wrap, asarray, unary, linear, any_dual, stack as _stack, primal, as_dual,
                     vectorize, broadcast_to as _bto)
from . import linalg, fft  # noqa: F401

ndarray = JArray
pi, inf, nan, e, newaxis = _np.pi, _np.inf, _np.nan, _np.e, None
float32, float64, int32, int64, bool_ = _np.float32, _np.float64, _np.int32, _np.int64, _np.bool_


def _unwrapped(fn):
    def call(*args, **kwargs):
        return wrap(fn(*args, **kwargs))
    call.__name__ = getattr(fn, "__name__", "fn")
    return call


def __getattr__(name):
    attr = getattr(_np, name)
    if callable(attr) and not isinstance(attr, type):
        return _unwrapped(attr)
    return attr


array = asarray


def zeros_like(x, dtype=None, shape=None):
    return wrap(_np.zeros_like(primal(x), dtype=dtype, shape=shape))


def ones_like(x, dtype=None, shape=None):
    return wrap(_np.ones_like(primal(x), dtype=dtype, shape=shape))


def exp(x):
    return unary(x, _np.exp, exp, exp)


def log(x):
    return unary(x, _np.log, log, lambda p: 1.0 / p)


def sqrt(x):
    return unary(x, _np.sqrt, sqrt, lambda p: 0.5 / sqrt(p))


def abs(x):  # noqa: A001
    return unary(x, _np.abs, abs, lambda p: _np.sign(primal(p)))


def sum(x, axis=None, **kw):  # noqa: A001
    return linear(x, lambda a: _np.sum(a, axis=axis, **kw))


def nansum(x, axis=None, **kw):
    if isinstance(x, Dual):
        keep = ~_np.isnan(primal(x))
        return sum(where(keep, x, 0.0), axis=axis, **kw)
    return wrap(_np.nansum(x, axis=axis, **kw))


def reshape(x, shape, **kw):
    return linear(x, lambda a: _np.reshape(a, shape))


def ravel(x):
    return linear(x, _np.ravel)


def diag(x, k=0):
    return linear(x, lambda a: _np.diag(a, k))


def transpose(x, axes=None):
    return linear(x, lambda a: _np.transpose(a, axes))


def broadcast_to(x, shape):
    return _bto(x, shape)


def dot(a, b):
    if any_dual(a, b):
        return a @ b if isinstance(a, Dual) else b.__rmatmul__(a)
    return wrap(_np.dot(a, b))


def matmul(a, b):
    return dot(a, b) if any_dual(a, b) else wrap(_np.matmul(a, b))


def where(cond, x=None, y=None):
    if x is None:
        return wrap(_np.where(primal(cond)))
    cond = primal(cond)
    if any_dual(x, y):
        tag = max(v.tag for v in (x, y) if isinstance(v, Dual))
        xd = x if (isinstance(x, Dual) and x.tag == tag) else Dual(x, 0.0 * primal(x), tag)
        yd = y if (isinstance(y, Dual) and y.tag == tag) else Dual(y, 0.0 * primal(y), tag)
        return Dual(where(cond, xd.p, yd.p), where(cond, xd.t, yd.t), tag)
    return wrap(_np.where(cond, x, y))


def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    if isinstance(x, Dual):
        ok = _np.isfinite(primal(x))
        return Dual(nan_to_num(x.p, nan=nan, posinf=posinf, neginf=neginf), where(ok, x.t, 0.0), x.tag)
    if isinstance(nan, (_np.ndarray, Dual)):  # jnp accepts an array for `nan` (filtering.py: nan=y_hat)
        return wrap(_np.where(_np.isnan(x), primal(nan), _np.nan_to_num(x, nan=0.0, posinf=posinf, neginf=neginf)))
    with _np.errstate(over="ignore"):
        return wrap(_np.nan_to_num(x, copy=True, nan=nan, posinf=posinf, neginf=neginf))


def clip(x, a_min=None, a_max=None, **kw):
    a_min = kw.get("min", a_min)
    a_max = kw.get("max", a_max)
    if isinstance(x, Dual):
        p = primal(x)
        inside = _np.ones(_np.shape(p), bool)
        if a_min is not None:
            inside &= p >= a_min
        if a_max is not None:
            inside &= p <= a_max
        return Dual(clip(x.p, a_min, a_max), where(inside, x.t, 0.0), x.tag)
    return wrap(_np.clip(x, a_min, a_max))


def maximum(a, b):
    if any_dual(a, b):
        return where(primal(a) >= primal(b), a, b)
    return wrap(_np.maximum(a, b))


def minimum(a, b):
    if any_dual(a, b):
        return where(primal(a) <= primal(b), a, b)
    return wrap(_np.minimum(a, b))


def isfinite(x):
    return wrap(_np.isfinite(primal(x)))


def isnan(x):
    return wrap(_np.isnan(primal(x)))


def ndim(x):
    return len(x.shape) if isinstance(x, Dual) else _np.ndim(x)


def shape(x):
    return x.shape if isinstance(x, Dual) else _np.shape(x)


def stack(arrays, axis=0):
    return _stack(list(arrays), axis)


def concatenate(arrays, axis=0):
    arrays = list(arrays)
    if any_dual(*arrays):
        tag = max(a.tag for a in arrays if isinstance(a, Dual))
        ds = [a if (isinstance(a, Dual) and a.tag == tag) else Dual(a, 0.0 * primal(a), tag) for a in arrays]
        return Dual(concatenate([d.p for d in ds], axis), concatenate([d.t for d in ds], axis), tag)
    return wrap(_np.concatenate([_np.asarray(a) for a in arrays], axis=axis))


def insert(arr, obj, values, axis=None):
    if any_dual(arr, values):
        d = arr if isinstance(arr, Dual) else as_dual(arr, like=values)
        v = values if isinstance(values, Dual) else as_dual(values, like=d)
        return Dual(insert(d.p, obj, v.p, axis), insert(d.t, obj, v.t, axis), d.tag)
    return wrap(_np.insert(_np.asarray(arr), obj, _np.asarray(values), axis=axis))


def append(arr, values, axis=None):
    return wrap(_np.append(_np.asarray(arr), _np.asarray(values), axis=axis))


def take(a, indices, axis=None, **kw):
    return linear(a, lambda v: _np.take(v, _np.asarray(indices), axis=axis))


def einsum(subscripts, *operands, **kw):
    if any_dual(*operands):
        if len(operands) != 2:
            raise NotImplementedError
        a, b = operands
        tag = max(v.tag for v in operands if isinstance(v, Dual))
        ad = a if (isinstance(a, Dual) and a.tag == tag) else Dual(a, 0.0 * primal(a), tag)
        bd = b if (isinstance(b, Dual) and b.tag == tag) else Dual(b, 0.0 * primal(b), tag)
        return Dual(einsum(subscripts, ad.p, bd.p),
                    einsum(subscripts, ad.p, bd.t) + einsum(subscripts, ad.t, bd.p), tag)
    return wrap(_np.einsum(subscripts, *[_np.asarray(o) for o in operands]))


def flatnonzero(a, size=None, fill_value=0):
    idx = _np.flatnonzero(a)
    if size is not None:
        out = _np.full((size,), fill_value, dtype=idx.dtype)
        out[:min(size, idx.size)] = idx[:size]
        idx = out
    return wrap(idx)


def searchsorted(a, v, side="left", **kw):
    return wrap(_np.searchsorted(_np.asarray(a), _np.asarray(v), side=side))


def finfo(dtype):
    return _np.finfo(dtype)


def iinfo(dtype):
    return _np.iinfo(dtype)

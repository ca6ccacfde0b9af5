"""`jax.lax` control flow as Python loops over pytrees (TEST INFRASTRUCTURE; ../README.md)."""
import numpy as _np

from ._core import tree_map, tree_leaves, tree_stack, _slice_leaf, wrap, primal, Dual, linear
from . import numpy as jnp


def cond(pred, true_fun, false_fun, *operands, operand=None):
    if operand is not None:
        operands = (operand,)
    return true_fun(*operands) if bool(primal(pred)) else false_fun(*operands)


def select(pred, on_true, on_false):
    return jnp.where(pred, on_true, on_false)


def scan(f, init, xs, length=None, reverse=False, unroll=1):
    if length is None:
        length = tree_leaves(xs)[0].shape[0]
    order = range(length - 1, -1, -1) if reverse else range(length)
    carry, ys = init, []
    for i in order:
        carry, y = f(carry, tree_map(lambda l: _slice_leaf(l, i), xs))
        ys.append(y)
    if reverse:
        ys = ys[::-1]
    return carry, (tree_stack(ys) if ys and ys[0] is not None else None)


def fori_loop(lower, upper, body, init):
    v = init
    for i in range(int(lower), int(upper)):
        v = body(i, v)
    return v


def _sl(tree, s):
    return tree_map(lambda l: linear(l, lambda a: a[s]), tree)


def _interleave(a, b):
    """a[0], b[0], a[1], b[1], ... (len(a) == len(b) or len(b) + 1)"""
    def f(x, y):
        x, y = _np.asarray(x), _np.asarray(y)
        out = _np.empty((x.shape[0] + y.shape[0],) + x.shape[1:], dtype=x.dtype)
        out[0::2], out[1::2] = x, y
        return wrap(out)
    return tree_map(f, a, b)


def associative_scan(fn, elems, reverse=False, axis=0):
    """Work-efficient recursive odd/even prefix scan (the schedule jax.lax.associative_scan uses)."""
    if axis != 0:
        raise NotImplementedError
    if reverse:
        flip = lambda t: tree_map(lambda l: wrap(_np.asarray(l)[::-1]), t)
        return flip(associative_scan(fn, flip(elems)))  # jax flips the inputs and outputs, fn's argument order is unchanged
    n = tree_leaves(elems)[0].shape[0]

    def _scan(e, n):
        if n < 2:
            return e
        reduced = fn(_sl(e, slice(0, n - 1, 2)), _sl(e, slice(1, n, 2)))
        odd = _scan(reduced, n // 2)
        if n == 2:
            return _interleave(_sl(e, slice(0, 1)), odd)
        if n % 2 == 0:
            even = fn(_sl(odd, slice(0, -1)), _sl(e, slice(2, n, 2)))
        else:
            even = fn(odd, _sl(e, slice(2, n, 2)))
        even = tree_map(lambda a, b: jnp.concatenate([a, b]), _sl(e, slice(0, 1)), even)
        return _interleave(even, odd)

    return _scan(elems, n)

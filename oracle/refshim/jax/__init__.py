"""NumPy-backed stand-in for the slice of `jax` the reference imports (TEST INFRASTRUCTURE; ../README.md)."""
from ._core import jit, vmap, grad, value_and_grad, tree_map, JArray as Array  # noqa: F401
from . import numpy, lax, random, scipy, tree_util, dtypes  # noqa: F401

__version__ = "0.0-refshim"


def default_backend():
    return "cpu"


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def disable_jit(disable=True):
    import contextlib
    return contextlib.nullcontext()


def hessian(fun, argnums=0):
    return jacfwd(grad(fun, argnums), argnums)


def jacfwd(fun, argnums=0):
    """Jacobian by forward-mode duals: one pass per input coordinate; output shape = out.shape + in.shape."""
    import numpy as np
    from . import _core as c

    def jfun(*args):
        x = args[argnums]
        shape = np.shape(c.primal(x))
        tag = next(c._level)
        cols = []
        for idx in np.ndindex(*shape):
            e = np.zeros(shape)
            e[idx] = 1.0
            a = list(args)
            a[argnums] = c.Dual(x, c.wrap(e) if shape else 1.0, tag)
            out = fun(*a)
            cols.append(c.tree_map(lambda o: c._tangent_wrt(o, tag) + 0.0 * c.primal(o), out))
        st = c.tree_stack(cols)  # leading axis = input coordinates
        return c.tree_map(lambda l: c.linear(l, lambda v: np.moveaxis(v, 0, -1).reshape(v.shape[1:] + shape)), st)

    return jfun


jacobian = jacrev = jacfwd

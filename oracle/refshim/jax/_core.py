"""Core of the NumPy-backed `jax` stand-in (TEST INFRASTRUCTURE; see ../README.md).

Arrays (`JArray`), forward-mode dual numbers (`Dual`), pytrees, `vmap`, `grad`, `vectorize`.
"""
import dataclasses as _dc
import itertools
import re

import numpy as np


# ------------------------------------------------------------------------------------------------
# arrays with the .at[...] protocol
# ------------------------------------------------------------------------------------------------

class _AtIndexer:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtRef(self._arr, idx)


class _AtRef:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def _idx_np(self):
        return _unwrap_index(self._idx)

    def set(self, value):
        if isinstance(value, Dual) or isinstance(self._arr, Dual):
            a = as_dual(self._arr)
            v = as_dual(value, like=a)
            return Dual(_AtRef(wrap(a.p), self._idx).set(v.p), _AtRef(wrap(a.t), self._idx).set(v.t), a.tag)
        out = np.array(self._arr, copy=True)
        try:
            out[self._idx_np()] = value
        except IndexError:
            # XLA scatter semantics: out-of-bounds updates are DROPPED (single integer-array index per axis)
            idx = self._idx_np()
            idx = idx if isinstance(idx, tuple) else (idx,)
            arrs = np.broadcast_arrays(*[np.asarray(i) for i in idx])
            ok = np.ones(arrs[0].shape, bool)
            for ax, a in enumerate(arrs):
                ok &= (a >= -out.shape[ax]) & (a < out.shape[ax])
            vals = np.broadcast_to(np.asarray(value), arrs[0].shape)
            out[tuple(a[ok] for a in arrs)] = vals[ok]
        return wrap(out)

    def add(self, value):
        out = np.array(self._arr, copy=True)
        np.add.at(out, self._idx_np(), value)
        return wrap(out)


def _unwrap_index(idx):
    if isinstance(idx, tuple):
        return tuple(np.asarray(i) if isinstance(i, JArray) else i for i in idx)
    return np.asarray(idx) if isinstance(idx, JArray) else idx


class JArray(np.ndarray):
    """ndarray with jax's functional-update protocol."""

    @property
    def at(self):
        return _AtIndexer(self)

    def __hash__(self):  # jax arrays are hashable by identity in some contexts (static args)
        return id(self)

    def __iter__(self):
        for i in range(self.shape[0]):
            yield np.ndarray.__getitem__(self, i)

    # jax arrays are immutable: `a -= b` rebinds `a` to a new array and never mutates the caller's buffer
    # (math/utils.py:38 `log_weights -= logsumexp(log_weights)` relies on it)
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def __getitem__(self, idx):
        try:
            return super().__getitem__(idx)
        except IndexError:
            if not any(isinstance(i, np.ndarray) for i in (idx if isinstance(idx, tuple) else (idx,))):
                raise
            # XLA gather semantics: out-of-bounds integer indices are CLAMPED
            tup = idx if isinstance(idx, tuple) else (idx,)
            fixed = []
            for ax, i in enumerate(tup):
                if isinstance(i, np.ndarray) and i.dtype.kind in "iu":
                    fixed.append(np.clip(np.asarray(i), -self.shape[ax], self.shape[ax] - 1))
                else:
                    fixed.append(i)
            return super().__getitem__(tuple(fixed))


def wrap(x):
    """ndarray / scalar -> JArray (no copy); tuples / lists element-wise; everything else unchanged."""
    if isinstance(x, JArray) or isinstance(x, Dual):
        return x
    if isinstance(x, np.ndarray):
        return x.view(JArray) if x.dtype != object else x
    if isinstance(x, (np.floating, np.integer, np.bool_)):
        return np.asarray(x).view(JArray)
    if isinstance(x, tuple) and not hasattr(x, "_fields"):
        return tuple(wrap(v) for v in x)
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


def asarray(x, dtype=None):
    if isinstance(x, Dual):
        return x
    return wrap(np.asarray(x, dtype=dtype))


# ------------------------------------------------------------------------------------------------
# forward-mode dual numbers (nestable through level tags)
# ------------------------------------------------------------------------------------------------

_level = itertools.count(1)


class Dual:
    """value + eps * tangent, eps^2 = 0.  `p` and `t` are arrays, scalars or lower-level Duals."""
    __array_ufunc__ = None  # ndarray (op) Dual defers to Dual.__r<op>__
    __array_priority__ = 1000

    def __init__(self, p, t, tag):
        self.p, self.t, self.tag = p, t, tag

    # structure
    @property
    def shape(self):
        return np.shape(self.p) if not isinstance(self.p, Dual) else self.p.shape

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return self.p.dtype if hasattr(self.p, "dtype") else np.dtype(float)

    @property
    def T(self):
        return Dual(_T(self.p), _T(self.t), self.tag)

    @property
    def at(self):
        return _AtIndexer(self)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        return Dual(_index(self.p, idx), _index(self.t, idx), self.tag)

    def __iter__(self):
        for i in range(self.shape[0]):
            yield self[i]

    def reshape(self, *shape):
        return Dual(self.p.reshape(*shape), self.t.reshape(*shape), self.tag)

    # arithmetic
    def _coerce(self, other):
        """Returns (p, t) of `other` as seen from this level (a constant has a zero tangent)."""
        if isinstance(other, Dual) and other.tag == self.tag:
            return other.p, other.t
        if isinstance(other, Dual) and other.tag > self.tag:
            return None  # the other operand is the higher level: let it drive
        return other, 0.0

    def __add__(self, o):
        c = self._coerce(o)
        if c is None:
            return o.__radd__(self)
        return Dual(self.p + c[0], self.t + c[1], self.tag)

    __radd__ = __add__

    def __sub__(self, o):
        c = self._coerce(o)
        if c is None:
            return o.__rsub__(self)
        return Dual(self.p - c[0], self.t - c[1], self.tag)

    def __rsub__(self, o):
        c = self._coerce(o)
        return Dual(c[0] - self.p, c[1] - self.t, self.tag)

    def __neg__(self):
        return Dual(-self.p, -self.t, self.tag)

    def __mul__(self, o):
        c = self._coerce(o)
        if c is None:
            return o.__rmul__(self)
        return Dual(self.p * c[0], self.p * c[1] + self.t * c[0], self.tag)

    __rmul__ = __mul__

    def __truediv__(self, o):
        c = self._coerce(o)
        if c is None:
            return o.__rtruediv__(self)
        q = self.p / c[0]
        return Dual(q, (self.t - q * c[1]) / c[0], self.tag)

    def __rtruediv__(self, o):
        c = self._coerce(o)
        q = c[0] / self.p
        return Dual(q, (c[1] - q * self.t) / self.p, self.tag)

    def __pow__(self, k):
        if isinstance(k, Dual):
            raise NotImplementedError("Dual ** Dual")
        return Dual(self.p ** k, k * self.p ** (k - 1) * self.t, self.tag)

    def __matmul__(self, o):
        c = self._coerce(o)
        if c is None:
            return o.__rmatmul__(self)
        t = self.t @ c[0]
        if not (np.isscalar(c[1]) and c[1] == 0.0):
            t = t + self.p @ c[1]
        return Dual(self.p @ c[0], t, self.tag)

    def __rmatmul__(self, o):
        c = self._coerce(o)
        t = c[0] @ self.t
        if not (np.isscalar(c[1]) and c[1] == 0.0):
            t = t + c[1] @ self.p
        return Dual(c[0] @ self.p, t, self.tag)

    # comparisons act on the primal
    def __lt__(self, o):
        return primal(self) < primal(o)

    def __le__(self, o):
        return primal(self) <= primal(o)

    def __gt__(self, o):
        return primal(self) > primal(o)

    def __ge__(self, o):
        return primal(self) >= primal(o)

    def __repr__(self):
        return "Dual(%r, %r, tag=%d)" % (self.p, self.t, self.tag)


def _T(x):
    return x.T if hasattr(x, "T") else x


def _index(x, idx):
    if isinstance(x, Dual):
        return x[idx]
    if np.ndim(x) == 0:
        return x
    return wrap(np.asarray(x)[_unwrap_index(idx)])


def primal(x):
    while isinstance(x, Dual):
        x = x.p
    return x


def as_dual(x, like=None, tag=None):
    if isinstance(x, Dual):
        return x
    tag = like.tag if like is not None else tag
    return Dual(x, np.zeros(np.shape(primal(x))) if np.ndim(primal(x)) else 0.0, tag)


def unary(x, f_np, f_lifted, df_lifted):
    """Lift an element-wise function: f(p), f'(p) * t.  `f_lifted` / `df_lifted` accept Duals (nesting)."""
    if isinstance(x, Dual):
        return Dual(f_lifted(x.p), df_lifted(x.p) * x.t, x.tag)
    return wrap(f_np(x))


def linear(x, f):
    """Lift a function that is linear in x (reshape, sum, indexing, ...)."""
    if isinstance(x, Dual):
        return Dual(linear(x.p, f), linear(x.t, f), x.tag)
    return wrap(f(np.asarray(x)))


def any_dual(*xs):
    return any(isinstance(x, Dual) for x in xs)


def stack(items, axis=0):
    items = list(items)
    if any_dual(*items):
        tag = max(x.tag for x in items if isinstance(x, Dual))
        ds = [x if (isinstance(x, Dual) and x.tag == tag) else
              Dual(x, _zeros_like(x), tag) for x in items]
        return Dual(stack([d.p for d in ds], axis), stack([d.t for d in ds], axis), tag)
    return wrap(np.stack([np.asarray(x) for x in items], axis=axis))


def _zeros_like(x):
    if isinstance(x, Dual):
        return Dual(_zeros_like(x.p), _zeros_like(x.t), x.tag)
    return np.zeros(np.shape(x)) if np.ndim(x) else 0.0


def broadcast_to(x, shape):
    return linear(x, lambda a: np.broadcast_to(a, shape))


# ------------------------------------------------------------------------------------------------
# pytrees
# ------------------------------------------------------------------------------------------------

def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def _is_node(x):
    from chex import is_dataclass_node
    return x is None or isinstance(x, (tuple, list, dict)) or is_dataclass_node(x)


def tree_map(f, tree, *rest, is_leaf=None):
    from chex import is_dataclass_node
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    if tree is None:
        return None
    if _is_namedtuple(tree):
        return type(tree)(*[tree_map(f, t, *[r[i] for r in rest], is_leaf=is_leaf) for i, t in enumerate(tree)])
    if isinstance(tree, tuple):
        return tuple(tree_map(f, t, *[r[i] for r in rest], is_leaf=is_leaf) for i, t in enumerate(tree))
    if isinstance(tree, list):
        return [tree_map(f, t, *[r[i] for r in rest], is_leaf=is_leaf) for i, t in enumerate(tree)]
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest], is_leaf=is_leaf) for k, v in tree.items()}
    if is_dataclass_node(tree):
        kw = {fl.name: tree_map(f, getattr(tree, fl.name), *[getattr(r, fl.name) for r in rest], is_leaf=is_leaf)
              for fl in _dc.fields(tree)}
        return type(tree)(**kw)
    return f(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(lambda x: out.append(x), tree)
    return out


def tree_stack(trees):
    first = trees[0]
    return tree_map(lambda *xs: stack(xs), first, *trees[1:])


def _slice_leaf(leaf, i, axis=0):
    if isinstance(leaf, Dual):
        return Dual(_slice_leaf(leaf.p, i, axis), _slice_leaf(leaf.t, i, axis), leaf.tag)
    if isinstance(leaf, np.ndarray):
        out = np.take(leaf, i, axis=axis) if leaf.dtype != object else leaf[i]
        return wrap(out)
    return leaf[i]


# ------------------------------------------------------------------------------------------------
# transformations
# ------------------------------------------------------------------------------------------------

def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0, **_kw):
    def vmapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            for leaf in tree_leaves(a):
                n = leaf.shape[ax]
                break
            if n is not None:
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree_map(lambda l: _slice_leaf(l, i, ax), a) for a, ax in zip(args, axes)]
            outs.append(fun(*sl))
        if n == 0:
            # empty batch (dc_map.py:104 passes an empty "unchanged" set through when T is a power of two): JAX
            # returns outputs with a zero-length mapped axis; evaluate `fun` once on a placeholder element for shapes
            def fake(l, ax):
                shp = tuple(d for j, d in enumerate(np.shape(l)) if j != ax)
                return wrap(np.zeros(shp, dtype=np.asarray(l).dtype))
            sl = [a if ax is None else tree_map(lambda l, ax=ax: fake(l, ax), a) for a, ax in zip(args, axes)]
            proto = fun(*sl)
            return tree_map(lambda r: wrap(np.zeros((0,) + np.shape(r), dtype=np.asarray(r).dtype)), proto)
        res = tree_stack(outs)
        if out_axes != 0:
            res = tree_map(lambda l: linear(l, lambda a: np.moveaxis(a, 0, out_axes)), res)
        return res

    return vmapped


def _tangent_wrt(out, tag):
    if isinstance(out, Dual):
        if out.tag == tag:
            return out.t
        # a higher level wraps this one (cannot happen with properly nested grads)
        return Dual(_tangent_wrt(out.p, tag), _tangent_wrt(out.t, tag), out.tag)
    return 0.0


def grad(fun, argnums=0, has_aux=False):
    def gfun(*args):
        x = args[argnums]
        shape = primal(x).shape if hasattr(primal(x), "shape") else ()
        tag = next(_level)
        pieces = []
        for idx in np.ndindex(*shape):
            e = np.zeros(shape)
            e[idx] = 1.0
            xd = Dual(x, wrap(e) if shape else 1.0, tag)
            a = list(args)
            a[argnums] = xd
            out = fun(*a)
            pieces.append(_tangent_wrt(out, tag))
        if not shape:
            return wrap(pieces[0]) if not isinstance(pieces[0], Dual) else pieces[0]
        if any_dual(*pieces):
            return stack(pieces).reshape(*shape)
        return wrap(np.array([float(p) for p in pieces]).reshape(shape))

    return gfun


def value_and_grad(fun, argnums=0):
    g = grad(fun, argnums)
    return lambda *args: (fun(*args), g(*args))


# ------------------------------------------------------------------------------------------------
# jnp.vectorize with a gufunc signature: explicit loop over the broadcast batch dimensions
# ------------------------------------------------------------------------------------------------

_ARG = r"\(([^)]*)\)"


def _parse(sig):
    ins, outs = sig.replace(" ", "").split("->")
    f = lambda s: [tuple(d for d in g.split(",") if d) for g in re.findall(_ARG, s)]
    return f(ins), f(outs)


def vectorize(pyfunc=None, *, signature=None, excluded=None):
    if pyfunc is None:
        return lambda f: vectorize(f, signature=signature, excluded=excluded)
    if signature is None:
        raise NotImplementedError("refshim.vectorize needs a signature")
    in_core, out_core = _parse(signature)

    excl = sorted(excluded or ())

    def wrapped(*all_args):
        fixed = {i: all_args[i] for i in excl if i < len(all_args)}
        args = [a for i, a in enumerate(all_args) if i not in fixed]

        def call(*vals):
            vals = list(vals)
            for i in sorted(fixed):
                vals.insert(i, fixed[i])
            return pyfunc(*vals)

        args = [a if isinstance(a, Dual) else asarray(a) for a in args]
        batch = [a.shape[:a.ndim - len(c)] for a, c in zip(args, in_core)]
        bshape = np.broadcast_shapes(*batch)
        if bshape == ():
            return wrap(call(*args))
        full = [broadcast_to(a, bshape + tuple(a.shape[a.ndim - len(c):])) for a, c in zip(args, in_core)]
        outs = []
        for idx in np.ndindex(*bshape):
            outs.append(call(*[_index(a, idx) for a in full]))
        single = not isinstance(outs[0], tuple)
        if single:
            outs = [(o,) for o in outs]
        res = []
        for k in range(len(outs[0])):
            st = stack([o[k] for o in outs])
            core = st.shape[1:]
            res.append(st.reshape(*(tuple(bshape) + tuple(core))))
        return res[0] if single else tuple(res)

    wrapped.__name__ = getattr(pyfunc, "__name__", "vectorized")
    wrapped.__doc__ = getattr(pyfunc, "__doc__", None)
    return wrapped

import numpy as _np
import scipy.linalg as _sl

from .._core import wrap, Dual, linear


def solve(a, b, assume_a="gen", lower=False, **kw):
    with _np.errstate(all="ignore"):
        return wrap(_sl.solve(_np.asarray(a), _np.asarray(b), assume_a=assume_a, lower=lower, check_finite=False))


def cho_solve(c_and_lower, b, **kw):
    c, lower = c_and_lower
    with _np.errstate(all="ignore"):
        return wrap(_sl.cho_solve((_np.asarray(c), lower), _np.asarray(b), check_finite=False))


def solve_triangular(a, b, trans=0, lower=False, unit_diagonal=False, **kw):
    f = lambda v: _tri(_np.asarray(a), v, trans, lower, unit_diagonal)
    return linear(b, f) if isinstance(b, Dual) else wrap(f(_np.asarray(b)))


def _tri(a, b, trans, lower, unit):
    with _np.errstate(all="ignore"):
        if not _np.all(_np.isfinite(a)):
            # substitution by hand so that the clipped (huge) entries of mvn.logpdf behave as in XLA (x / inf = 0)
            if trans not in (0, "N") or not lower or b.ndim != 1:
                raise NotImplementedError
            x = _np.zeros_like(b, dtype=float)
            for i in range(a.shape[0]):
                x[i] = (b[i] - _np.dot(_np.nan_to_num(a[i, :i] * x[:i], nan=0.0), _np.ones(i))) / (1.0 if unit else a[i, i])
            return x
        return _sl.solve_triangular(a, b, trans=trans, lower=lower, unit_diagonal=unit, check_finite=False)


def eigh(a, **kw):
    w, v = _np.linalg.eigh(_np.asarray(a))
    return wrap(w), wrap(v)


def cholesky(a, lower=False):
    from ..numpy.linalg import cholesky as _c
    L = _c(a)
    return L if lower else wrap(_np.asarray(L).T)

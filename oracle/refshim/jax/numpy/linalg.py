import numpy as _np

from .._core import wrap, Dual


def cholesky(a):
    """XLA semantics: an all-NaN factor where the factorisation fails (NumPy raises instead)."""
    a = _np.asarray(a)
    if a.ndim > 2:
        return wrap(_np.stack([_np.asarray(cholesky(x)) for x in a]))
    try:
        with _np.errstate(all="ignore"):
            if not _np.all(_np.isfinite(a)):
                # LAPACK with non-finite input: run the textbook column algorithm so that inf / nan propagate
                return wrap(_chol_ref(a))
            return wrap(_np.linalg.cholesky(a))
    except _np.linalg.LinAlgError:
        return wrap(_np.full(a.shape, _np.nan))


def _chol_ref(a):
    n = a.shape[0]
    L = _np.zeros_like(a, dtype=float)
    for j in range(n):
        s = a[j, j] - _np.dot(L[j, :j], L[j, :j])
        if not (s > 0):
            return _np.full(a.shape, _np.nan)
        L[j, j] = _np.sqrt(s)
        for i in range(j + 1, n):
            L[i, j] = (a[i, j] - _np.dot(L[i, :j], L[j, :j])) / L[j, j]
    return L


def __getattr__(name):
    fn = getattr(_np.linalg, name)

    def call(*args, **kwargs):
        kwargs.pop("hermitian", None) if name == "svd" else None
        return wrap(fn(*args, **kwargs))
    return call

import numpy as _np
import scipy.special as _sp

from .._core import wrap


def logsumexp(a, axis=None, b=None, keepdims=False, **kw):
    with _np.errstate(all="ignore"):
        return wrap(_sp.logsumexp(_np.asarray(a), axis=axis, b=b, keepdims=keepdims))

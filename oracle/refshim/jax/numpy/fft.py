import numpy as _np

from .._core import wrap


def __getattr__(name):
    fn = getattr(_np.fft, name)
    return lambda *a, **k: wrap(fn(*a, **k))

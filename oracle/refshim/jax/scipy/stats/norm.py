import math

from ... import numpy as jnp


def logpdf(x, loc=0.0, scale=1.0):
    """Closed form, Dual-aware: -(x - loc)^2 / (2 scale^2) - log(scale) - log(2 pi) / 2."""
    z = (x - loc) / scale
    return -0.5 * (z * z) - jnp.log(scale + 0.0 * z) - 0.5 * math.log(2.0 * math.pi)


def pdf(x, loc=0.0, scale=1.0):
    return jnp.exp(logpdf(x, loc, scale))

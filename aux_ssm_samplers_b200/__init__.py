"""B200-native (sm_100a) hot paths of aux-ssm-samplers behind the reference's Python API.

``kalman``                 auxiliary Kalman sampler (filter / backward sampler / log-densities / MH sweep)
``stochastic_volatility``  SV model adapters used by the headline benchmark
``csmc``                   conditional SMC particle Gibbs (fused sweep + unfused primitives)
``diagnostics``            cross-chain split-R-hat / ESS over all-gathered chain statistics
``linearisation``          extended / cubature / Gauss--Hermite linearisation (``aux_samplers.extended`` etc.)
``draws``                  key stand-ins (externally supplied or device-generated draws)
"""
from . import csmc, diagnostics, draws, kalman, linearisation, stochastic_volatility  # noqa: F401
from ._lib import AuxssmError  # noqa: F401
from .kalman import LGSSM, get_kernel  # noqa: F401
from .csmc import get_generic_kernel, get_independent_kernel  # noqa: F401
from .linearisation import extended, gauss_hermite, cubature  # noqa: F401  (aux_samplers/__init__.py:2)

"""B200-native (sm_100a) hot paths of aux-ssm-samplers behind the reference's Python API.

``kalman``                 auxiliary Kalman sampler (filter / backward sampler / log-densities / MH sweep)
``stochastic_volatility``  SV model adapters used by the headline benchmark
``draws``                  key stand-ins (externally supplied or device-generated draws)
"""
from . import draws, kalman, stochastic_volatility  # noqa: F401
from ._lib import AuxssmError  # noqa: F401
from .kalman import LGSSM, get_kernel  # noqa: F401

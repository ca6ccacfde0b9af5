"""B200-native (sm_100a) hot paths of aux-ssm-samplers behind the reference's Python API.

``kalman``                 auxiliary Kalman sampler (filter / backward sampler / log-densities / MH sweep)
``latent_gaussian``        the same sweep fused into one launch for narrow latent-Gaussian models
``stochastic_volatility``  SV model adapters used by the headline benchmark
``csmc``                   conditional SMC particle Gibbs (fused sweep, parallel-in-time sweep, unfused primitives)
``linearisation``          extended / cubature / Gauss--Hermite linearisation (``aux_samplers.extended`` etc.)
``mcmc``                   step-size adaptation, running statistics, CUDA-graph outer loop
``diagnostics``            cross-chain split-R-hat / ESS over all-gathered chain statistics
``draws``                  key stand-ins (externally supplied or device-generated draws)
``raw``                    the C ABI with cuda-python device memory: no PyTorch in the process

Submodules are imported on first use (PEP 562), so ``import aux_ssm_samplers_b200.raw`` does not import torch.
"""
import importlib

_SUBMODULES = ("csmc", "diagnostics", "draws", "kalman", "latent_gaussian", "linearisation", "mcmc", "raw",
               "stochastic_volatility")
_NAMES = {  # public name -> defining submodule (aux_samplers/__init__.py exports the same constructors)
    "AuxssmError": "_abi", "LGSSM": "kalman", "get_kernel": "kalman", "get_generic_kernel": "csmc",
    "get_independent_kernel": "csmc", "extended": "linearisation", "gauss_hermite": "linearisation",
    "cubature": "linearisation", "delta_adaptation": "mcmc",
}


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    if name in _NAMES:
        return getattr(importlib.import_module("." + _NAMES[name], __name__), name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def __dir__():
    return sorted(list(globals()) + list(_SUBMODULES) + list(_NAMES))

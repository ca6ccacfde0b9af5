"""`jax.dtypes` stand-in (TEST INFRASTRUCTURE; ../README.md): keys are Python objects here, so the key dtype is
NumPy's object dtype (`_primitives/csmc/pit/dc_map.py:152` pads key arrays by edge replication)."""
import numpy as _np

prng_key = _np.object_

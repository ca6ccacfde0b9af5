"""CPU oracle for the two hot paths of aux-ssm-samplers (TEST INFRASTRUCTURE ONLY).

This package is a NumPy restatement of the reference's JAX algorithms with every
``jax.random.*`` call replaced by an explicit array argument.  It is *not* part of
the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``aux_ssm_samplers_b200``) never imports anything from here and has no CPU fallback.

Pinning status
--------------
The reference (pure Python/JAX) cannot be imported in this image: ``jax``, ``jaxlib``
and ``chex`` are absent and there is no network, and the reference ships no stored
golden vectors.  The oracle is therefore pinned against the pins the reference's own
tests use, restated in ``oracle/explicit.py``:

* ``explicit_kalman_filter`` / ``explicit_kalman_smoothing``
  (reference ``aux_samplers/_primitives/test_kalman/common.py:5-79``),
* the Chopin--Singh Alg. 4 conditional systematic resampler
  (reference ``aux_samplers/_primitives/test_csmc/test_resamplings.py:44-78``),
* ``scipy.stats.multivariate_normal.logpdf`` (doctest of
  ``aux_samplers/_primitives/math/mvn/base.py:37-46``).

Parts of the path that no reference test touches (posterior_logpdf, the MH step,
``normalize``, csmc/generic.py, csmc/independent.py, the backward-sampling index
sequence) and the pieces named by BASELINE.json but absent from the reference
(killing resampling, R-hat) are **parity unpinned**: they follow the reference's code
(or the literature) by reading, with no executable reference output to compare to.
"""

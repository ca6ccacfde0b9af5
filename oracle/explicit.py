"""Textbook pins used to validate the oracle (oracle only; test infrastructure).

Independent, deliberately naive formulations of what the reference's own tests
compare against:

* a masked-row Kalman filter and an RTS smoother -- the role played by
  ``explicit_kalman_filter`` / ``explicit_kalman_smoothing`` in reference
  ``aux_samplers/_primitives/test_kalman/common.py:5-79``;
* conditional systematic resampling, Algorithm 4 of Chopin & Singh (2015), the
  role of ``_sys_resampling`` in reference ``test_csmc/test_resamplings.py:44-78``.
"""
import numpy as np
from scipy.stats import multivariate_normal


def kalman_filter_masked(ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs):
    """Kalman filter that *drops* non-finite observation rows.

    Model: x_0 ~ N(m0, P0); x_t = F_{t-1} x_{t-1} + b_{t-1} + N(0, Q_{t-1});
    y_t = H_t x_t + c_t + N(0, R_t) for t = 0..T-1.  Returns (ms, Ps, ell).
    """
    T = ys.shape[0]
    dx = m0.shape[0]
    ms = np.empty((T, dx))
    Ps = np.empty((T, dx, dx))
    m, P, ell = np.asarray(m0, float), np.asarray(P0, float), 0.0
    for t in range(T):
        if t > 0:
            F = Fs[t - 1]
            m = F @ m + bs[t - 1]
            P = F @ P @ F.T + Qs[t - 1]
        keep = np.isfinite(ys[t])
        if keep.any():
            H = Hs[t][keep]
            R = Rs[t][np.ix_(keep, keep)]
            resid = ys[t][keep] - (H @ m + cs[t][keep])
            S = H @ P @ H.T + R
            ell += multivariate_normal.logpdf(resid, np.zeros(resid.shape[0]), S)
            gain = np.linalg.solve(S, H @ P).T
            m = m + gain @ resid
            P = P - gain @ H @ P
        ms[t], Ps[t] = m, P
    return ms, Ps, ell


def rts_smoother(ms, Ps, Fs, Qs, bs):
    """Rauch--Tung--Striebel smoother marginals from filtering marginals."""
    T = ms.shape[0]
    sm, sP = ms.copy(), Ps.copy()
    for t in range(T - 2, -1, -1):
        F = Fs[t]
        Ppred = F @ Ps[t] @ F.T + Qs[t]
        gain = np.linalg.solve(Ppred, F @ Ps[t]).T
        sm[t] = ms[t] + gain @ (sm[t + 1] - (F @ ms[t] + bs[t]))
        sP[t] = Ps[t] + gain @ (sP[t + 1] - Ppred) @ gain.T
    return sm, sP


def conditional_systematic_alg4(W, u, v, w, M):
    """Chopin & Singh (2015), Algorithm 4, for one triple of uniforms (u, v, w).

    Note: as in the reference's test helper, ``np.nonzero`` returns a 1-tuple so
    ``n_zero`` is the tuple length (always 1) -- i.e. the cyclic shift of step (c)
    is never applied.  We restate the *mathematical* algorithm here (shift applied
    when more than one offspring of particle 0 exists) and expose the degenerate
    reading through ``apply_shift=False`` in the oracle, see ``oracle/csmc.py``.
    """
    W = np.asarray(W, float)
    nW1 = M * W[0]
    if nW1 <= 1.0:
        U = nW1 * u
    else:
        r1 = nW1 - np.floor(nW1)
        if v < r1 * np.ceil(nW1) / nW1:
            U = r1 * u
        else:
            U = r1 + (1.0 - r1) * u
    grid = (np.arange(M) + U) / M
    a_bar = np.searchsorted(np.cumsum(W), grid)
    zero_loc = np.flatnonzero(a_bar == 0)
    n_zero = zero_loc.shape[0]
    if n_zero <= 1:
        return np.clip(a_bar, 0, len(W) - 1)
    shift = zero_loc[int(np.floor(n_zero * w))]
    return np.clip(np.roll(a_bar, -shift), 0, len(W) - 1)

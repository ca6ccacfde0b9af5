"""NumPy restatement of the reference's Kalman hot path (oracle only; test infrastructure).

Follows, function by function:

* ``aux_samplers/_primitives/kalman/filtering.py`` (filtering, sequential and
  parallel-in-time via the 5-tuple associative operator),
* ``aux_samplers/_primitives/kalman/sampling.py`` (backward sampling, sequential and
  reverse associative scan),
* ``aux_samplers/_primitives/kalman/base.py:73-166`` (trajectory log-densities),
* ``aux_samplers/kalman/generic.py:53-106`` (the auxiliary Kalman MH sweep).

Every function is vectorised over arbitrary *leading* axes (chains / inner batch);
the time axis, when present, is the first axis after those leading axes is NOT assumed:
time-indexed arrays are laid out ``[..., T, ...]`` with the leading axes first (the
layout a ``jax.vmap`` over chains produces).  Random draws are explicit arguments.

Pinned by tests/test_oracle_kalman.py against ``oracle/explicit.py``.
"""
import numpy as np

from .linalg import chol, cho_solve, mT, mv, mvn_logpdf, norm_logpdf, solve_lower, sym, LOG_2PI

# --------------------------------------------------------------------------------------
# filtering.py
# --------------------------------------------------------------------------------------


def _mask_obs(y, H, c, R):
    """Missing-observation masking, reference ``filtering.py:91-101`` / ``:204-212``."""
    nan = ~np.isfinite(y)
    eye = np.eye(y.shape[-1], dtype=bool)
    R_ = np.where(nan[..., None, :] | nan[..., :, None], 0.0, R)
    R_ = np.where(eye & nan[..., :, None], np.inf, R_)
    H_ = np.where(nan[..., :, None], 0.0, H)
    c_ = np.where(nan, 0.0, c)
    return nan, H_, c_, R_


def sequential_update(y, m, P, H, c, R):
    """Reference ``filtering.py:84-130``.  Returns (m, P, ell_inc)."""
    y, m, P, H, c, R = _lead_bcast((y, 1), (m, 1), (P, 2), (H, 2), (c, 1), (R, 2))
    nan, H_, c_, R_ = _mask_obs(y, H, c, R)
    any_obs = np.any(~nan, axis=-1)
    dy = y.shape[-1]
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        y_hat = mv(H_, m) + c_
        y_ = np.where(np.isnan(y), y_hat, y)
        y_diff = y_ - y_hat
        y_diff = np.where(nan, 0.0, y_diff)  # +-inf observations are "missing" too (isfinite)
        S = R_ + H_ @ P @ mT(H_)
        if dy == 1:
            chol_S = np.sqrt(S)
            ell_inc = norm_logpdf(y_[..., 0], y_hat[..., 0], chol_S[..., 0, 0])
            G = (P @ mT(H_)) / S
        else:
            chol_S = chol(S)
            ell_inc = mvn_logpdf(y_, y_hat, chol_S)
            fmax = np.finfo(chol_S.dtype).max
            chol_S = np.where(np.isfinite(chol_S), chol_S, np.where(chol_S < 0, -fmax, fmax))
            G = mT(cho_solve(chol_S, H_ @ P))
        m_new = m + mv(G, y_diff)
        S0 = np.where(np.isfinite(S), S, 0.0)
        P_new = sym(P - G @ S0 @ mT(G))
        ell_inc = np.where(np.isnan(ell_inc), 0.0, ell_inc)
    m_out = np.where(any_obs[..., None], m_new, m)
    P_out = np.where(any_obs[..., None, None], P_new, P)
    ell = np.where(any_obs, ell_inc, 0.0)
    return m_out, P_out, ell


def _lead_bcast(*pairs):
    """Broadcast the leading (batch) axes of arrays whose trailing core ranks differ."""
    arrs = [np.asarray(a) for a, _ in pairs]
    lead = np.broadcast_shapes(*[a.shape[: a.ndim - r] for a, (_, r) in zip(arrs, pairs)])
    return [np.broadcast_to(a, lead + a.shape[a.ndim - r:]) for a, (_, r) in zip(arrs, pairs)]


def sequential_predict(m, P, F, b, Q):
    """Reference ``filtering.py:135-139``."""
    return mv(F, m) + b, sym(Q + F @ P @ mT(F))


def sequential_predict_update(m, P, F, b, Q, y, H, c, R):
    """Reference ``filtering.py:144-147``."""
    m, P = sequential_predict(m, P, F, b, Q)
    return sequential_update(y, m, P, H, c, R)


def filtering_init_one(F, Q, b, H, R, c, y, m, P):
    """Scan element (A, b, C, eta, J) for one time step, reference ``filtering.py:197-250``."""
    F, Q, b, H, R, c, y, m, P = _lead_bcast((F, 2), (Q, 2), (b, 1), (H, 2), (R, 2), (c, 1), (y, 1), (m, 1), (P, 2))
    dy = y.shape[-1]
    m_ = mv(F, m) + b
    P_ = F @ P @ mT(F) + Q
    nan, H_, c_, R_ = _mask_obs(y, H, c, R)
    any_obs = np.any(~nan, axis=-1)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        S = H_ @ P_ @ mT(H_) + R_
        if dy == 1:
            S_invH_T = mT(H_) / S[..., 0:1, 0:1]
        else:
            cS = chol(S)
            cS = np.where(np.isfinite(cS), cS, np.finfo(cS.dtype).max)
            S_invH_T = mT(cho_solve(cS, H_))
        K = P_ @ S_invH_T
        A = F - K @ H_ @ F
        y_diff_b = np.where(nan, 0.0, y - mv(H_, b) - c_)
        y_diff_m = np.where(nan, 0.0, y - mv(H_, m_) - c_)
        b_std = m_ + mv(K, y_diff_m)
        S0 = np.where(np.isfinite(S), S, 0.0)
        C = P_ - K @ S0 @ mT(K)
        temp = mT(F) @ S_invH_T
        eta = mv(temp, y_diff_b)
        J = temp @ H_ @ F
    sel1, sel2 = any_obs[..., None], any_obs[..., None, None]
    A = np.where(sel2, A, F)
    b_std = np.where(sel1, b_std, m_)
    C = sym(np.where(sel2, C, P_))
    eta = np.where(sel1, eta, 0.0)
    J = np.where(sel2, sym(J), 0.0)
    return A, b_std, C, eta, J


def filtering_op(e1, e2):
    """Associative combine of two elements, reference ``filtering.py:164-183``."""
    A1, b1, C1, eta1, J1 = e1
    A2, b2, C2, eta2, J2 = e2
    d = b1.shape[-1]
    eye = np.eye(d, dtype=b1.dtype)
    IpCJ = eye + C1 @ J2
    IpJC = eye + J2 @ C1
    if d == 1:
        X = A2 / IpCJ
        Y = A1 / IpJC
    else:
        X = mT(np.linalg.solve(mT(IpCJ), mT(A2)))
        Y = mT(np.linalg.solve(mT(IpJC), A1))
    A = X @ A1
    b = mv(X, b1 + mv(C1, eta2)) + b2
    C = X @ (C1 @ mT(A2)) + C2
    eta = mv(Y, eta2 - mv(J2, b1)) + eta1
    J = Y @ (J2 @ A1) + J1
    return A, b, sym(C), eta, sym(J)


def associative_scan(fn, elems, axis, reverse=False):
    """Odd/even recursive prefix scan with the call structure of ``jax.lax.associative_scan``.

    ``elems`` is a tuple of arrays; ``axis`` is the (non-negative) scan axis of each,
    counted from the front.  ``fn`` must be vectorised over that axis.
    """
    if reverse:
        flipped = tuple(np.flip(e, axis) for e in elems)
        out = associative_scan(lambda a, b: fn(a, b), flipped, axis)
        return tuple(np.flip(o, axis) for o in out)

    def sl(e, s):
        idx = [slice(None)] * e.ndim
        idx[axis] = s
        return e[tuple(idx)]

    n = elems[0].shape[axis]
    if n < 2:
        return elems
    reduced = fn(tuple(sl(e, slice(0, n - 1, 2)) for e in elems), tuple(sl(e, slice(1, n, 2)) for e in elems))
    odd = associative_scan(fn, reduced, axis)
    if n % 2 == 0:
        even = fn(tuple(sl(e, slice(0, -1)) for e in odd), tuple(sl(e, slice(2, None, 2)) for e in elems))
    else:
        even = fn(odd, tuple(sl(e, slice(2, None, 2)) for e in elems))
    even = tuple(np.concatenate([sl(e, slice(0, 1)), ev], axis=axis) for e, ev in zip(elems, even))
    out = []
    for ev, od in zip(even, odd):
        shape = list(ev.shape)
        shape[axis] = n
        o = np.empty(shape, dtype=ev.dtype)
        idx = [slice(None)] * o.ndim
        idx[axis] = slice(0, None, 2)
        o[tuple(idx)] = ev
        idx[axis] = slice(1, None, 2)
        o[tuple(idx)] = od
        out.append(o)
    return tuple(out)


def _tslice(a, core_rank, s):
    """Slice the time axis (the axis just before the ``core_rank`` trailing axes)."""
    idx = [slice(None)] * a.ndim
    idx[a.ndim - core_rank - 1] = s
    return a[tuple(idx)]


def filtering(ys, lgssm, parallel):
    """Reference ``filtering.py:18-63`` + ``:66-79``.

    Layout: ``ys [..., T, dy]``, ``m0 [..., dx]``, ``P0 [..., dx, dx]``,
    ``Fs, Qs [..., T-1, dx, dx]``, ``bs [..., T-1, dx]``, ``Hs [..., T, dy, dx]``,
    ``Rs [..., T, dy, dy]``, ``cs [..., T, dy]``.  Returns ``ms [..., T, dx]``,
    ``Ps [..., T, dx, dx]``, ``ell [...]``.  (The reference's inner-batch form
    ``[T, B, ...]`` is the same thing with the axes swapped; its ``ell`` is summed
    over B by the caller, ``filtering.py:43-45``.)
    """
    m0, P0, Fs, Qs, bs, Hs, Rs, cs = [np.asarray(a) for a in lgssm]
    ys = np.asarray(ys)
    T = ys.shape[-2]
    y0, H0, c0, R0 = _tslice(ys, 1, 0), _tslice(Hs, 2, 0), _tslice(cs, 1, 0), _tslice(Rs, 2, 0)
    m0f, P0f, ell0 = sequential_update(y0, m0, P0, H0, c0, R0)
    y1, H1, c1, R1 = (_tslice(ys, 1, slice(1, None)), _tslice(Hs, 2, slice(1, None)),
                      _tslice(cs, 1, slice(1, None)), _tslice(Rs, 2, slice(1, None)))
    lead = m0f.shape[:-1]
    dx = m0f.shape[-1]
    ms = np.empty(lead + (T, dx), dtype=m0f.dtype)
    Ps = np.empty(lead + (T, dx, dx), dtype=m0f.dtype)
    ms[..., 0, :], Ps[..., 0, :, :] = m0f, P0f
    if T == 1:
        return ms, Ps, ell0
    if not parallel:
        m, P, ell = m0f, P0f, ell0
        for t in range(T - 1):
            m, P, inc = sequential_predict_update(
                m, P, _tslice(Fs, 2, t), _tslice(bs, 1, t), _tslice(Qs, 2, t),
                _tslice(y1, 1, t), _tslice(H1, 2, t), _tslice(c1, 1, t), _tslice(R1, 2, t))
            ell = ell + inc
            ms[..., t + 1, :], Ps[..., t + 1, :, :] = m, P
        return ms, Ps, ell
    # parallel: filtering.py:49-63
    m_init = np.zeros(lead + (T - 1, dx), dtype=m0f.dtype)
    P_init = np.zeros(lead + (T - 1, dx, dx), dtype=m0f.dtype)
    m_init[..., 0, :], P_init[..., 0, :, :] = m0f, P0f
    elems = filtering_init_one(Fs, Qs, bs, H1, R1, c1, y1, m_init, P_init)
    axis = len(lead)
    _, bsc, Csc, _, _ = associative_scan(filtering_op, elems, axis)
    ms[..., 1:, :], Ps[..., 1:, :, :] = bsc, Csc
    *_, incs = sequential_predict_update(ms[..., :-1, :], Ps[..., :-1, :, :], Fs, bs, Qs, y1, H1, c1, R1)
    return ms, Ps, ell0 + np.nansum(incs, axis=-1)


# --------------------------------------------------------------------------------------
# sampling.py
# --------------------------------------------------------------------------------------


def mean_and_chol(F, Q, b, m, P):
    """Reference ``sampling.py:61-105``.  Returns (inc_m, L, gain)."""
    d = m.shape[-1]
    S = sym(F @ P @ mT(F) + Q)
    with np.errstate(invalid="ignore", divide="ignore"):
        if d == 1:
            gain = P * F / S
        else:
            gain = P @ mT(np.linalg.solve(S, F))
        inc_Sig = sym(P - gain @ S @ mT(gain))
        inc_m = m - mv(gain, mv(F, m) + b)
        L = np.sqrt(inc_Sig) if d == 1 else chol(inc_Sig)
    L = np.nan_to_num(L)
    return inc_m, L, gain


def sample_last_step(m, P, eps):
    """Reference ``sampling.py:116-124``."""
    with np.errstate(invalid="ignore"):
        L = np.sqrt(P) if P.shape[-1] == 1 else chol(P)
    L = np.nan_to_num(L)
    return np.zeros_like(P), m + mv(L, eps)


def sampling_op(e1, e2):
    """Reference ``sampling.py:52-55``."""
    G1, e1v = e1
    G2, e2v = e2
    return G2 @ G1, mv(G2, e1v) + e2v


def sampling(eps, ms, Ps, lgssm, parallel):
    """Reference ``sampling.py:11-40``; ``eps [..., T, dx]`` replaces ``normal(key, ms.shape)`` (``:128``)."""
    Fs, Qs, bs = np.asarray(lgssm[2]), np.asarray(lgssm[3]), np.asarray(lgssm[4])
    T = ms.shape[-2]
    lead = ms.shape[:-2]
    dx = ms.shape[-1]
    gains = np.empty(lead + (T, dx, dx), dtype=ms.dtype)
    incs = np.empty(lead + (T, dx), dtype=ms.dtype)
    if T > 1:
        inc_m, L, g = mean_and_chol(Fs, Qs, bs, ms[..., :-1, :], Ps[..., :-1, :, :])
        gains[..., :-1, :, :] = g
        incs[..., :-1, :] = inc_m + mv(L, eps[..., :-1, :])
    gains[..., -1, :, :], incs[..., -1, :] = sample_last_step(ms[..., -1, :], Ps[..., -1, :, :], eps[..., -1, :])
    if parallel:
        _, samples = associative_scan(sampling_op, (gains, incs), len(lead), reverse=True)
        return samples
    xs = np.empty_like(incs)
    xs[..., -1, :] = incs[..., -1, :]
    for t in range(T - 2, -1, -1):
        xs[..., t, :] = mv(gains[..., t, :, :], xs[..., t + 1, :]) + incs[..., t, :]
    return xs


# --------------------------------------------------------------------------------------
# base.py
# --------------------------------------------------------------------------------------


def prior_logpdf(xs, lgssm):
    """Reference ``base.py:100-134``."""
    m0, P0, Fs, Qs, bs = [np.asarray(a) for a in lgssm[:5]]
    pred = mv(Fs, xs[..., :-1, :]) + bs
    with np.errstate(invalid="ignore", divide="ignore"):
        if m0.shape[-1] == 1:
            out = np.nan_to_num(norm_logpdf(xs[..., 0, 0], m0[..., 0], np.sqrt(P0)[..., 0, 0]), nan=0.0)
            trans = norm_logpdf(xs[..., 1:, 0], pred[..., 0], np.sqrt(Qs)[..., 0, 0])
        else:
            out = np.nan_to_num(mvn_logpdf(xs[..., 0, :], m0, chol(P0)), nan=0.0)
            trans = mvn_logpdf(xs[..., 1:, :], pred, chol(Qs))
    return out + np.nansum(trans, axis=-1)


def log_likelihood(ys, xs, lgssm):
    """Reference ``base.py:138-166``."""
    Hs, Rs, cs = [np.asarray(a) for a in lgssm[5:8]]
    pred = mv(Hs, xs) + cs
    with np.errstate(invalid="ignore", divide="ignore"):
        if cs.shape[-1] == 1:
            out = norm_logpdf(ys[..., 0], pred[..., 0], np.sqrt(Rs)[..., 0, 0])
        else:
            out = mvn_logpdf(ys, pred, chol(Rs))
    return np.nansum(out, axis=-1)


def posterior_logpdf(ys, xs, ell, lgssm):
    """Reference ``base.py:73-96``."""
    return log_likelihood(ys, xs, lgssm) - ell + prior_logpdf(xs, lgssm)


# --------------------------------------------------------------------------------------
# kalman/generic.py
# --------------------------------------------------------------------------------------


def get_alpha(log_lgssm_prop, log_lgssm_rev, log_target_prop, log_target_rev, delta, u, x, x_prop):
    """Reference ``kalman/generic.py:98-106``."""
    log_alpha = log_target_prop - log_target_rev + log_lgssm_rev - log_lgssm_prop
    sd = np.sqrt(delta)
    corr = np.sum(((x_prop - u) / sd) ** 2 - ((x - u) / sd) ** 2, axis=(-2, -1))
    return np.exp(np.minimum(0.0, log_alpha - corr))


def kalman_sweep(x, delta, normals_aux, normals_eps, uniform_accept,
                 dynamics_factory, observations_factory, log_likelihood_fn, parallel):
    """One auxiliary-Kalman MH step, reference ``kalman/generic.py:53-90``.

    ``normals_aux`` replaces ``normal(auxiliary_key, x.shape)`` (``:61``), ``normals_eps``
    the draw inside ``sampling`` (``sampling.py:128``), ``uniform_accept`` the bernoulli
    draw (``:73``; ``bernoulli(key, p)`` is ``uniform(key) < p``).  Returns
    (x_new, accepted, aux) with ``aux`` a dict of intermediates for parity checks.
    """
    u = x + np.sqrt(0.5 * delta) * normals_aux

    def do_one(x_lin, x_prop=None):
        m0, P0, Fs, Qs, bs = dynamics_factory(x_lin)
        ys, Hs, Rs, cs = observations_factory(x_lin, u, delta)
        lg = (m0, P0, Fs, Qs, bs, Hs, Rs, cs)
        ms, Ps, ell = filtering(ys, lg, parallel)
        if x_prop is None:
            x_prop = sampling(normals_eps, ms, Ps, lg, parallel)
        return posterior_logpdf(ys, x_prop, ell, lg), log_likelihood_fn(x_prop), x_prop, (ms, Ps, ell)

    lp_prop, lt_prop, x_prop, fwd = do_one(x)
    lp_rev, lt_rev, _, rev = do_one(x_prop, x)
    alpha = get_alpha(lp_prop, lp_rev, lt_prop, lt_rev, delta, u, x, x_prop)
    accept = uniform_accept < alpha
    x_new = np.where(np.asarray(accept)[..., None, None], x_prop, x)
    aux = dict(u=u, x_prop=x_prop, alpha=alpha, log_lgssm_prop=lp_prop, log_lgssm_rev=lp_rev,
               log_target_prop=lt_prop, log_target_rev=lt_rev, ell_fwd=fwd[2], ell_rev=rev[2])
    return x_new, accept, aux

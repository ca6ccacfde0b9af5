"""NumPy restatement of the reference's conditional-SMC hot path (oracle only; test infrastructure).

Follows:

* ``aux_samplers/_primitives/csmc/csmc.py`` (forward conditional particle filter ``_csmc:69-107``,
  ancestor tracing ``_backward_scanning_pass:110-124``, backward sampling
  ``_backward_sampling_pass:127-149``, ``get_kernel/kernel:16-66``),
* ``aux_samplers/_primitives/csmc/resamplings.py`` (conditional multinomial ``:14-37`` and
  systematic ``:40-86``),
* ``aux_samplers/_primitives/math/utils.py:24-39`` (``normalize``),
* ``aux_samplers/csmc/generic.py:56-72`` and ``aux_samplers/csmc/independent.py:60-73,144-268``
  (auxiliary particle Gibbs wrappers),
* model families from ``_primitives/test_csmc/common.py:11-94`` and
  ``examples/stochastic_volatility/{auxiliary_csmc.py:17-48, auxiliary_guided_csmc.py:29-167}``.

Every ``jax.random`` draw is an explicit argument:
``normals [T, N, d]`` (``M0.sample`` at t=0, ``Mt.sample`` for t>=1: csmc.py:74,91),
``u_resample [T-1, N]`` (resamplings.py:35; killing also needs ``u_kill [T-1, N]``),
``u_backward [T]`` (csmc.py:111 or :131,138).

``jax.random.choice(key, M, p=w, shape)`` is restated from upstream JAX (not verifiable here) as
``cdf = cumsum(p); r = cdf[-1] * (1 - U); idx = searchsorted(cdf, r, side="left")``.

CANONICAL SUMMATION ORDER.  Ancestor indices depend on the floating-point order in which the CDF
is accumulated, and XLA's ``cumsum`` order is unspecified.  Oracle and CUDA kernels therefore share
one fixed order (``canonical_cdf`` / ``canonical_sum`` below): an inclusive Hillis--Steele scan
inside each group of 32 consecutive particles (offsets 1, 2, 4, 8, 16), the same scan over the
(zero-padded) group totals for the group offsets, one add of the group offset; sums use the
xor-butterfly (16, 8, 4, 2, 1) inside groups of 32 and the same butterfly over the group totals.  ``mismatches_vs_plain_cumsum`` reports how often a plain
``np.cumsum`` CDF would have picked a different index.

Parity status: multinomial index-0 pin / frequencies and systematic == Chopin--Singh Alg. 4 are
pinned by the reference's tests (tests/test_oracle_csmc.py); killing resampling is NOT in the
reference (spec from the literature, see ``killing``): parity unpinned.  The kernels of
csmc/generic.py, csmc/independent.py and the backward-sampling index sequence are unpinned too.
"""
import numpy as np

from .linalg import LOG_2PI, chol, cho_solve, mv, mvn_logpdf, norm_logpdf, solve_lower

# --------------------------------------------------------------------------------------------
# canonical reductions (shared with the CUDA kernels)
# --------------------------------------------------------------------------------------------

_LANES = np.arange(32)


def _pad32(a, fill):
    n = a.shape[-1]
    m = (n + 31) // 32 * 32
    if m == n:
        return a
    pad = np.full(a.shape[:-1] + (m - n,), fill, dtype=a.dtype)
    return np.concatenate([a, pad], axis=-1)


def canonical_sum(a):
    """Sum over the last axis in the kernels' order (dtype preserved): xor butterfly inside groups of 32,
    then the same butterfly over the (zero-padded) group totals, recursively."""
    a = np.asarray(a)
    while True:
        v = _pad32(a, 0).reshape(a.shape[:-1] + (-1, 32)).copy()
        for off in (16, 8, 4, 2, 1):
            v = v + v[..., _LANES ^ off]
        tot = v[..., 0]
        if tot.shape[-1] == 1:
            return tot[..., 0]
        a = tot


def canonical_max(a):
    return np.max(a, axis=-1)


def _scan32(v):
    """Hillis--Steele inclusive scan over the last axis of a [..., 32] array."""
    v = v.copy()
    for off in (1, 2, 4, 8, 16):
        shifted = np.zeros_like(v)
        shifted[..., off:] = v[..., :-off]
        v = np.where(_LANES >= off, v + shifted, v)
    return v


def canonical_cdf(w):
    """Inclusive prefix sum over the last axis in the kernels' order (dtype preserved): Hillis--Steele scan
    inside groups of 32, the canonical CDF of the (zero-padded) group totals -- the same rule, recursively -- gives
    the group offsets, one add of the offset (group 0 is not offset).  Up to 1024 entries that is two levels (what
    the particle kernels do with one thread per particle); the N^2 stitching weights of the parallel-in-time cSMC
    need a third."""
    w = np.asarray(w)
    n = w.shape[-1]
    v = _scan32(_pad32(w, 0).reshape(w.shape[:-1] + (-1, 32)))
    ng = v.shape[-2]
    out = v.copy()
    if ng > 1:
        tot = canonical_cdf(v[..., 31])                 # inclusive canonical scan of the group totals
        out[..., 1:, :] = v[..., 1:, :] + tot[..., :-1, None]
    return out.reshape(w.shape[:-1] + (-1,))[..., :n]


def normalize(log_w):
    """``normalize`` (math/utils.py:24-39) with the canonical sum: exp(lw - logsumexp(lw))."""
    log_w = np.asarray(log_w)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        amax = canonical_max(log_w)
        amax = np.where(np.isfinite(amax), amax, 0).astype(log_w.dtype)
        s = canonical_sum(np.exp(log_w - amax[..., None]))
        lse = (np.log(s) + amax).astype(log_w.dtype)
        return np.exp(log_w - lse[..., None])


def _search_left(cdf, r):
    """First index j with cdf[j] >= r (searchsorted side='left'), batched over leading axes."""
    return np.sum(cdf[..., None, :] < r[..., :, None], axis=-1).astype(np.int64)


def categorical(w, u):
    """``jax.random.choice(key, N, p=w, shape=())`` with uniform ``u``; batched over leading axes."""
    cdf = canonical_cdf(w)
    r = (cdf[..., -1] * (1 - u)).astype(cdf.dtype)
    return np.minimum(_search_left(cdf, r[..., None])[..., 0], w.shape[-1] - 1)


# --------------------------------------------------------------------------------------------
# resamplings.py
# --------------------------------------------------------------------------------------------


def multinomial(w, u):
    """Conditional multinomial resampling (resamplings.py:14-37): N draws, index 0 pinned."""
    cdf = canonical_cdf(w)
    r = (cdf[..., -1:] * (1 - u)).astype(cdf.dtype)
    idx = np.minimum(_search_left(cdf, r), w.shape[-1] - 1)
    idx[..., 0] = 0
    return idx


def killing(w, u_kill, u):
    """Conditional killing resampling -- NOT in the reference (BASELINE.json names it).

    Spec (Del Moral's killing / "residual-rejection" scheme made conditional the way the reference
    pins multinomial, resamplings.py:36): particle i survives (A_i = i) when
    ``u_kill_i * max_j w_j < w_i``; otherwise A_i is a multinomial draw from w; A_0 = 0 always.
    Parity unpinned.
    """
    wmax = np.max(w, axis=-1, keepdims=True)
    keep = u_kill * wmax < w
    idx = np.where(keep, np.arange(w.shape[-1]), multinomial(w, u))
    idx[..., 0] = 0
    return idx


def systematic(w, u, v, ww, N=None):
    """Conditional systematic resampling (resamplings.py:40-86; Chopin & Singh Alg. 4) for one weight vector and
    one triple of uniforms (u, v, ww).  All arithmetic in ``w.dtype``; the CDF is accumulated in the canonical
    order shared with the kernels (the reference's ``jnp.cumsum`` order is unspecified).

    ``searchsorted`` of an increasing grid gives a non-decreasing index vector, so the zeros of ``idx`` are its
    first ``n_zero`` entries and ``zero_loc[k] = k`` (resamplings.py:81-84): the roll is by ``floor(n_zero * W)``.
    """
    w = np.asarray(w)
    dt = w.dtype.type
    u, v, ww = dt(u), dt(v), dt(ww)
    M = w.shape[0]
    N = M if N is None else N
    cdf = canonical_cdf(w)
    tmp = dt(N) * w[0]
    tmp_floor = np.floor(tmp)
    if tmp <= 1:
        uniform = tmp * u
    else:
        rem = tmp - tmp_floor
        p_cond = rem * (tmp_floor + dt(1)) / tmp
        uniform = rem * u if v < p_cond else rem + (dt(1) - rem) * u
    linspace = (np.arange(N, dtype=w.dtype) + uniform) / dt(N)
    idx = _search_left(cdf, linspace)
    n_zero = int(np.sum(idx == 0))
    zero_loc = np.flatnonzero(idx == 0)
    if n_zero != 1:
        roll_idx = int(np.floor(dt(n_zero) * ww))
        shift = zero_loc[roll_idx] if roll_idx < zero_loc.shape[0] else -1
        idx = np.roll(idx, -shift)
    return np.clip(idx, 0, M - 1)


def mismatches_vs_plain_cumsum(w, u):
    """Number of multinomial indices that differ between the canonical CDF and ``np.cumsum``."""
    a = multinomial(w, u)
    cdf = np.cumsum(w, axis=-1, dtype=w.dtype)
    r = (cdf[..., -1:] * (1 - u)).astype(cdf.dtype)
    b = np.minimum(_search_left(cdf, r), w.shape[-1] - 1)
    b[..., 0] = 0
    return int(np.sum(a != b))


# --------------------------------------------------------------------------------------------
# model families (closed set compiled into the CUDA sweep; see aux_ssm_samplers_b200/csmc.py)
# --------------------------------------------------------------------------------------------


class LinearGaussianModel:
    """Feynman--Kac pieces with linear-Gaussian prior dynamics and a closed set of potentials.

    M0 = N(m0, P0); Mt = N(F x + b, Q) (time-invariant); potential (for G0 and Gt):
      "flat"      0                                        (test_csmc/common.py:67-81)
      "gaussian"  sum_i log N(y_ti; x_i, sig)              (test_csmc/common.py:57-63, d = 1 there)
      "sv"        sum_i log N(y_ti; 0, exp(x_i / 2))       (auxiliary_csmc.py:41-46)
    ``GaussianDynamics(rho)`` of the reference's tests is F = rho, Q = 1 - rho^2, b = 0, m0 = 0, P0 = 1.
    """

    def __init__(self, m0, P0, F, Q, b, potential="flat", ys=None, sig=1.0, dtype=np.float64):
        self.dtype = dtype
        self.m0, self.P0, self.F, self.Q, self.b = [np.asarray(a, dtype=dtype) for a in (m0, P0, F, Q, b)]
        self.d = self.m0.shape[0]
        # time-varying dynamics: F [T-1, d, d], Q [T-1, d, d], b [T-1, d] -- the leading axis of `Mt.params`, which
        # the reference scans over (csmc.py:103): entry t - 1 is the transition INTO time step t
        self.time_varying = self.F.ndim == 3
        self.cP0 = chol(self.P0)
        self.cQ = np.stack([chol(q) for q in self.Q]) if self.time_varying else chol(self.Q)
        self.potential, self.sig = potential, dtype(sig)
        self.ys = None if ys is None else np.asarray(ys, dtype=dtype)

    def _dyn(self, t):
        if not self.time_varying:
            return self.F, self.b, self.cQ
        return self.F[t - 1], self.b[t - 1], self.cQ[t - 1]

    # prior pieces (t = the time step the transition leads INTO; ignored for time-invariant dynamics)
    def m0_sample(self, eps):
        return self.m0 + mv(self.cP0, eps)

    def m0_logpdf(self, x):
        return mvn_logpdf(x, self.m0, self.cP0)

    def mt_mean(self, x, t=None):
        F, b, _ = self._dyn(t)
        return mv(F, x) + b

    def mt_sample(self, eps, x, t=None):
        return self.mt_mean(x, t) + mv(self._dyn(t)[2], eps)

    def mt_logpdf(self, x_new, x, t=None):
        return mvn_logpdf(x_new, self.mt_mean(x, t), self._dyn(t)[2])

    def pot(self, t, x):
        if self.potential == "flat":
            return np.zeros(x.shape[:-1], dtype=self.dtype)
        y = self.ys[t]
        with np.errstate(all="ignore"):
            if self.potential == "gaussian":
                return np.sum(norm_logpdf(y, x, self.sig), axis=-1)
            if self.potential == "sv":
                return np.sum(norm_logpdf(y, 0.0, np.exp(0.5 * x)), axis=-1)
        raise ValueError(self.potential)


def grad_log_target(model, u):
    """Gradient at ``u`` [T, d] of ``_log_pdf`` (csmc/independent.py:121-137): log M0(u_0) + G0(u_0) +
    sum_t [Gt(u_t) + log Mt(u_t | u_{t-1})], which the reference obtains with ``jax.grad``; analytic here for the
    closed family set (linear-Gaussian prior; flat / Gaussian / SV potential)."""
    u = np.asarray(u, dtype=model.dtype)
    r0 = cho_solve(model.cP0, u[0] - model.m0)
    res = u[1:] - (mv(model.F, u[:-1]) + model.b)
    rt = cho_solve(model.cQ, res.T).T  # one solve per time step
    g = np.zeros_like(u)
    g[0] -= r0
    g[1:] -= rt
    g[:-1] += rt @ model.F
    if model.potential == "gaussian":
        g += (model.ys - u) / model.sig ** 2
    elif model.potential == "sv":
        g += 0.5 * model.ys ** 2 * np.exp(-u) - 0.5
    return g


class BootstrapFK:
    """Bootstrap Feynman--Kac model: propose from the prior, weight by the potential."""

    def __init__(self, model):
        self.m = model

    def m0_sample(self, eps):
        return self.m.m0_sample(eps)

    def g0(self, x):
        return self.m.pot(0, x)

    def mt_sample(self, t, eps, x_prev):
        return self.m.mt_sample(eps, x_prev, t)

    def gt(self, t, x, x_prev):
        return self.m.pot(t, x)


class IndependentAuxFK:
    """Auxiliary FK model with separable proposals, csmc/independent.py:60-73,144-268.

    M0' = N(u_0 + (delta_0/2) grad_0, (delta_0/2) I), Mt' likewise (independent of x_{t-1});
    G0' = G0 + log M0;  Gt' = log Mt(x_t | x_{t-1}) + Gt;  with ``gradient`` the weights also carry
    log N(x; u, s) - log N(x; mean, s) (GradientAuxiliaryG0/Gt, :173-190, :252-268).
    ``grad`` is the gradient of the log target at u (zeros when gradient=False).

    ``gradient=True`` (or 1) is the reference to the letter: ``GradientAuxiliaryG0`` corrects every particle, while
    ``GradientAuxiliaryGt.__call__`` (independent.py:262-266) calls ``jnp.sum`` WITHOUT ``axis=-1``: its correction
    is summed over all particles, i.e. one constant per time step, which cancels in every ``normalize`` and in the
    backward weights -- it is left out here (log-weight histories differ from the reference's by that constant).
    ``gradient=2`` applies the per-particle correction at every step (what the proposal's importance weight calls
    for; not what the reference computes).  Pinned by tests/golden/ref_csmc.npz (ind_g1_*).
    """

    def __init__(self, model, u, sqrt_half_delta, grad=None, gradient=False):
        self.m, self.u, self.s = model, u, np.asarray(sqrt_half_delta, dtype=model.dtype)
        self.grad = np.zeros_like(u) if grad is None else grad
        self.gradient = int(gradient)

    def _mean(self, t):
        return self.u[t] + self.s[t] ** 2 * self.grad[t]

    def m0_sample(self, eps):
        return self._mean(0) + self.s[0] * eps

    def mt_sample(self, t, eps, x_prev):
        return self._mean(t) + self.s[t] * eps

    def _corr(self, t, x):
        if not self.gradient or (self.gradient == 1 and t > 0):
            return 0.0
        return (np.sum(norm_logpdf(x, self.u[t], self.s[t]), axis=-1)
                - np.sum(norm_logpdf(x, self._mean(t), self.s[t]), axis=-1))

    def g0(self, x):
        return self.m.pot(0, x) + self.m.m0_logpdf(x) + self._corr(0, x)

    def gt(self, t, x, x_prev):
        return self.m.mt_logpdf(x, x_prev, t) + self.m.pot(t, x) + self._corr(t, x)


class GuidedAuxFK:
    """Guided auxiliary FK model, examples/stochastic_volatility/auxiliary_guided_csmc.py:29-167
    (``grad=False`` branch): Gaussian proposal N(mu_t, Lambda_t) with K_t = Q (Q + s_t^2 I)^{-1},
    mu_t = x_pred + K_t (u_t - x_pred), Lambda_t = Q - K_t Q; weights = potential + transition +
    log N(x; u_t, s_t) - log N(x; mu_t, Lambda_t)."""

    def __init__(self, model, u, sqrt_half_delta):
        self.m, self.u, self.s = model, u, np.asarray(sqrt_half_delta, dtype=model.dtype)
        d = model.d
        eye = np.eye(d, dtype=model.dtype)

        def gain(P, s):
            K = np.linalg.solve(P + s ** 2 * eye, P).T
            Lam = P - K @ P
            cL = chol(0.5 * (Lam + Lam.T))
            cL = np.where(np.isfinite(cL), cL, s * eye)
            return K.astype(model.dtype), cL.astype(model.dtype)

        T = u.shape[0]
        self.K, self.cL = [None] * T, [None] * T
        self.K[0], self.cL[0] = gain(model.P0, self.s[0])
        for t in range(1, T):
            self.K[t], self.cL[t] = gain(model.Q, self.s[t])

    def _mu(self, t, x_pred):
        return x_pred + mv(self.K[t], self.u[t] - x_pred)

    def m0_sample(self, eps):
        return self._mu(0, self.m.m0) + mv(self.cL[0], eps)

    def g0(self, x):
        mu = self._mu(0, self.m.m0)
        return (self.m.pot(0, x) + self.m.m0_logpdf(x) + np.sum(norm_logpdf(x, self.u[0], self.s[0]), axis=-1)
                - mvn_logpdf(x, mu, self.cL[0]))

    def mt_sample(self, t, eps, x_prev):
        return self._mu(t, self.m.mt_mean(x_prev)) + mv(self.cL[t], eps)

    def gt(self, t, x, x_prev):
        mu = self._mu(t, self.m.mt_mean(x_prev))
        return (self.m.pot(t, x) + self.m.mt_logpdf(x, x_prev)
                + np.sum(norm_logpdf(x, self.u[t], self.s[t]), axis=-1) - mvn_logpdf(x, mu, self.cL[t]))


# --------------------------------------------------------------------------------------------
# csmc.py
# --------------------------------------------------------------------------------------------


def csmc_forward(fk, x_star, normals, u_resample, resampling="multinomial", u_kill=None):
    """``_csmc`` (csmc.py:69-107) for ONE chain.  Returns (w_T, xs [T,N,d], log_ws [T,N], As [T-1,N])."""
    T, N, d = normals.shape
    dtype = normals.dtype
    xs = np.empty((T, N, d), dtype=dtype)
    log_ws = np.empty((T, N), dtype=dtype)
    As = np.empty((max(T - 1, 0), N), dtype=np.int64)
    x = fk.m0_sample(normals[0]).astype(dtype)
    x[0] = x_star[0]
    lw = np.asarray(fk.g0(x), dtype=dtype)
    w = normalize(lw)
    xs[0], log_ws[0] = x, lw
    for t in range(1, T):
        if resampling == "multinomial":
            A = multinomial(w, u_resample[t - 1])
        elif resampling == "killing":
            A = killing(w, u_kill[t - 1], u_resample[t - 1])
        elif resampling == "systematic":  # u_resample is [T-1, 3]: (U, V, W) of resamplings.py:65
            A = systematic(w, *u_resample[t - 1])
        else:
            raise ValueError(resampling)
        x_prev = x[A]
        x = fk.mt_sample(t, normals[t], x_prev).astype(dtype)
        x[0] = x_star[t]
        lw = np.asarray(fk.gt(t, x, x_prev), dtype=dtype)
        w = normalize(lw)
        xs[t], log_ws[t], As[t - 1] = x, lw, A
    return w, xs, log_ws, As


def backward_scanning_pass(u_T, w_T, xs, As):
    """Ancestor tracing (csmc.py:110-124).  Returns (x [T,d], B [T])."""
    T = xs.shape[0]
    B = np.empty((T,), dtype=np.int64)
    B[-1] = categorical(w_T, np.asarray(u_T))
    for t in range(T - 2, -1, -1):
        B[t] = As[t][B[t + 1]]
    return xs[np.arange(T), B], B


def backward_sampling_pass(model, u_backward, w_T, xs, log_ws):
    """Backward sampling (csmc.py:127-149) with ``Pt.logpdf`` = the transition density of ``model`` (the prior model,
    or a different ``Pt`` model: csmc.py:47-50).
    ``u_backward[0]`` draws B_T (keys[0]); ``u_backward[k]`` serves time T-1-k (keys[1:] scanned over
    the reversed history)."""
    T = xs.shape[0]
    dtype = xs.dtype
    B = np.empty((T,), dtype=np.int64)
    x = np.empty((T, xs.shape[-1]), dtype=dtype)
    B[-1] = categorical(w_T, np.asarray(u_backward[0]))
    x[-1] = xs[-1, B[-1]]
    for k, t in enumerate(range(T - 2, -1, -1)):
        lw = np.asarray(model.mt_logpdf(x[t + 1], xs[t], t + 1), dtype=dtype) + log_ws[t]
        w = normalize(lw)
        B[t] = categorical(w, np.asarray(u_backward[k + 1]))
        x[t] = xs[t, B[t]]
    return x, B


def csmc_kernel(fk, model, x_star, normals, u_resample, u_backward, backward, resampling="multinomial",
                u_kill=None, pt_model=None):
    """``kernel`` (csmc.py:52-59) for one chain: returns (x [T,d], updated [T], ancestors [T], history).
    ``pt_model``: the model whose transition density is ``Pt.logpdf`` in the backward pass (default: ``model``)."""
    w_T, xs, log_ws, As = csmc_forward(fk, x_star, normals, u_resample, resampling, u_kill)
    if backward:
        x, B = backward_sampling_pass(model if pt_model is None else pt_model, u_backward, w_T, xs, log_ws)
    else:
        x, B = backward_scanning_pass(u_backward[0], w_T, xs, As)
    return x, B != 0, B, (w_T, xs, log_ws, As)


def aux_u(x, delta, normals_aux):
    """u = x + sqrt(delta_t / 2) eps (csmc/generic.py:61-67); delta scalar or [T]."""
    T = x.shape[0]
    s = np.sqrt(0.5 * np.broadcast_to(np.asarray(delta, dtype=x.dtype), (T,))).astype(x.dtype)
    return x + s[:, None] * normals_aux, s

"""Conditional SMC particle Gibbs on B200: host-side mirror of the reference's cSMC API.

Same names and argument meaning as the reference:

* ``CSMCState``, ``Distribution``, ``UnivariatePotential``, ``Dynamics``, ``Potential``
  (aux_samplers/_primitives/csmc/base.py:18-71)
* ``get_kernel(M0, G0, Mt, Gt, N, backward=False, Pt=None)`` -> ``(init, kernel)``; ``kernel(key, state)``
  (aux_samplers/_primitives/csmc/csmc.py:16-66)
* ``get_generic_kernel(factory, N, backward=False, Pt=None)`` -> ``kernel(key, state, delta)``
  (aux_samplers/csmc/generic.py:14-79; raises ValueError when ``backward`` without ``Pt``, :44-47)
* ``get_independent_kernel(M0, G0, Mt, Gt, N, backward=False, Pt=None, gradient=False, parallel=False)``
  (aux_samplers/csmc/independent.py:18-75)
* unfused primitives ``normalize``, ``multinomial``, ``killing``, ``categorical``
  (math/utils.py:24-39, resamplings.py:14-37, csmc.py:111)

The reference calls user Python callables inside its time loop.  A fused CUDA sweep cannot, so there are two paths:

* *descriptors* from a closed family set (see include/auxssm.h) -- a linear-Gaussian prior (``GaussianDistribution`` +
  ``LinearGaussianDynamics``, time-invariant or time-varying) with a ``Flat`` / ``GaussianObservation`` /
  ``StochasticVolatility`` potential, proposed from the prior (bootstrap), from the independent auxiliary proposal, or
  from the guided auxiliary proposal -- run the whole sweep in ONE persistent kernel;
* any other model given as objects with the reference's callable interface (``M0.sample(key, N)``, ``G0(x)``,
  ``Mt.sample(key, x, params_t)``, ``Gt(x_t, x_prev, params_t)``, ``Pt.logpdf``; csmc/base.py:24-71) runs through
  ``_get_callable_kernel``: the time loop stays on the host like ``lax.scan``, the user's callables are evaluated once
  per step, and normalisation / conditional resampling / categorical draws use the unfused kernels.

``state.x`` is ``[C, T, d]`` (chains leading) or ``[T, d]``.  ``key`` is a draws object (``draws.py``);
call order: normals ``[C,T,d]`` (auxiliary kernels only), normals ``[C,T,N,d]``, uniforms ``[C,T-1,N]``,
(killing: uniforms ``[C,T-1,N]``), uniforms ``[C,T]``.
"""
import ctypes
from typing import NamedTuple, Optional

import torch

from . import _lib
from ._lib import AuxssmError, CsmcDesc


class CSMCState(NamedTuple):
    x: torch.Tensor
    updated: torch.Tensor


def cholesky(A):
    """Lower Cholesky factors of ``A[..., d, d]`` computed by libauxssm (``auxssm_cholesky_*``): an all-NaN factor
    where the factorisation fails (``jnp.linalg.cholesky``), a plain square root when ``d == 1``."""
    dev = _lib.require_cuda(A)
    Ac = A.contiguous()
    L = torch.empty_like(Ac)
    d = A.shape[-1]
    _lib.call("auxssm_cholesky_" + _lib.suffix(A.dtype), _lib.stream_ptr(dev), Ac.numel() // (d * d), d, _lib.ptr(Ac),
              _lib.ptr(L))
    return L


# ---- model-family descriptors (mirroring csmc/base.py's abstract classes) -------------------------

class Distribution:
    pass


class UnivariatePotential:
    pass


class Dynamics:
    params = None


class Potential:
    params = None


class GaussianDistribution(Distribution, UnivariatePotential):
    """x_0 ~ N(m0, P0) (the role of ``M0`` in auxiliary_csmc.py:25-31)."""

    def __init__(self, m0, P0):
        self.m0, self.P0 = m0, P0
        self.chol_P0 = cholesky(P0)  # auxssm_cholesky_* (jnp.linalg.cholesky semantics; sqrt when d = 1)

    def logpdf(self, x):  # presence of this attribute is what the reference checks (csmc.py:49)
        raise NotImplementedError("evaluated inside the fused kernel")


class LinearGaussianDynamics(Dynamics):
    """x_t ~ N(F x_{t-1} + b, Q) (``Mt`` of auxiliary_csmc.py:33-40; ``GaussianDynamics(rho)`` of the reference's
    tests is F = rho, Q = 1 - rho^2, b = 0).  Time-invariant ``F [d,d], Q [d,d], b [d]``, or time-varying
    ``F [T-1,d,d], Q [T-1,d,d], b [T-1,d]`` -- the leading axis of the reference's ``Mt.params``, scanned over t
    (csmc.py:103): entry t - 1 is the transition into time step t."""

    def __init__(self, F, Q, b):
        self.F, self.Q, self.b = F, Q, b
        self.time_varying = F.dim() == 3
        if self.time_varying and not (Q.dim() == 3 and b.dim() == 2 and Q.shape[0] == F.shape[0] == b.shape[0]):
            raise ValueError("time-varying dynamics need F [T-1,d,d], Q [T-1,d,d] and b [T-1,d]")
        self.chol_Q = cholesky(Q)
        self.params = (F, b, self.chol_Q) if self.time_varying else None

    def logpdf(self, x_t_p_1, x_t, params):
        raise NotImplementedError("evaluated inside the fused kernel")


class FlatPotential(Potential, UnivariatePotential):
    """test_csmc/common.py:67-81."""
    kind = _lib.POT_FLAT
    ys, sig = None, 1.0


class GaussianObservationPotential(Potential, UnivariatePotential):
    """sum_i log N(y_ti; x_i, sig)  (test_csmc/common.py:57-63); ``ys`` is [T, d] (or [C, T, d])."""
    kind = _lib.POT_GAUSSIAN

    def __init__(self, ys, sig=1.0):
        self.ys, self.sig = ys, float(sig)


class StochasticVolatilityPotential(Potential, UnivariatePotential):
    """sum_i log N(y_ti; 0, exp(x_i / 2))  (auxiliary_csmc.py:41-46)."""
    kind = _lib.POT_SV
    sig = 1.0

    def __init__(self, ys):
        self.ys = ys


class _Aux(NamedTuple):
    proposal: int
    u: torch.Tensor
    sqrt_half_delta: torch.Tensor
    grad: Optional[torch.Tensor]
    gradient: bool
    K: Optional[torch.Tensor]
    chol_Lam: Optional[torch.Tensor]


class AuxiliaryFeynmanKac(NamedTuple):
    """What a ``factory(u, sqrt_half_delta)`` returns in place of the reference's four JAX callables."""
    M0: GaussianDistribution
    G0: UnivariatePotential
    Mt: LinearGaussianDynamics
    Gt: Potential
    aux: _Aux

    def __iter__(self):  # so that ``m0, g0, mt, gt = factory(...)`` keeps working
        return iter((self.M0, self.G0, self.Mt, self.Gt))


def _check_family(M0, G0, Mt, Gt):
    if not isinstance(M0, GaussianDistribution) or not isinstance(Mt, LinearGaussianDynamics):
        raise NotImplementedError(
            "the fused cSMC sweep supports GaussianDistribution / LinearGaussianDynamics priors only; "
            "use the unfused primitives (normalize, multinomial, killing, categorical) for other models")
    for g in (G0, Gt):
        if not hasattr(g, "kind"):
            raise NotImplementedError("unsupported potential %r (closed family set, see include/auxssm.h)" % (g,))
    if G0.kind != Gt.kind:
        raise NotImplementedError("G0 and Gt must be the same potential family")
    return Gt


def _sweep(x_star, M0, Mt, pot, N, backward, resampling, key, aux=None, return_history=False, Pt=None):
    squeeze = x_star.dim() == 2
    x = x_star.unsqueeze(0) if squeeze else x_star
    x = x.contiguous()
    dev = _lib.require_cuda(x)
    dtype = x.dtype
    sfx = _lib.suffix(dtype)
    C, T, d = x.shape
    desc = CsmcDesc()
    desc.T, desc.N, desc.d, desc.n_chains = T, N, d, C
    desc.proposal = aux.proposal if aux is not None else _lib.PROP_BOOTSTRAP
    desc.potential = pot.kind
    desc.resampling = {"multinomial": _lib.RESAMPLE_MULTINOMIAL, "killing": _lib.RESAMPLE_KILLING,
                       "systematic": _lib.RESAMPLE_SYSTEMATIC}[resampling]
    desc.backward = 1 if backward else 0
    desc.gradient = int(aux.gradient) if aux is not None else 0
    desc.sig = float(pot.sig)

    def cvt(t):
        if t is None:
            return None
        _lib.require_cuda(x, t)
        return t.to(dtype=dtype).contiguous()

    if Mt.time_varying:
        if Mt.F.shape[0] != T - 1:
            raise AuxssmError("time-varying dynamics carry T - 1 = %d transitions, got %d" % (T - 1, Mt.F.shape[0]))
        desc.dyn_time_varying = 1
    ptf_ = ptb_ = ptq_ = None
    if backward and Pt is not None and Pt is not Mt:  # csmc.py:47-50: Pt.logpdf drives the backward weights
        if Pt.time_varying != Mt.time_varying or Pt.F.shape != Mt.F.shape:
            raise AuxssmError("Pt must have the same (time-varying or not) layout as Mt")
        ptf_, ptb_, ptq_ = cvt(Pt.F), cvt(Pt.b), cvt(Pt.chol_Q)
        desc.pt_F, desc.pt_b, desc.pt_chol_Q = ptf_.data_ptr(), ptb_.data_ptr(), ptq_.data_ptr()
    ys = None
    if pot.ys is not None:
        ys = cvt(pot.ys)
        desc.ys_chain_stride = 0 if ys.dim() == 2 else T * d
    u = shd = grad = K = cL = None
    if aux is not None:
        u, shd, grad, K, cL = cvt(aux.u), cvt(aux.sqrt_half_delta), cvt(aux.grad), cvt(aux.K), cvt(aux.chol_Lam)
        desc.gains_time_stride = 0 if (K is None or K.dim() == 2) else d * d
    in_kernel = bool(getattr(key, "in_kernel", False)) and hasattr(key, "counter")
    if in_kernel:
        # counter-based draws generated inside the sweep (include/auxssm.h: rng_seed / rng_counter_ptr): the
        # [C,T,N,d] normals and [C,T-1,N] uniforms never exist in HBM
        normals = u_res = u_kill = u_bwd = None
        desc.rng_seed = key.seed
        desc.rng_counter_ptr = key.counter.data_ptr()
    else:
        normals = key.normal((C, T, N, d), dtype, dev).contiguous()
        # systematic resampling consumes the three uniforms (U, V, W) of resamplings.py:65 per step
        u_res = key.uniform((C, max(T - 1, 0), 3 if resampling == "systematic" else N), dtype, dev).contiguous()
        u_kill = key.uniform((C, max(T - 1, 0), N), dtype, dev).contiguous() if resampling == "killing" else None
        u_bwd = key.uniform((C, T), dtype, dev).contiguous()
    x_out = torch.empty_like(x)
    anc = torch.empty((C, T), dtype=torch.int32, device=dev)
    xs_h = lw_h = a_h = None
    scratch = None
    L = _lib.lib()
    if return_history:
        xs_h = torch.empty((C, T, N, d), dtype=dtype, device=dev)
        lw_h = torch.empty((C, T, N), dtype=dtype, device=dev)
        a_h = torch.empty((C, max(T - 1, 0), N), dtype=torch.int32, device=dev)
        nbytes = 0
    else:
        nbytes = L.auxssm_csmc_scratch_bytes(desc, x.element_size())
        scratch = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
    # converted copies must outlive the (asynchronous) call: keep them in locals, never pass temporaries
    m0_, lp0_, f_, b_, lq_ = cvt(M0.m0), cvt(M0.chol_P0), cvt(Mt.F), cvt(Mt.b), cvt(Mt.chol_Q)
    _lib.call("auxssm_csmc_sweep_" + sfx, _lib.stream_ptr(dev), desc, _lib.ptr(x), _lib.ptr(m0_),
              _lib.ptr(lp0_), _lib.ptr(f_), _lib.ptr(b_), _lib.ptr(lq_),
              _lib.ptr(ys), _lib.ptr(u), _lib.ptr(shd), _lib.ptr(grad), _lib.ptr(K), _lib.ptr(cL),
              _lib.ptr(normals), _lib.ptr(u_res), _lib.ptr(u_kill), _lib.ptr(u_bwd), _lib.ptr(x_out), _lib.ptr(anc),
              _lib.ptr(xs_h), _lib.ptr(lw_h), _lib.ptr(a_h), _lib.ptr(scratch), int(nbytes))
    if in_kernel:
        key.advance(4 if resampling == "killing" else 3)
    updated = anc != 0
    if squeeze:
        x_out, updated, anc = x_out[0], updated[0], anc[0]
    state = CSMCState(x=x_out, updated=updated)
    if return_history:
        return state, dict(ancestors=anc, xs=xs_h, log_ws=lw_h, As=a_h)
    return state


def get_kernel(M0, G0, Mt, Gt, N, backward=False, Pt=None, *, resampling="multinomial", _aux=None,
               return_history=False):
    """cSMC kernel of csmc.py:16-66: returns ``(init, kernel)``; ``kernel(key, state) -> CSMCState``."""
    if backward and Pt is None:
        Pt = Mt
    elif backward and not hasattr(Pt, "logpdf"):
        raise ValueError("When `backward` is True, `Pt` must implement a valid logpdf method.")
    if _aux is None and _is_user_model(M0, G0, Mt, Gt):
        return _get_callable_kernel(M0, G0, Mt, Gt, N, backward, Pt, resampling)
    pot = _check_family(M0, G0, Mt, Gt)
    if backward and Pt is not Mt and not isinstance(Pt, LinearGaussianDynamics):
        raise NotImplementedError("backward sampling needs a LinearGaussianDynamics `Pt` (closed family set)")

    def kernel(key, state):
        return _sweep(state.x, M0, Mt, pot, N, backward, resampling, key, _aux, return_history, Pt if backward else None)

    def init(x_star):
        T = x_star.shape[-2]
        shape = x_star.shape[:-2] + (T,)
        return CSMCState(x=x_star, updated=torch.ones(shape, dtype=torch.bool, device=x_star.device))

    return init, kernel


# ---- user-defined models: the reference's callables, one time step at a time --------------------------------------

def _is_user_model(M0, G0, Mt, Gt):
    """True when the four pieces are user objects with the reference's callable interface (csmc/base.py:24-71)
    rather than descriptors of the closed family set."""
    closed = isinstance(M0, GaussianDistribution) or isinstance(Mt, LinearGaussianDynamics) or hasattr(G0, "kind") \
        or hasattr(Gt, "kind")
    return (not closed) and callable(getattr(M0, "sample", None)) and callable(G0) and \
        callable(getattr(Mt, "sample", None)) and callable(Gt)


def _params_at(params, k):
    """Entry k of a parameter pytree whose leaves have a leading time axis (what ``lax.scan`` does with
    ``Mt.params`` / ``Gt.params``, csmc.py:103)."""
    if params is None:
        return None
    if isinstance(params, (tuple, list)):
        return type(params)(_params_at(p, k) for p in params)
    if isinstance(params, dict):
        return {n: _params_at(p, k) for n, p in params.items()}
    return params[k]


def _get_callable_kernel(M0, G0, Mt, Gt, N, backward, Pt, resampling):
    """``get_kernel`` of csmc.py:16-66 for ARBITRARY user models: ``M0.sample(key, N)``, ``G0(x)``,
    ``Mt.sample(key, x_prev, params_t)``, ``Gt(x_t, x_prev, params_t)`` and ``Pt.logpdf(x_next, xs_t, params_t)`` are the
    user's own (torch) callables, exactly the reference's signatures; they are called once per time step from the host,
    like ``lax.scan`` calls them (csmc.py:82-103, :134-146), and everything between them -- ``normalize``
    (math/utils.py:24-39), conditional resampling (resamplings.py:14-86), the categorical draws of the backward pass
    (csmc.py:111,131,138) -- runs in the library's unfused kernels.  Slower than the fused sweeps (2-4 launches per time
    step), but any model works.  Callables see tensors with a leading chain axis when ``state.x`` is ``[C, T, d]``
    (``M0.sample(key, N) -> [C, N, d]`` etc.) and none when it is ``[T, d]``; ``key`` is the draws object (``draws.py``).
    Draw order: per step the callables' own draws, then ``uniform((.., N))`` (or ``(.., 3)`` systematic; killing draws
    the kill uniforms first), and in the backward pass one ``uniform((..,))`` per step starting with B_T."""
    def resample(key, w):
        shape = w.shape
        if resampling == "multinomial":
            return multinomial(key.uniform(shape, w.dtype, w.device), w)
        if resampling == "killing":
            uk = key.uniform(shape, w.dtype, w.device)
            return killing(uk, key.uniform(shape, w.dtype, w.device), w)
        if resampling == "systematic":
            return systematic(key.uniform(shape[:-1] + (3,), w.dtype, w.device), w)
        raise ValueError(resampling)

    def take(x, idx):  # x [.., N, d], idx [.., K] -> [.., K, d]
        return torch.gather(x, -2, idx.long().unsqueeze(-1).expand(idx.shape + (x.shape[-1],)))

    def kernel(key, state):
        x_star = state.x
        T = x_star.shape[-2]
        batch = x_star.shape[:-2]
        xs, lws, As = [], [], []
        x = M0.sample(key, N)
        _lib.require_cuda(x_star, x)
        x = x.clone()
        x[..., 0, :] = x_star[..., 0, :]                       # csmc.py:76
        lw = G0(x)
        w = normalize(lw)
        xs.append(x); lws.append(lw)
        for t in range(1, T):
            A = resample(key, w)
            xp = take(x, A)
            x = Mt.sample(key, xp, _params_at(getattr(Mt, "params", None), t - 1)).clone()
            x[..., 0, :] = x_star[..., t, :]                   # :92
            lw = Gt(x, xp, _params_at(getattr(Gt, "params", None), t - 1))
            w = normalize(lw)
            xs.append(x); lws.append(lw); As.append(A)
        B = categorical(key.uniform(batch, w.dtype, w.device), w)
        out_x, out_B = [None] * T, [None] * T
        out_B[T - 1] = B
        out_x[T - 1] = take(xs[T - 1], B.unsqueeze(-1)).squeeze(-2)
        for t in range(T - 2, -1, -1):
            if backward:                                       # csmc.py:134-146
                lwb = Pt.logpdf(out_x[t + 1].unsqueeze(-2), xs[t], _params_at(getattr(Pt, "params", None), t)) + lws[t]
                B = categorical(key.uniform(batch, w.dtype, w.device), normalize(lwb))
            else:                                              # :116-121
                B = torch.gather(As[t], -1, B.long().unsqueeze(-1)).squeeze(-1)
            out_B[t] = B
            out_x[t] = take(xs[t], B.unsqueeze(-1)).squeeze(-2)
        anc = torch.stack(out_B, dim=-1)
        return CSMCState(x=torch.stack(out_x, dim=-2), updated=anc != 0)

    def init(x_star):
        T = x_star.shape[-2]
        return CSMCState(x=x_star, updated=torch.ones(x_star.shape[:-2] + (T,), dtype=torch.bool, device=x_star.device))

    return init, kernel


def _aux_u(key, x, delta):
    """u = x + sqrt(delta_t / 2) eps with per-time-step delta (csmc/generic.py:61-67)."""
    from .kalman import auxiliary_draw
    C, T, d = x.shape
    if torch.is_tensor(delta):
        dl = torch.as_tensor(delta, dtype=x.dtype, device=x.device)
    else:
        from .kalman import _as_delta
        dl = _as_delta(delta, 1, x.dtype, x.device)[0]  # cached device scalar: no blocking host copy per sweep
    dl = dl.expand(T) if dl.dim() == 0 else dl
    shared = dl.dim() == 1
    dl2 = dl.expand(C, T) if shared else dl
    eps = key.normal((C, T, d), x.dtype, x.device)
    u = auxiliary_draw(x.reshape(C * T, d), eps.reshape(C * T, d), dl2.reshape(C * T)).reshape(C, T, d)
    shd = torch.sqrt(0.5 * dl)
    return u, (shd.expand(C, T) if shared else shd)  # shared schedule: stride-0 chain axis (see guided_factory)


def get_generic_kernel(factory, N, backward=False, Pt=None, *, resampling="multinomial", return_history=False):
    """Auxiliary particle Gibbs with generic proposals (csmc/generic.py:14-79).

    ``factory(u [C,T,d], sqrt_half_delta [C,T]) -> AuxiliaryFeynmanKac`` (see ``independent_factory`` and
    ``stochastic_volatility.guided_factory``), or four user callables ``(M0, G0, Mt, Gt)`` with the reference's
    interface, which run on the per-step path (``_get_callable_kernel``).  ``kernel(key, state, delta)``; ``delta``
    scalar, ``[T]`` or ``[C,T]``.
    """
    if backward and Pt is None:
        raise ValueError("If backward is True, the true dynamics `Pt` must be provided.")
    elif backward and not hasattr(Pt, "logpdf"):
        raise ValueError("`Pt` must implement a valid logpdf method.")

    def kernel(key, state, delta):
        x = state.x
        squeeze = x.dim() == 2
        xb = (x.unsqueeze(0) if squeeze else x).contiguous()
        u, shd = _aux_u(key, xb, delta)
        fk = factory(u, shd)
        if not isinstance(fk, AuxiliaryFeynmanKac):
            # csmc/generic.py:69-72 with the user's own callables: the per-step path on the unfused kernels.  A user
            # factory sees exactly what the reference hands it: u [T, d], scale [T] when the state is one chain.
            m0, g0, mt, gt = factory(u[0], shd[0]) if squeeze else fk
            if not _is_user_model(m0, g0, mt, gt):
                raise NotImplementedError("factory must return an AuxiliaryFeynmanKac (closed family set) or four user "
                                          "callables with the reference's interface (csmc/base.py:24-71)")
            _, inner = _get_callable_kernel(m0, g0, mt, gt, N, backward, Pt, resampling)
            return inner(key, CSMCState(x, state.updated))
        _, inner = get_kernel(fk.M0, fk.G0, fk.Mt, fk.Gt, N, backward, Pt, resampling=resampling, _aux=fk.aux,
                              return_history=return_history)
        out = inner(key, CSMCState(xb, state.updated))
        if squeeze:
            if return_history:
                st, h = out
                return CSMCState(st.x[0], st.updated[0]), h
            return CSMCState(out.x[0], out.updated[0])
        return out

    def init(x):
        T = x.shape[-2]
        return CSMCState(x=x, updated=torch.zeros(x.shape[:-2] + (T,), dtype=torch.bool, device=x.device))

    return init, kernel


def grad_log_target(u, M0, G0, Mt, Gt):
    """Gradient at ``u`` [C,T,d] of log M0(u_0) + G0(u_0) + sum_t (Gt(u_t) + log Mt(u_t | u_{t-1})), the quantity
    the reference obtains with ``jax.grad(_log_pdf)`` (independent.py:62,121-137); analytic for the closed
    family set, computed by ``auxssm_csmc_grad_log_target_*`` (one thread per (chain, time step))."""
    pot = _check_family(M0, G0, Mt, Gt)
    dev = _lib.require_cuda(u)
    dtype = u.dtype
    u = u.contiguous()
    C, T, d = u.shape
    desc = CsmcDesc()
    desc.T, desc.N, desc.d, desc.n_chains = T, 1, d, C
    desc.potential, desc.sig = pot.kind, float(pot.sig)
    cvt = lambda t: t.to(dtype=dtype).contiguous() if _lib.require_cuda(u, t) else None
    ys = None
    if pot.ys is not None:
        ys = cvt(pot.ys)
        desc.ys_chain_stride = 0 if ys.dim() == 2 else T * d
    m0_, lp0_, f_, b_, lq_ = cvt(M0.m0), cvt(M0.chol_P0), cvt(Mt.F), cvt(Mt.b), cvt(Mt.chol_Q)
    grad = torch.empty_like(u)
    _lib.call("auxssm_csmc_grad_log_target_" + _lib.suffix(dtype), _lib.stream_ptr(dev), desc, _lib.ptr(u),
              _lib.ptr(m0_), _lib.ptr(lp0_), _lib.ptr(f_), _lib.ptr(b_), _lib.ptr(lq_), _lib.ptr(ys), _lib.ptr(grad))
    return grad


def independent_factory(M0, G0, Mt, Gt, gradient=False):
    """``factory`` of independent.py:60-73: proposal N(u_t + (delta_t/2) grad_t, (delta_t/2) I).

    ``gradient=True`` reproduces the reference to the letter: ``GradientAuxiliaryG0`` (independent.py:173-190)
    corrects the weight of every particle at t = 0, whereas ``GradientAuxiliaryGt`` (:252-268) sums its correction
    over all particles -- a per-step constant that cancels in every normalisation.  ``gradient="per_particle"``
    applies the per-particle correction at every step instead (the importance weight the proposal calls for)."""
    _check_family(M0, G0, Mt, Gt)
    mode = {False: 0, True: 1, "per_particle": 2, 0: 0, 1: 1, 2: 2}[gradient]

    def factory(u, scale):
        grad = grad_log_target(u, M0, G0, Mt, Gt) if mode else None
        return AuxiliaryFeynmanKac(M0, G0, Mt, Gt, _Aux(_lib.PROP_INDEPENDENT, u, scale, grad, mode, None, None))

    return factory


def get_independent_kernel(M0, G0, Mt, Gt, N, backward=False, Pt=None, gradient=False, parallel=False, *,
                           resampling="multinomial", return_history=False):
    """Auxiliary particle Gibbs with independent (separable) proposals (independent.py:18-75).
    ``gradient``: ``False`` / ``True`` (the reference's weights, see ``independent_factory``) / ``"per_particle"``."""
    if parallel:
        return _get_parallel_kernel(M0, G0, Mt, Gt, N, gradient)
    return get_generic_kernel(independent_factory(M0, G0, Mt, Gt, gradient), N, backward, Pt,
                              resampling=resampling, return_history=return_history)


def pit_sweep(x_star, M0, Mt, pot, N, key, u, sqrt_half_delta, grad=None):
    """One parallel-in-time cSMC sweep (``_primitives/csmc/pit/csmc.py:68-114``) with the independent auxiliary
    proposals of independent.py:78-97: ``auxssm_pit_csmc_sweep_*``.  ``x_star, u [C,T,d]``, ``sqrt_half_delta [C,T]``,
    ``grad [C,T,d]`` or None.  Draws: ``key.normal((C,T,N,d))`` then ``key.uniform((C,T,N))`` (row t: the stitch whose
    right-hand block starts at t), or generated inside the kernels for an in-kernel Philox key."""
    x = x_star.contiguous()
    dev = _lib.require_cuda(x, u, sqrt_half_delta, grad)
    dtype = x.dtype
    C, T, d = x.shape
    desc = CsmcDesc()
    desc.T, desc.N, desc.d, desc.n_chains = T, N, d, C
    desc.proposal, desc.potential, desc.sig = _lib.PROP_INDEPENDENT, pot.kind, float(pot.sig)
    desc.gradient = 0 if grad is None else 1

    def cvt(t):
        if t is None:
            return None
        _lib.require_cuda(x, t)
        return t.to(dtype=dtype).contiguous()

    ys = None
    if pot.ys is not None:
        ys = cvt(pot.ys)
        desc.ys_chain_stride = 0 if ys.dim() == 2 else T * d
    u_, shd_, grad_ = cvt(u), cvt(sqrt_half_delta.expand(C, T)), cvt(grad)
    in_kernel = bool(getattr(key, "in_kernel", False)) and hasattr(key, "counter")
    if in_kernel:
        normals = uniforms = None
        desc.rng_seed = key.seed
        desc.rng_counter_ptr = key.counter.data_ptr()
    else:
        normals = key.normal((C, T, N, d), dtype, dev).contiguous()
        uniforms = key.uniform((C, T, N), dtype, dev).contiguous()
    x_out = torch.empty_like(x)
    anc = torch.empty((C, T), dtype=torch.int32, device=dev)
    L = _lib.lib()
    nbytes = L.auxssm_pit_csmc_scratch_bytes(desc, x.element_size())
    if nbytes == 0:
        _lib.check(-1, "auxssm_pit_csmc_scratch_bytes")
    scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    m0_, lp0_, f_, b_, lq_ = cvt(M0.m0), cvt(M0.chol_P0), cvt(Mt.F), cvt(Mt.b), cvt(Mt.chol_Q)
    _lib.call("auxssm_pit_csmc_sweep_" + _lib.suffix(dtype), _lib.stream_ptr(dev), desc, _lib.ptr(x), _lib.ptr(m0_),
              _lib.ptr(lp0_), _lib.ptr(f_), _lib.ptr(b_), _lib.ptr(lq_), _lib.ptr(ys), _lib.ptr(u_), _lib.ptr(shd_),
              _lib.ptr(grad_), _lib.ptr(normals), _lib.ptr(uniforms), _lib.ptr(x_out), _lib.ptr(anc), _lib.ptr(scratch),
              int(nbytes))
    if in_kernel:
        key.advance(2)
    return x_out, anc


def _get_parallel_kernel(M0, G0, Mt, Gt, N, gradient):
    """``_get_parallel_kernel`` of independent.py:78-118: the parallel-in-time particle Gibbs kernel.
    ``kernel(key, state, delta) -> CSMCState``; ``delta`` scalar, ``[T]`` or ``[C, T]``."""
    pot = _check_family(M0, G0, Mt, Gt)
    if gradient not in (False, True, 0, 1):
        raise NotImplementedError("parallel=True supports gradient=False / True (independent.py:80-86)")

    def kernel(key, state, delta):
        x = state.x
        squeeze = x.dim() == 2
        xb = (x.unsqueeze(0) if squeeze else x).contiguous()
        u, shd = _aux_u(key, xb, delta)                                       # independent.py:99-102
        grad = grad_log_target(u, M0, G0, Mt, Gt) if gradient else None      # :81 (jax.grad(_log_pdf))
        x_out, anc = pit_sweep(xb, M0, Mt, pot, N, key, u, shd, grad)
        upd = anc != 0
        if squeeze:
            x_out, upd = x_out[0], upd[0]
        return CSMCState(x=x_out, updated=upd)

    def init(x):
        T = x.shape[-2]
        return CSMCState(x=x, updated=torch.zeros(x.shape[:-2] + (T,), dtype=torch.bool, device=x.device))

    return init, kernel


def guided_gains(P0, Q, sqrt_half_delta):
    """K_t, chol(Lambda_t) for every t (auxiliary_guided_csmc.py:155-157, :44-46); sqrt_half_delta is [T]."""
    dev = _lib.require_cuda(P0, Q, sqrt_half_delta)
    sfx = _lib.suffix(P0.dtype)
    T, d = sqrt_half_delta.shape[0], P0.shape[-1]
    K = torch.empty((T, d, d), dtype=P0.dtype, device=dev)
    cL = torch.empty_like(K)
    P0c, Qc, sc = P0.contiguous(), Q.contiguous(), sqrt_half_delta.contiguous()
    _lib.call("auxssm_guided_gains_" + sfx, _lib.stream_ptr(dev), T, d, _lib.ptr(P0c), _lib.ptr(Qc), _lib.ptr(sc),
              _lib.ptr(K), _lib.ptr(cL))
    return K, cL


def guided_factory(M0, G0, Mt, Gt):
    """``factory`` of auxiliary_guided_csmc.py:19-24 (``grad=False``).  Requires delta shared by all chains."""
    _check_family(M0, G0, Mt, Gt)

    def factory(u, scale):
        # the gains are computed once from chain 0's schedule: every chain must share it.  `_aux_u` hands over a
        # stride-0 (expanded) chain axis when delta is a scalar or [T]; a genuine [C, T] delta is rejected.
        if scale.dim() == 2 and scale.shape[0] > 1 and scale.stride(0) != 0:
            raise AuxssmError("guided_factory needs one delta schedule shared by all chains (scalar or [T]); "
                              "got a per-chain [C, T] delta")
        K, cL = guided_gains(M0.P0.to(u), Mt.Q.to(u), scale[0].contiguous())
        return AuxiliaryFeynmanKac(M0, G0, Mt, Gt, _Aux(_lib.PROP_GUIDED, u, scale, None, False, K, cL))

    return factory


# ---- unfused primitives -----------------------------------------------------------------------------

def _rows(t):
    return t.reshape(-1, t.shape[-1]).contiguous()


def normalize(log_weights):
    """exp(lw - logsumexp(lw)) over the last axis (math/utils.py:24-39)."""
    dev = _lib.require_cuda(log_weights)
    lw = _rows(log_weights)
    w = torch.empty_like(lw)
    _lib.call("auxssm_normalize_" + _lib.suffix(lw.dtype), _lib.stream_ptr(dev), lw.shape[0], lw.shape[1],
              _lib.ptr(lw), _lib.ptr(w))
    return w.reshape(log_weights.shape)


def multinomial(uniforms, weights):
    """Conditional multinomial resampling (resamplings.py:14-37) with supplied uniforms [..., N]."""
    dev = _lib.require_cuda(weights, uniforms)
    w, u = _rows(weights), _rows(uniforms)
    idx = torch.empty(w.shape, dtype=torch.int32, device=dev)
    _lib.call("auxssm_resample_multinomial_" + _lib.suffix(w.dtype), _lib.stream_ptr(dev), w.shape[0], w.shape[1],
              _lib.ptr(w), _lib.ptr(u), _lib.ptr(idx))
    return idx.reshape(weights.shape)


def killing(uniforms_kill, uniforms, weights):
    """Conditional killing resampling (not in the reference; spec in oracle/csmc.py)."""
    dev = _lib.require_cuda(weights, uniforms, uniforms_kill)
    w, u, uk = _rows(weights), _rows(uniforms), _rows(uniforms_kill)
    idx = torch.empty(w.shape, dtype=torch.int32, device=dev)
    _lib.call("auxssm_resample_killing_" + _lib.suffix(w.dtype), _lib.stream_ptr(dev), w.shape[0], w.shape[1],
              _lib.ptr(w), _lib.ptr(uk), _lib.ptr(u), _lib.ptr(idx))
    return idx.reshape(weights.shape)


def systematic(uniforms3, weights, N=None):
    """Conditional systematic resampling (resamplings.py:40-86): ``weights [..., M]``, the three uniforms (U, V, W)
    of resamplings.py:65 as ``uniforms3 [..., 3]``; returns ``N`` (default ``M``) int32 indices per row."""
    dev = _lib.require_cuda(weights, uniforms3)
    w, u = _rows(weights), _rows(uniforms3)
    M = w.shape[1]
    N = M if N is None else int(N)
    idx = torch.empty((w.shape[0], N), dtype=torch.int32, device=dev)
    _lib.call("auxssm_resample_systematic_" + _lib.suffix(w.dtype), _lib.stream_ptr(dev), w.shape[0], M, N, _lib.ptr(w),
              _lib.ptr(u), _lib.ptr(idx))
    return idx.reshape(weights.shape[:-1] + (N,))


def categorical(uniforms, weights):
    """One ``jax.random.choice(key, N, p=w, shape=())`` draw per row; uniforms [...], weights [..., N]."""
    dev = _lib.require_cuda(weights, uniforms)
    w = _rows(weights)
    u = uniforms.reshape(-1).contiguous()
    idx = torch.empty((w.shape[0],), dtype=torch.int32, device=dev)
    _lib.call("auxssm_categorical_" + _lib.suffix(w.dtype), _lib.stream_ptr(dev), w.shape[0], w.shape[1],
              _lib.ptr(w), _lib.ptr(u), _lib.ptr(idx))
    return idx.reshape(weights.shape[:-1])

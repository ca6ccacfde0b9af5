"""Auxiliary Kalman sampler for latent-Gaussian targets, with the whole MH sweep fused into one kernel launch.

Target:  pi(x_{0:T-1}) ~ N(x_0; m0, P0) prod_t N(x_t; F_{t-1} x_{t-1} + b_{t-1}, Q_{t-1}) prod_t exp(g(x_t; y_t)),
g = ``"flat"`` | ``"gaussian"`` (sum_i log N(y_ti; x_ti, sig)) | ``"sv"`` (sum_i log N(y_ti; 0, exp(x_ti / 2))).

``get_kernel(ys, m0, P0, Fs, Qs, bs, parallel, order, potential)`` mirrors the reference's model adapter
``examples/stochastic_volatility/auxiliary_kalman.py:10-57`` (dynamics factory = the prior, first / second order
observations factory, ``log_likelihood_fn`` = prior + potential) on top of ``aux_samplers/kalman/generic.py``, for
time-varying dynamics ``Fs, Qs [T-1, d, d]``, ``bs [T-1, d]`` (or time-invariant ``[d, d]`` / ``[d]``).

* d <= 4 and every CTA of the launch co-resident: ONE launch per sweep (``auxssm_kalman_sweep_*``: warp-shuffle
  scans + decoupled look-back, csrc/kalman_sweep_fused.cu) -- the few-chain, launch-bound regime of configs[0];
* otherwise the sweep is composed from the library's entry points exactly like ``kalman.get_kernel`` does.
"""
import torch

from . import _lib, kalman
from ._lib import AuxssmError, SweepDesc

_POT = {"flat": _lib.POT_FLAT, "gaussian": _lib.POT_GAUSSIAN, "sv": _lib.POT_SV}


def _canon_dyn(t, core, Tm1, C, name):
    """-> ([C or 1, Tm1 or 1, *core] contiguous tensor, (chain stride, time stride) in elements)"""
    nc = len(core)
    if t.dim() == nc:
        t = t.reshape((1, 1) + tuple(core))
    elif t.dim() == nc + 1:
        t = t.unsqueeze(0)
    elif t.dim() != nc + 2:
        raise AuxssmError("%s: expected %d-%d dims, got %s" % (name, nc, nc + 2, tuple(t.shape)))
    if tuple(t.shape[2:]) != tuple(core) or t.shape[0] not in (1, C) or t.shape[1] not in (1, Tm1):
        raise AuxssmError("%s: shape %s does not match [C?, T-1?, %s]" % (name, tuple(t.shape), core))
    t = t.contiguous()
    n = 1
    for s in core:
        n *= s
    return t, (t.shape[1] * n if t.shape[0] > 1 else 0, n if t.shape[1] > 1 else 0)


def fused_sweep(x, m0, P0, Fs, Qs, bs, ys, normals_aux, normals_eps, uniforms, delta, order=1, potential="flat",
                sig=1.0):
    """One MH sweep of every chain in one launch.  ``x, normals_*``: [C, T, d]; ``uniforms``: [C]; ``delta`` scalar
    or [C].  Returns ``(x_new, accepted[bool, C], alpha[C])``.  Raises AuxssmError (code -4, UNSUPPORTED) for d > 4
    or when the launch would not be co-resident."""
    dev = _lib.require_cuda(x, m0, P0, Fs, Qs, bs, ys if potential != "flat" else None, normals_aux, normals_eps,
                            uniforms)
    dtype = x.dtype
    C, T, d = x.shape
    desc = SweepDesc()
    desc.lg.T, desc.lg.dx, desc.lg.dy, desc.lg.n_chains, desc.lg.n_batch = T, d, d, C, 1
    Fc, (desc.lg.Fs.chain, desc.lg.Fs.time) = _canon_dyn(Fs, (d, d), T - 1, C, "Fs")
    Qc, (desc.lg.Qs.chain, desc.lg.Qs.time) = _canon_dyn(Qs, (d, d), T - 1, C, "Qs")
    bc, (desc.lg.bs.chain, desc.lg.bs.time) = _canon_dyn(bs, (d,), T - 1, C, "bs")
    m0c = m0.contiguous()
    P0c = P0.contiguous()
    desc.lg.m0.chain = d if m0c.dim() == 2 and m0c.shape[0] > 1 else 0
    desc.lg.P0.chain = d * d if P0c.dim() == 3 and P0c.shape[0] > 1 else 0
    desc.order, desc.potential, desc.sig = int(order), _POT[potential], float(sig)
    ysc = None
    if potential != "flat":
        ysc = ys.to(dtype).contiguous()
        desc.ys_chain_stride = 0 if ysc.dim() == 2 else T * d
    dl = kalman._as_delta(delta, C, dtype, dev)
    desc.delta_stride = 1 if dl.numel() > 1 else 0
    for t in (Fc, Qc, bc, m0c, P0c, normals_aux, normals_eps, uniforms):
        if t.dtype != dtype:
            raise AuxssmError("all arrays must have dtype %s" % dtype)
    xc, na, ne, un = x.contiguous(), normals_aux.contiguous(), normals_eps.contiguous(), uniforms.contiguous()
    x_out = torch.empty_like(xc)
    acc = torch.empty((C,), dtype=torch.int32, device=dev)
    alpha = torch.empty((C,), dtype=dtype, device=dev)
    L = _lib.lib()
    nbytes = L.auxssm_kalman_sweep_scratch_bytes(desc, x.element_size())
    if nbytes == 0:
        _lib.check(-4, "auxssm_kalman_sweep_scratch_bytes")
    scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    p = _lib.ptr
    _lib.call("auxssm_kalman_sweep_" + _lib.suffix(dtype), _lib.stream_ptr(dev), desc, p(xc), p(m0c), p(P0c), p(Fc),
              p(Qc), p(bc), p(ysc), p(na), p(ne), p(un), p(dl), p(x_out), p(acc), p(alpha), p(scratch), scratch.numel())
    return x_out, acc.bool(), alpha


def get_kernel(ys, m0, P0, Fs, Qs, bs, parallel=True, order=1, potential="sv", sig=1.0, fused="auto"):
    """``(init, kernel)`` with ``kernel(key, state, delta) -> KalmanSampler`` (kalman/generic.py:19-95).

    ``fused="auto"`` takes the one-launch sweep when ``d <= 4`` and the launch is co-resident, ``True`` insists on it,
    ``False`` always composes filter / sampler / log-density calls."""
    d = m0.shape[-1]

    def composed():
        def dynamics_factory(x):
            Tm1 = x.shape[1] - 1
            ex = lambda t, core: t if t.dim() > len(core) else t.reshape((1,) + tuple(core)).expand((Tm1,) + tuple(core))
            return m0, P0, ex(Fs, (d, d)), ex(Qs, (d, d)), ex(bs, (d,))

        def obs_factory(x, u, delta):
            if potential == "sv":
                from . import stochastic_volatility as sv
                return sv.observations_factory(ys, x, u, delta, order)
            raise NotImplementedError("the composed path serves the SV potential; use fused=True for flat / gaussian")

        def log_likelihood_fn(x):
            from . import stochastic_volatility as sv
            lg = kalman.LGSSM(*dynamics_factory(x), None, None, None)
            return sv.log_potential(ys, x, out=kalman.prior_logpdf(x, lg))

        return kalman.get_kernel(dynamics_factory, obs_factory, log_likelihood_fn, parallel)[1]

    slow = [None]

    def kernel(key, state, delta):
        x = state.x
        squeeze = x.dim() == 2
        xb = x.unsqueeze(0) if squeeze else x
        use_fused = fused is True or (fused == "auto" and d <= 4)
        if use_fused:
            dev, dtype = xb.device, xb.dtype
            n_aux = key.normal(tuple(xb.shape), dtype, dev)
            n_eps = key.normal(tuple(xb.shape), dtype, dev)
            un = key.uniform((xb.shape[0],), dtype, dev)
            try:
                x_new, acc, _ = fused_sweep(xb, m0, P0, Fs, Qs, bs, ys, n_aux, n_eps, un, delta, order, potential, sig)
                return kalman.KalmanSampler(x_new[0] if squeeze else x_new, acc[0] if squeeze else acc)
            except AuxssmError as e:
                if fused is True or "code -4" not in str(e):
                    raise
                # not co-resident: replay the SAME draws through the composed path
                from .draws import ExternalDraws
                key = ExternalDraws([n_aux, n_eps], [un])
        if slow[0] is None:
            slow[0] = composed()
        st = slow[0](key, kalman.KalmanSampler(xb, True), delta)
        return kalman.KalmanSampler(st.x[0], st.updated[0]) if squeeze else st

    def init(x):
        return kalman.KalmanSampler(x=x, updated=True)

    return init, kernel

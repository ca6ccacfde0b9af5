"""Auxiliary Kalman sampler on B200: host-side mirror of the reference's Kalman API.

Same names, argument meaning and error behaviour as the reference:

* ``LGSSM``                                  (aux_samplers/_primitives/kalman/base.py:12-69)
* ``filtering(ys, lgssm, parallel)``          (_primitives/kalman/filtering.py:18-46)
* ``sampling(key, ms, Ps, lgssm, parallel)``  (_primitives/kalman/sampling.py:11-40)
* ``posterior_logpdf / prior_logpdf / log_likelihood`` (_primitives/kalman/base.py:73-166)
* ``get_kernel(dynamics_factory, observations_factory, log_likelihood_fn, parallel)``
  -> ``(init, kernel)``; ``kernel(key, state, delta)``   (aux_samplers/kalman/generic.py:19-106)

Arrays are CUDA ``torch`` tensors (device memory only; all arithmetic happens in libauxssm.so).
Layouts: the default is chains-leading, ``[C, T, ...]`` (what ``jax.vmap`` over chains gives) or
unbatched ``[T, ...]``; pass ``inner_batch=True`` for the reference's batched-LGSSM layout
``[T, B, ...]`` (base.py:40-49), optionally with a leading chain axis ``[C, T, B, ...]``.
Any axis may be broadcast (size 1 or ``expand``-ed): it reaches the kernels as a stride of 0, so a
time-invariant ``F`` is passed once instead of tiled (auxiliary_kalman.py:23-25).
``Hs=None`` means ``H_t = I`` and ``cs=None`` means ``c_t = 0``.

Randomness: JAX keys do not exist here.  ``key`` is any object with ``normal(shape, dtype, device)``
and ``uniform(shape, dtype, device)`` methods -- see ``draws.py`` (``ExternalDraws`` replays supplied
arrays in call order, which is how parity with the oracle is checked).
"""
from typing import NamedTuple, Optional

import torch

from . import _lib
from ._lib import AuxssmError, LgssmDesc, Stride


class LGSSM(NamedTuple):
    m0: torch.Tensor
    P0: torch.Tensor
    Fs: torch.Tensor
    Qs: torch.Tensor
    bs: torch.Tensor
    Hs: Optional[torch.Tensor]
    Rs: torch.Tensor
    cs: Optional[torch.Tensor]


class Diagonal(NamedTuple):
    """Wraps ``diag(R_t)`` of shape ``[..., T, dy]`` to be used as ``LGSSM.Rs`` when every observation
    covariance is diagonal (B200-first layout: ``d`` instead of ``d^2`` values per step travel through HBM)."""
    diag: torch.Tensor


class KalmanSampler(NamedTuple):
    """Mirror of ``KalmanSampler(SamplerState)`` (kalman/generic.py:13-16)."""
    x: torch.Tensor
    updated: object


# --------------------------------------------------------------------------------------------
# layout handling
# --------------------------------------------------------------------------------------------

class _Dims(NamedTuple):
    C: int
    B: int
    T: int
    dx: int
    dy: int
    has_chain: bool
    inner_batch: bool


def _canon(t, core, n_time, dims, name):
    """View ``t`` as ``[C, n_time or 1, B, *core]`` without copying batch axes; returns (tensor, Stride)."""
    if t is None:
        return None, Stride(0, 0, 0)
    nc = len(core)
    want_nd = nc + (1 if n_time else 0) + (1 if dims.inner_batch else 0)
    if t.dim() == want_nd + 1:
        pass  # leading chain axis present
    elif t.dim() == want_nd:
        t = t.unsqueeze(0)
    else:
        raise AuxssmError("%s: expected %d or %d dims, got shape %s" % (name, want_nd, want_nd + 1, tuple(t.shape)))
    # now [C?, (T), (B), core]
    if not n_time:
        t = t.unsqueeze(1)
    if not dims.inner_batch:
        t = t.unsqueeze(2)
    if tuple(t.shape[3:]) != tuple(core):
        raise AuxssmError("%s: trailing shape %s, expected %s" % (name, tuple(t.shape[3:]), tuple(core)))
    tgt = (dims.C, n_time if n_time else 1, dims.B) + tuple(core)
    for have, want, ax in zip(t.shape[:3], tgt[:3], ("chain", "time", "batch")):
        if have != want and have != 1:
            raise AuxssmError("%s: %s axis has size %d, expected %d (or 1 to broadcast)" % (name, ax, have, want))
    # core axes must be dense row-major
    exp = 1
    dense = True
    for size, st in zip(reversed(t.shape[3:]), reversed(t.stride()[3:])):
        if size != 1 and st != exp:
            dense = False
        exp *= size
    if not dense:
        t = t.contiguous()
    t = t.expand(tgt)
    st = t.stride()
    return t, Stride(st[0] if dims.C > 1 else 0, st[2] if dims.B > 1 else 0, st[1] if (n_time and n_time > 1) else 0)


def _infer_dims(ys, m0, inner_batch, xs=None):
    ref = ys if ys is not None else xs
    core_nd = 2 + (1 if inner_batch else 0)  # [T, (B), d]
    if ref.dim() == core_nd + 1:
        has_chain, C = True, ref.shape[0]
        T = ref.shape[1]
    elif ref.dim() == core_nd:
        has_chain, C = False, 1
        T = ref.shape[0]
    else:
        raise AuxssmError("observations / trajectories must have %d or %d dims, got %s"
                          % (core_nd, core_nd + 1, tuple(ref.shape)))
    B = ref.shape[-2] if inner_batch else 1
    dx = m0.shape[-1]
    dy = ys.shape[-1] if ys is not None else dx
    if has_chain is False and m0.dim() == (2 if inner_batch else 1) + 1:
        # parameters carry a chain axis although ys does not: treat ys as shared
        has_chain, C = True, m0.shape[0]
    return _Dims(C, B, T, dx, dy, has_chain, inner_batch)


def _fill_desc(dims, parallel, time_chunks, hid, rdiag=False, fdiag=False):
    d = LgssmDesc()
    d.T, d.dx, d.dy = dims.T, dims.dx, dims.dy
    d.n_chains, d.n_batch = dims.C, dims.B
    d.parallel = 1 if parallel else 0
    d.time_chunks = int(time_chunks)
    d.flags = ((_lib.FLAG_H_IDENTITY if hid else 0) | (_lib.FLAG_R_DIAGONAL if rdiag else 0)
               | (_lib.FLAG_F_IS_DIAGONAL if fdiag else 0))
    return d


def _tensors(lgssm):
    return [t.diag if isinstance(t, Diagonal) else t for t in lgssm if t is not None]


def _out_shape(dims, n_time, core):
    shp = ()
    if dims.has_chain:
        shp += (dims.C,)
    if n_time:
        shp += (n_time,)
    if dims.inner_batch:
        shp += (dims.B,)
    return shp + tuple(core)


def _scratch(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _prep_lgssm(d, dims, lgssm, dtype):
    m0, P0, Fs, Qs, bs, Hs, Rs, cs = lgssm
    Tm1 = dims.T - 1
    dx, dy = dims.dx, dims.dy
    out = {}
    out["m0"], d.m0 = _canon(m0, (dx,), 0, dims, "m0")
    out["P0"], d.P0 = _canon(P0, (dx, dx), 0, dims, "P0")
    if Tm1 > 0:
        out["Fs"], d.Fs = _canon(Fs, (dx, dx), Tm1, dims, "Fs")
        out["Qs"], d.Qs = _canon(Qs, (dx, dx), Tm1, dims, "Qs")
        out["bs"], d.bs = _canon(bs, (dx,), Tm1, dims, "bs")
    else:
        out["Fs"] = out["Qs"] = out["bs"] = None
    out["Hs"], d.Hs = _canon(Hs, (dy, dx), dims.T, dims, "Hs")
    if isinstance(Rs, Diagonal):
        out["Rs"], d.Rs = _canon(Rs.diag, (dy,), dims.T, dims, "Rs (diagonal)")
    else:
        out["Rs"], d.Rs = _canon(Rs, (dy, dy), dims.T, dims, "Rs")
    out["cs"], d.cs = _canon(cs, (dy,), dims.T, dims, "cs")
    for k, v in out.items():
        if v is not None and v.dtype != dtype:
            raise AuxssmError("%s has dtype %s, expected %s" % (k, v.dtype, dtype))
    return out


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------

def filtering(ys, lgssm, parallel, *, inner_batch=False, time_chunks=0, f_diagonal=False, marginals=True):
    """Kalman filter; returns ``(ms, Ps, ell)`` like the reference (filtering.py:18-46).
    ``marginals=False`` is the ell-only call (returns ``(None, None, ell)``; nothing but the likelihood is stored).

    ``ell`` is summed over the inner batch (filtering.py:43-45) and has shape ``[C]`` (or ``[]``).
    ``time_chunks`` (parallel only) forces the number of time chunks of the scan; 0 = automatic.
    ``f_diagonal=True`` is a HINT that every ``F_t`` is a diagonal matrix (layout unchanged); together
    with ``Hs=None``, ``Rs=Diagonal(...)`` and time-invariant ``Fs, Qs, bs`` it selects the one-warp-per-chain
    structured kernel.  The library verifies the hint on the device on every call and runs the general kernels
    when it does not hold: a wrong hint costs time, never correctness.
    """
    lgssm = LGSSM(*lgssm)
    dev = _lib.require_cuda(ys, *_tensors(lgssm))
    dtype = ys.dtype
    sfx = _lib.suffix(dtype)
    dims = _infer_dims(ys, lgssm.m0, inner_batch)
    hid = lgssm.Hs is None
    d = _fill_desc(dims, parallel, time_chunks, hid, isinstance(lgssm.Rs, Diagonal), f_diagonal)
    p = _prep_lgssm(d, dims, lgssm, dtype)
    ysc, d.ys = _canon(ys, (dims.dy,), dims.T, dims, "ys")
    ms = Ps = None
    if marginals:
        ms = torch.empty(_out_shape(dims, dims.T, (dims.dx,)), dtype=dtype, device=dev)
        Ps = torch.empty(_out_shape(dims, dims.T, (dims.dx, dims.dx)), dtype=dtype, device=dev)
        _, d.ms = _canon(ms, (dims.dx,), dims.T, dims, "ms")
        _, d.Ps = _canon(Ps, (dims.dx, dims.dx), dims.T, dims, "Ps")
    ell = torch.empty((dims.C,), dtype=dtype, device=dev)
    L = _lib.lib()
    nbytes = L.auxssm_kalman_filter_scratch_bytes(d, ys.element_size())
    scratch = _scratch(nbytes, dev)
    _lib.call("auxssm_kalman_filter_" + sfx,
        _lib.stream_ptr(dev), d, _lib.ptr(ysc), _lib.ptr(p["m0"]), _lib.ptr(p["P0"]), _lib.ptr(p["Fs"]),
        _lib.ptr(p["Qs"]), _lib.ptr(p["bs"]), _lib.ptr(p["Hs"]), _lib.ptr(p["Rs"]), _lib.ptr(p["cs"]),
        _lib.ptr(ms), _lib.ptr(Ps), _lib.ptr(ell), _lib.ptr(scratch), scratch.numel())
    return ms, Ps, (ell if dims.has_chain else ell[0])


def sampling(key, ms, Ps, lgssm, parallel, *, inner_batch=False, time_chunks=0, f_diagonal=False):
    """Backward sampler (sampling.py:11-40).  ``key`` supplies ``normal(ms.shape)`` (sampling.py:128);
    a plain tensor of that shape is accepted as the draws themselves."""
    lgssm = LGSSM(*lgssm)
    dev = _lib.require_cuda(ms, Ps, lgssm.Fs, lgssm.Qs, lgssm.bs)
    dtype = ms.dtype
    sfx = _lib.suffix(dtype)
    eps = key if torch.is_tensor(key) else key.normal(tuple(ms.shape), dtype, dev)
    _lib.require_cuda(ms, eps)
    if tuple(eps.shape) != tuple(ms.shape):
        raise AuxssmError("eps must have the shape of ms, got %s vs %s" % (tuple(eps.shape), tuple(ms.shape)))
    dims = _infer_dims(None, lgssm.m0, inner_batch, xs=ms)
    dims = dims._replace(dx=ms.shape[-1], dy=ms.shape[-1])
    d = _fill_desc(dims, parallel, time_chunks, True, False, f_diagonal)
    dx, Tm1 = dims.dx, dims.T - 1
    msc, d.ms = _canon(ms, (dx,), dims.T, dims, "ms")
    Psc, d.Ps = _canon(Ps, (dx, dx), dims.T, dims, "Ps")
    epsc, d.eps = _canon(eps, (dx,), dims.T, dims, "eps")
    Fs = Qs = bs = None
    if Tm1 > 0:
        Fs, d.Fs = _canon(lgssm.Fs, (dx, dx), Tm1, dims, "Fs")
        Qs, d.Qs = _canon(lgssm.Qs, (dx, dx), Tm1, dims, "Qs")
        bs, d.bs = _canon(lgssm.bs, (dx,), Tm1, dims, "bs")
    xs = torch.empty(_out_shape(dims, dims.T, (dx,)), dtype=dtype, device=dev)
    _, d.xs = _canon(xs, (dx,), dims.T, dims, "xs")
    L = _lib.lib()
    scratch = _scratch(L.auxssm_kalman_sample_scratch_bytes(d, ms.element_size()), dev)
    _lib.call("auxssm_kalman_sample_" + sfx,
        _lib.stream_ptr(dev), d, _lib.ptr(msc), _lib.ptr(Psc), _lib.ptr(Fs), _lib.ptr(Qs), _lib.ptr(bs),
        _lib.ptr(epsc), _lib.ptr(xs), _lib.ptr(scratch), scratch.numel())
    return xs


def _logpdf(which, ys, xs, ell, lgssm, inner_batch, f_diagonal=False):
    lgssm = LGSSM(*lgssm)
    dev = _lib.require_cuda(xs, ys, ell if torch.is_tensor(ell) else None, *_tensors(lgssm))
    dtype = xs.dtype
    sfx = _lib.suffix(dtype)
    dims = _infer_dims(ys, lgssm.m0, inner_batch, xs=xs)
    dims = dims._replace(dx=xs.shape[-1])
    hid = lgssm.Hs is None
    d = _fill_desc(dims, False, 0, hid, isinstance(lgssm.Rs, Diagonal), f_diagonal)
    p = _prep_lgssm(d, dims, lgssm, dtype)
    ysc = None
    if ys is not None:
        ysc, d.ys = _canon(ys, (dims.dy,), dims.T, dims, "ys")
    xsc, d.xs = _canon(xs, (dims.dx,), dims.T, dims, "xs")
    ellc = None
    if which & _lib.LOGPDF_MINUS_ELL:
        ellc = torch.as_tensor(ell, dtype=dtype, device=dev).reshape(-1).contiguous()
        if ellc.numel() != dims.C:
            raise AuxssmError("ell must have one entry per chain")
    out = torch.empty((dims.C,), dtype=dtype, device=dev)
    L = _lib.lib()
    scratch = _scratch(L.auxssm_lgssm_logpdf_scratch_bytes(d, xs.element_size()), dev)
    _lib.call("auxssm_lgssm_logpdf_" + sfx,
        _lib.stream_ptr(dev), d, which, _lib.ptr(ysc), _lib.ptr(xsc), _lib.ptr(ellc), _lib.ptr(p["m0"]),
        _lib.ptr(p["P0"]), _lib.ptr(p["Fs"]), _lib.ptr(p["Qs"]), _lib.ptr(p["bs"]), _lib.ptr(p["Hs"]),
        _lib.ptr(p["Rs"]), _lib.ptr(p["cs"]), _lib.ptr(out), _lib.ptr(scratch), scratch.numel())
    return out if dims.has_chain else out[0]


def posterior_logpdf(ys, xs, ell, lgssm, *, inner_batch=False, f_diagonal=False):
    """log p(xs | ys) of a trajectory (base.py:73-96).  ``f_diagonal``: same promise as in ``filtering``."""
    return _logpdf(_lib.LOGPDF_POSTERIOR, ys, xs, ell, lgssm, inner_batch, f_diagonal)


def prior_logpdf(xs, lgssm, *, inner_batch=False, f_diagonal=False):
    """log p(xs) (base.py:100-134)."""
    lg = LGSSM(*lgssm)
    return _logpdf(_lib.LOGPDF_PRIOR, None, xs, None, lg._replace(Hs=None, Rs=None, cs=None), inner_batch, f_diagonal)


def log_likelihood(ys, xs, lgssm, *, inner_batch=False):
    """log p(ys | xs) (base.py:138-166)."""
    return _logpdf(_lib.LOGPDF_LOGLIK, ys, xs, None, lgssm, inner_batch)


def auxiliary_draw(x, normals, delta):
    """u = x + sqrt(delta / 2) * normals (generic.py:61); x is ``[C, ...]``, delta scalar or ``[C]``."""
    dev = _lib.require_cuda(x, normals, delta if torch.is_tensor(delta) and delta.dim() > 0 else None)
    sfx = _lib.suffix(x.dtype)
    x = x.contiguous()
    normals = normals.contiguous()
    C = x.shape[0]
    dl = _as_delta(delta, C, x.dtype, dev)
    u = torch.empty_like(x)
    _lib.call("auxssm_auxiliary_draw_" + sfx,
        _lib.stream_ptr(dev), C, x.numel() // C, _lib.ptr(x), _lib.ptr(normals), _lib.ptr(dl),
        1 if dl.numel() > 1 else 0, _lib.ptr(u))
    return u


_DELTA_CACHE = {}


def _as_delta(delta, C, dtype, dev):
    if not torch.is_tensor(delta):
        # A host scalar: ``torch.as_tensor(0.005, device=cuda)`` is a BLOCKING pageable copy -- it waits for everything
        # queued on the stream (measured: 6 hidden synchronisations per sweep, ~1 ms of host time each).  The device image
        # of a scalar step size is made once by an asynchronous fill and reused.
        key = (float(delta), dtype, dev)
        dl = _DELTA_CACHE.get(key)
        if dl is None and torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture the fill belongs to the graph (and its memory to the graph's pool): not cached
            return torch.full((1,), float(delta), dtype=dtype, device=dev)
        if dl is None:
            if len(_DELTA_CACHE) >= 256:
                _DELTA_CACHE.clear()
            dl = _DELTA_CACHE[key] = torch.full((1,), float(delta), dtype=dtype, device=dev)
        return dl
    dl = torch.as_tensor(delta, dtype=dtype, device=dev).reshape(-1)
    if dl.numel() not in (1, C):
        raise AuxssmError("delta must be a scalar or have one entry per chain")
    return dl.contiguous()


def mh_accept(x, x_prop, u, delta, log_lgssm_prop, log_lgssm_rev, log_target_prop, log_target_rev, uniforms):
    """MH accept/reject of generic.py:70-74 / 98-106.  Returns (x_new, accepted[bool], alpha)."""
    dev = _lib.require_cuda(x, x_prop, u, log_lgssm_prop, log_lgssm_rev, log_target_prop, log_target_rev, uniforms)
    sfx = _lib.suffix(x.dtype)
    C = x.shape[0]
    dl = _as_delta(delta, C, x.dtype, dev)
    x_new = x.clone(memory_format=torch.contiguous_format)
    accepted = torch.empty((C,), dtype=torch.int32, device=dev)
    alpha = torch.empty((C,), dtype=x.dtype, device=dev)
    vec = lambda t: t.reshape(-1).to(x.dtype).contiguous()
    args = [vec(log_lgssm_prop), vec(log_lgssm_rev), vec(log_target_prop), vec(log_target_rev)]
    un = vec(uniforms)
    uc, xp = u.contiguous(), x_prop.contiguous()
    _lib.call("auxssm_kalman_mh_" + sfx,
        _lib.stream_ptr(dev), C, x.numel() // C, _lib.ptr(dl), 1 if dl.numel() > 1 else 0,
        _lib.ptr(args[0]), _lib.ptr(args[1]), _lib.ptr(args[2]), _lib.ptr(args[3]), _lib.ptr(uc), _lib.ptr(xp),
        _lib.ptr(un), _lib.ptr(x_new), _lib.ptr(accepted), _lib.ptr(alpha))
    return x_new, accepted.bool(), alpha


# --------------------------------------------------------------------------------------------
# kernel constructor
# --------------------------------------------------------------------------------------------

def get_kernel(dynamics_factory, observations_factory, log_likelihood_fn, parallel, *, time_chunks=0,
               f_diagonal=False, prior_cancels=False):
    """Auxiliary Kalman sampler: returns ``(init, kernel)`` exactly like the reference
    (kalman/generic.py:19-95; note the reference's docstring lists them the other way round).

    ``dynamics_factory(x) -> (m0, P0, Fs, Qs, bs)``, ``observations_factory(x, u, delta) -> (ys, Hs, Rs, cs)``
    and ``log_likelihood_fn(x) -> [C]`` operate on CUDA tensors with a leading chain axis.
    ``kernel(key, state, delta) -> KalmanSampler``; ``state.x`` is ``[C, T, d]`` or ``[T, d]``.

    ``prior_cancels=True`` is for adapters whose target is "the proposal's own dynamics x a potential" (the SV and
    latent-Gaussian models): the acceptance ratio of generic.py:98-106 contains ``log p(x)`` under the dynamics once
    inside ``log_target`` and once, with the opposite sign, inside ``posterior_logpdf`` -- for the proposal and for the
    reverse move -- so the four evaluations cancel analytically.  With the flag the proposal densities are evaluated
    as ``log p(y | x) - ell`` (base.py:138-166 minus ``ell``) and ``log_likelihood_fn`` must return the target WITHOUT
    its ``log p(x)`` term; the acceptance probability is the same up to the rounding of the cancelled terms (|log p(x)|
    is ~1e4 here, so dropping them makes ``alpha`` more, not less, accurate in fp32).  Default: the reference's form.
    """

    def do_one(sampling_key, u, delta, x, x_prop=None):
        m0, P0, Fs, Qs, bs, *_ = dynamics_factory(x)
        ys, Hs, Rs, cs, *_ = observations_factory(x, u, delta)
        lgssm = LGSSM(m0, P0, Fs, Qs, bs, Hs, Rs, cs)
        # the reverse proposal (generic.py:67) only needs the likelihood: ell-only filter call, nothing stored
        ms, Ps, ell = filtering(ys, lgssm, parallel, time_chunks=time_chunks, f_diagonal=f_diagonal,
                                marginals=x_prop is None)
        if x_prop is None:
            x_prop = sampling(sampling_key, ms, Ps, lgssm, parallel, time_chunks=time_chunks,
                              f_diagonal=f_diagonal)
        if prior_cancels:
            log_lgssm = _logpdf(_lib.LOGPDF_LOGLIK | _lib.LOGPDF_MINUS_ELL, ys, x_prop, ell, lgssm, False, f_diagonal)
        else:
            log_lgssm = posterior_logpdf(ys, x_prop, ell, lgssm, f_diagonal=f_diagonal)
        return log_lgssm, log_likelihood_fn(x_prop), x_prop

    def kernel(key, state, delta):
        x = state.x
        squeeze = x.dim() == 2
        if squeeze:
            x = x.unsqueeze(0)
        dev, dtype = x.device, x.dtype
        normals_aux = key.normal(tuple(x.shape), dtype, dev)         # generic.py:61
        u = auxiliary_draw(x, normals_aux, delta)
        lp_prop, lt_prop, x_prop = do_one(key, u, delta, x)           # generic.py:64
        lp_rev, lt_rev, _ = do_one(None, u, delta, x_prop, x)         # generic.py:67
        uniforms = key.uniform((x.shape[0],), dtype, dev)             # generic.py:73
        x_new, accepted, _ = mh_accept(x, x_prop, u, delta, lp_prop, lp_rev, lt_prop, lt_rev, uniforms)
        if squeeze:
            return KalmanSampler(x=x_new[0], updated=accepted[0])
        return KalmanSampler(x=x_new, updated=accepted)

    def init(x):
        return KalmanSampler(x=x, updated=True)

    return init, kernel

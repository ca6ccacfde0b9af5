"""Outer MCMC loop on the device.

Mirrors the pieces of the reference that surround a sweep:

* ``delta_adaptation(delta, target_rate, acceptance_rate, adaptation_rate, min_delta, max_delta)``
  (aux_samplers/common.py:4-32), computed by ``auxssm_delta_adaptation_*`` on device tensors;
* the body of ``loop`` in examples/stochastic_volatility/experiment.py:98-127 -- sweep, windowed / overall
  acceptance, expected squared jump distance, step-size update -- as ``DeviceLoop``: every quantity stays in device
  memory, the step size is a device tensor fed back into the next sweep, and ``n_inner`` iterations are captured
  once in a CUDA graph and replayed, so the host issues ONE launch per ``n_inner`` sweeps instead of ~15 per sweep
  (the few-chain regime of BASELINE configs[0] is launch-bound otherwise).  Randomness inside the graph comes from
  ``draws.PhiloxDraws`` whose call counter lives in device memory.
"""
import torch

from . import _lib
from ._lib import AuxssmError


def delta_adaptation(delta, target_rate, acceptance_rate, adaptation_rate, min_delta=1e-20, max_delta=1e20, out=None):
    """common.py:4-32 on device tensors: ``clip(delta * exp(adaptation_rate (acceptance_rate - target_rate)))``.
    ``delta`` is a CUDA tensor of any shape; ``acceptance_rate`` has the same number of entries or exactly one."""
    dev = _lib.require_cuda(delta, acceptance_rate)
    dl = delta.contiguous()
    acc = acceptance_rate.to(dl.dtype).contiguous()
    if acc.numel() not in (1, dl.numel()):
        raise AuxssmError("acceptance_rate must have one entry or as many as delta")
    res = torch.empty_like(dl) if out is None else out
    _lib.call("auxssm_delta_adaptation_" + _lib.suffix(dl.dtype), _lib.stream_ptr(dev), dl.numel(), _lib.ptr(dl),
              _lib.ptr(acc), 1 if acc.numel() > 1 else 0, float(target_rate), float(adaptation_rate), float(min_delta),
              float(max_delta), _lib.ptr(res))
    return res


def mcmc_stats(stats, x_old, x_new, updated, beta=0.9):
    """Running statistics of the outer loop (experiment.py:98-127), in place on ``stats [C, 4]`` =
    {n, acceptance mean, windowed acceptance, mean squared jump distance}; ``updated`` is bool/int ``[C]`` or ``[C, T]``."""
    dev = _lib.require_cuda(stats, x_old, x_new, updated)
    C = x_old.shape[0]
    upd = updated.reshape(C, -1).to(torch.int32).contiguous()
    xo, xn = x_old.contiguous(), x_new.contiguous()
    _lib.call("auxssm_mcmc_stats_" + _lib.suffix(xo.dtype), _lib.stream_ptr(dev), C, xo.numel() // C, upd.shape[1],
              _lib.ptr(xo), _lib.ptr(xn), _lib.ptr(upd), float(beta), _lib.ptr(stats))
    return stats


class DeviceLoop:
    """``n_inner`` sweeps + statistics + step-size adaptation per host launch (one captured CUDA graph).

    ``kernel(key, state, delta) -> state`` is any kernel of this package whose state carries ``x [C, T, d]`` and
    ``updated``; ``key`` must be graph-safe (``PhiloxDraws``).  ``delta`` is a device tensor ``[C]`` (one step size
    per chain, adapted from that chain's windowed acceptance) -- the reference adapts a scalar from a scalar
    acceptance rate; with C = 1 the two coincide.
    """

    def __init__(self, kernel, key, state, delta, *, target_rate=0.5, adaptation_rate=0.05, beta=0.9, min_delta=1e-20,
                 max_delta=1e20, n_inner=8, adapt=True, make_state=None, use_graph=True):
        x = state.x
        if x.dim() != 3:
            raise AuxssmError("DeviceLoop needs a chain-leading state x [C, T, d]")
        self.kernel, self.key, self.n_inner, self.adapt = kernel, key, int(n_inner), adapt
        self.target_rate, self.adaptation_rate, self.beta = target_rate, adaptation_rate, beta
        self.min_delta, self.max_delta = min_delta, max_delta
        self.make_state = make_state or (lambda x_, upd: type(state)(x_, upd))
        C = x.shape[0]
        dev = x.device
        self.x = x.clone(memory_format=torch.contiguous_format)           # static buffers the graph reads / writes
        self.updated = torch.ones((C,) + tuple(state.updated.shape[1:]) if torch.is_tensor(state.updated) else (C,),
                                  dtype=torch.bool, device=dev)
        self.delta = torch.as_tensor(delta, dtype=x.dtype, device=dev).reshape(-1).expand(C).contiguous().clone()
        self.stats = torch.zeros((C, 4), dtype=x.dtype, device=dev)
        self.graph = None
        self.launches_per_replay = None
        if use_graph:
            self._capture()

    def _body(self):
        for _ in range(self.n_inner):
            st = self.kernel(self.key, self.make_state(self.x, self.updated), self.delta)
            mcmc_stats(self.stats, self.x, st.x, st.updated, self.beta)
            if self.adapt:
                delta_adaptation(self.delta, self.target_rate, self.stats[:, 2], self.adaptation_rate, self.min_delta,
                                 self.max_delta, out=self.delta)
            self.x.copy_(st.x)
            self.updated.copy_(st.updated.reshape(self.updated.shape))

    def _capture(self):
        dev = self.x.device
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):   # warm-up outside capture: lazy module loads, cudaFuncSetAttribute, allocator
            snap = (self.x.clone(), self.delta.clone(), self.stats.clone(), self.key.counter.clone()
                    if hasattr(self.key, "counter") else None)
            self._body()
            self.x.copy_(snap[0]); self.delta.copy_(snap[1]); self.stats.copy_(snap[2])
            if snap[3] is not None:
                self.key.counter.copy_(snap[3])
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_replay = _lib.launch_count() - n0

    def run(self, n_outer=1):
        """Runs ``n_outer * n_inner`` sweeps; returns the current state (device tensors, no synchronisation)."""
        for _ in range(n_outer):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._body()
        return self.make_state(self.x, self.updated)


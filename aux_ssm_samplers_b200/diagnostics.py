"""Cross-chain MCMC diagnostics over chain-sharded ranks.

Chains are sharded across GPUs with no communication inside a sweep; the only collective in the
framework is the cheap exchange here: per-chain sufficient statistics of a few tracked scalars are
all-gathered (NCCL on GPUs, gloo in the CPU tests) and split-R-hat / ESS are formed from them.

* ``effective_sample_size`` follows the reference's ``examples/rare_event/ess.py:28-161`` (BlackJAX/Stan:
  FFT autocovariances averaged over chains, Geyer initial positive + monotone sequences, optional
  known variance).  Decomposition across ranks: every rank computes its chains' means and the SUM of
  its chains' autocovariance vectors; the sums are all-reduced and the per-chain means all-gathered --
  exactly the two quantities ``ess.py:80`` and ``:87-93`` need.
* ``split_rhat`` is not in the reference (BASELINE.json names it): Gelman et al. (BDA3) split-R-hat,
  from all-gathered per-half-chain (mean, variance).  Parity unpinned vs the reference.
"""
import math

import torch
import torch.distributed as dist


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def all_gather_chains(local):
    """Concatenate ``local`` ([C_local, ...]) over ranks along axis 0 (identity for one rank)."""
    if _world() == 1:
        return local
    local = local.contiguous()
    out = torch.empty((_world() * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


def _all_reduce_sum(t):
    if _world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def cross_chain_summary(tracked):
    """Mean / variance over ALL chains (all ranks) of tracked scalars ``[C_local, K]`` -> dict (for logs)."""
    allv = all_gather_chains(tracked.double())
    return {"n_chains": int(allv.shape[0]), "mean": [float(v) for v in allv.mean(0)],
            "var": [float(v) for v in allv.var(0, unbiased=True)], "world_size": _world()}


def split_rhat(samples):
    """Split-R-hat of ``samples`` [C_local, n] (this rank's chains); returns a python float.

    Each chain is split in two halves; with m half-chains of length h: W = mean of within variances,
    B/h = variance of half-chain means, R-hat = sqrt(((h-1)/h W + B/h) / W).
    """
    x = samples.double()
    n = x.shape[1]
    h = n // 2
    halves = torch.cat([x[:, :h], x[:, n - h:]], dim=0)  # [2 C_local, h]
    stats = torch.stack([halves.mean(1), halves.var(1, unbiased=True)], dim=1)  # [2 C_local, 2]
    stats = all_gather_chains(stats)
    means, variances = stats[:, 0], stats[:, 1]
    W = variances.mean()
    B_over_h = means.var(unbiased=True)
    var_plus = (h - 1.0) / h * W + B_over_h
    return float(torch.sqrt(var_plus / W))


def _next_fast_len(n):
    """Smallest 2^a 3^b 5^c >= n (what scipy.fftpack.next_fast_len returns for these sizes)."""
    best = None
    p5 = 1
    while p5 < 2 * n:
        p35 = p5
        while p35 < 2 * n:
            v = p35
            while v < n:
                v *= 2
            if best is None or v < best:
                best = v
            p35 *= 3
        p5 *= 5
    return best


def chain_autocovariances(samples):
    """Per-chain means ``[C]`` and biased autocovariances ``[C, n]`` (float64) of ``samples [C, n]``.
    CUDA tensors go through ``auxssm_autocov_*`` (direct sums in double); CPU tensors (the gloo tests) through an
    FFT, as ``ess.py:69-79`` does."""
    c_local, n = samples.shape
    if samples.is_cuda and samples.dtype in (torch.float32, torch.float64):
        from . import _lib
        x = samples.contiguous()
        mean = torch.empty((c_local,), dtype=torch.float64, device=x.device)
        acov = torch.empty((c_local, n), dtype=torch.float64, device=x.device)
        _lib.call("auxssm_autocov_" + _lib.suffix(x.dtype), _lib.stream_ptr(x.device), c_local, n, _lib.ptr(x),
                  _lib.ptr(mean), _lib.ptr(acov))
        return mean, acov
    x = samples.double()
    mean = x.mean(1)
    centered = x - mean[:, None]
    m = _next_fast_len(2 * n)
    f = torch.fft.rfft(centered, n=m, dim=1)
    return mean, torch.fft.irfft(f * torch.conj(f), n=m, dim=1)[:, :n] / n


def effective_sample_size(samples, var=None):
    """ESS of ``samples`` [C_local, n] over all ranks' chains, after ``ess.py:28-161``."""
    c_local, n = samples.shape
    dev = samples.device
    mean_chain, acov = chain_autocovariances(samples)
    acov_sum = _all_reduce_sum(acov.sum(0))                                 # sum over ALL chains
    n_chains = int(_all_reduce_sum(torch.tensor([float(c_local)], dtype=torch.float64, device=dev)).item())
    means = all_gather_chains(mean_chain).cpu().tolist()
    mean_acov = (acov_sum / n_chains).cpu().tolist()                        # ess.py:80
    # everything below is O(n) scalar work on the host (ess.py:81-161)
    mean_var0 = mean_acov[0] * n / (n - 1.0)
    weighted_var = mean_var0 * (n - 1.0) / n
    if n_chains > 1:                                                        # ess.py:87-93
        mbar = sum(means) / n_chains
        weighted_var = weighted_var + sum((v - mbar) ** 2 for v in means) / (n_chains - 1.0)
    if var is not None:
        weighted_var = mean_var0 = float(var)
    n_even = n - n % 2
    rho = [1.0] + [1.0 - (mean_var0 - mean_acov[t]) / weighted_var for t in range(1, n_even)]
    rho_even, rho_odd = rho[0::2], rho[1::2]
    k_pairs = len(rho_odd)
    rho_even = rho_even[:k_pairs] if len(rho_even) > k_pairs else rho_even
    # Geyer's initial positive sequence (ess.py:117-135)
    mask, carry, max_t = [], True, 0
    for t in range(k_pairs):
        carry = carry and (rho_even[t] + rho_odd[t] > 0.0)
        mask.append(carry)
        if carry:
            max_t = t
    odd = [rho_odd[t] if mask[t] else 0.0 for t in range(k_pairs)]
    mask_even = list(mask)
    # ess.py:131 `mask.at[max_t + 1].set(...)`: an out-of-bounds scatter is DROPPED by XLA, whereas the gathers
    # `rho_hat_even[indices]` (:131, :156) clamp the index
    if max_t + 1 < k_pairs:
        mask_even[max_t + 1] = rho_even[max_t + 1] > 0
    idx = min(max_t + 1, k_pairs - 1)
    even = [rho_even[t] if mask_even[t] else 0.0 for t in range(k_pairs)]
    # Geyer's initial monotone sequence (ess.py:137-149)
    prev = even[0] + odd[0]
    even_f, odd_f = list(even), list(odd)
    for t in range(k_pairs):
        cur = even[t] + odd[t]
        if cur > prev:
            even_f[t] = odd_f[t] = prev / 2.0
        else:
            prev = cur
    ess_raw = n_chains * n
    tau = -1.0 + 2.0 * sum(e + o for e, o in zip(even_f, odd_f)) - even_f[idx]
    tau = max(tau, 1.0 / math.log10(ess_raw))
    return ess_raw / tau

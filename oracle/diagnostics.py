"""Single-process NumPy statements of the cross-chain diagnostics (oracle only).

``effective_sample_size`` restates reference ``aux_samplers/examples/rare_event/ess.py:28-161`` on a
full ``[chains, samples]`` array; ``split_rhat`` is BDA3's split-R-hat (absent from the reference:
parity unpinned).
"""
import numpy as np
from scipy.fftpack import next_fast_len


def split_rhat(x):
    x = np.asarray(x, float)
    n = x.shape[1]
    h = n // 2
    halves = np.concatenate([x[:, :h], x[:, n - h:]], axis=0)
    W = halves.var(axis=1, ddof=1).mean()
    B_over_h = halves.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((h - 1.0) / h * W + B_over_h) / W))


def effective_sample_size(x, var=None):
    x = np.asarray(x, float)
    num_chains, num_samples = x.shape
    mean_across_chain = x.mean(axis=1, keepdims=True)
    centered = x - mean_across_chain
    m = next_fast_len(2 * num_samples)
    f = np.fft.rfft(centered, n=m, axis=1)
    autocov = np.fft.irfft(f * np.conjugate(f), n=m, axis=1)[:, :num_samples] / num_samples
    mean_autocov = autocov.mean(0)
    mean_var0 = mean_autocov[0] * num_samples / (num_samples - 1.0)
    weighted_var = mean_var0 * (num_samples - 1.0) / num_samples
    if num_chains > 1:
        weighted_var = weighted_var + mean_across_chain[:, 0].var(ddof=1)
    if var is not None:
        weighted_var = mean_var0 = float(var)
    n_even = num_samples - num_samples % 2
    rho_hat = np.concatenate([[1.0], 1.0 - (mean_var0 - mean_autocov[1:n_even]) / weighted_var])
    rho_even, rho_odd = rho_hat[0::2], rho_hat[1::2]
    k = len(rho_odd)
    rho_even = rho_even[:k]
    mask0 = (rho_even + rho_odd) > 0.0
    mask = np.logical_and.accumulate(mask0)
    max_t = int(np.max(np.flatnonzero(mask))) if mask.any() else 0
    # ess.py:126-131: `mask.at[max_t + 1].set(rho_even[max_t + 1] > 0)`; when max_t + 1 == k the scatter is OUT OF
    # BOUNDS and XLA drops it, whereas the two gathers (`rho_hat_even[indices]`, ess.py:131 and :156) clamp.
    idx = min(max_t + 1, k - 1)
    rho_odd = np.where(mask, rho_odd, 0.0)
    mask_even = mask.copy()
    if max_t + 1 < k:
        mask_even[max_t + 1] = rho_even[max_t + 1] > 0
    rho_even = np.where(mask_even, rho_even, 0.0)
    s = rho_even + rho_odd
    even_f, odd_f = rho_even.copy(), rho_odd.copy()
    prev = s[0]
    for t in range(k):
        if s[t] > prev:
            even_f[t] = odd_f[t] = prev / 2.0
        else:
            prev = s[t]
    ess_raw = num_chains * num_samples
    tau = -1.0 + 2.0 * np.sum(even_f + odd_f) - even_f[idx]
    tau = max(tau, 1.0 / np.log10(ess_raw))
    return float(ess_raw / tau)


def delta_adaptation(delta, target_rate, acceptance_rate, adaptation_rate, min_delta=1e-20, max_delta=1e20):
    """Reference ``aux_samplers/common.py:4-32``: delta * exp(rate (acc - target)), clipped."""
    rate = np.exp(adaptation_rate * (np.asarray(acceptance_rate) - target_rate))
    return np.clip(np.asarray(delta) * rate, min_delta, max_delta)

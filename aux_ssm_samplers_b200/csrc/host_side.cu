// Small kernels that keep the sampler's host side free of eager tensor arithmetic: the analytic gradient of the
// auxiliary target (replaces jax.grad(_log_pdf), csmc/independent.py:62,121-137), batched Cholesky factors of the
// model covariances (the jnp.linalg.cholesky calls of the model constructors, e.g. auxiliary_csmc.py:18-19), the
// step-size adaptation rule (common.py:4-32), running MCMC statistics of the outer loop
// (examples/stochastic_volatility/experiment.py:98-127) and counter-based Philox4x32-10 draws (stand-in for
// jax.random.normal / uniform: generated on the device, reproducible from (seed, counter), restated bit for bit in
// oracle/philox.py).
#include "common.cuh"
#include "philox.cuh"

namespace auxssm {

// ---- gradient of log M0(u_0) + G0(u_0) + sum_t [Gt(u_t) + log Mt(u_t | u_{t-1})] at u -------------------------
// One thread per (chain, t).  With r_0 = P0^{-1}(u_0 - m0), r_t = Q^{-1}(u_t - F u_{t-1} - b):
//   grad_t = -r_t + F^T r_{t+1} + d/du potential(u_t).
template <typename T>
__device__ void cho_solve_local(const T* __restrict__ L, T* v, int d) {  // v <- (L L^T)^{-1} v, L row-major [d, d]
  for (int i = 0; i < d; ++i) {
    T acc = v[i];
    for (int k = 0; k < i; ++k) acc -= L[i * d + k] * v[k];
    v[i] = acc / L[i * d + i];
  }
  for (int i = d - 1; i >= 0; --i) {
    T acc = v[i];
    for (int k = i + 1; k < d; ++k) acc -= L[k * d + i] * v[k];
    v[i] = acc / L[i * d + i];
  }
}

template <typename T>
__global__ void grad_log_target_kernel(int n_chains, int Tn, int d, int potential, T sig, int64_t ys_chain_stride,
                                       const T* __restrict__ u, const T* __restrict__ m0, const T* __restrict__ Lp0,
                                       const T* __restrict__ F, const T* __restrict__ b, const T* __restrict__ Lq,
                                       const T* __restrict__ ys, T* __restrict__ grad) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)n_chains * Tn) return;
  const int c = (int)(gid / Tn), t = (int)(gid % Tn);
  const T* uc = u + (int64_t)c * Tn * d;
  const T* ut = uc + (int64_t)t * d;
  T r[AUXSSM_MAX_DIM], rn[AUXSSM_MAX_DIM];
  if (t == 0) {
    for (int i = 0; i < d; ++i) r[i] = ut[i] - m0[i];
    cho_solve_local<T>(Lp0, r, d);
  } else {
    const T* up = ut - d;
    for (int i = 0; i < d; ++i) {
      T acc = b[i];
      for (int k = 0; k < d; ++k) acc += F[i * d + k] * up[k];
      r[i] = ut[i] - acc;
    }
    cho_solve_local<T>(Lq, r, d);
  }
  const bool has_next = t + 1 < Tn;
  if (has_next) {
    const T* un = ut + d;
    for (int i = 0; i < d; ++i) {
      T acc = b[i];
      for (int k = 0; k < d; ++k) acc += F[i * d + k] * ut[k];
      rn[i] = un[i] - acc;
    }
    cho_solve_local<T>(Lq, rn, d);
  }
  const T* yt = ys ? ys + (int64_t)c * ys_chain_stride + (int64_t)t * d : nullptr;
  T* g = grad + gid * d;
  for (int i = 0; i < d; ++i) {
    T v = -r[i];
    if (has_next) {
      T acc = T(0);
      for (int k = 0; k < d; ++k) acc += F[k * d + i] * rn[k];
      v += acc;
    }
    if (potential == AUXSSM_POT_GAUSSIAN) v += (yt[i] - ut[i]) / (sig * sig);
    if (potential == AUXSSM_POT_SV) v += T(0.5) * yt[i] * yt[i] * exp(-ut[i]) - T(0.5);
    g[i] = v;
  }
}

template <typename T>
static int grad_entry(void* stream, const auxssm_csmc_desc* dsc, const T* u, const T* m0, const T* Lp0, const T* F,
                      const T* b, const T* Lq, const T* ys, T* grad) {
  const char* who = "auxssm_csmc_grad_log_target";
  if (!dsc || !u || !m0 || !Lp0 || !F || !b || !Lq || !grad || dsc->T < 1 || dsc->d < 1 || dsc->n_chains < 1 ||
      dsc->d > AUXSSM_MAX_DIM || (dsc->potential != AUXSSM_POT_FLAT && !ys)) {
    set_error("%s: invalid argument", who);
    return AUXSSM_ERR_INVALID;
  }
  const int64_t n = (int64_t)dsc->n_chains * dsc->T;
  grad_log_target_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      dsc->n_chains, dsc->T, dsc->d, dsc->potential, T(dsc->sig), dsc->ys_chain_stride, u, m0, Lp0, F, b, Lq, ys, grad);
  AUXSSM_CHECK_LAUNCH(who);
  return AUXSSM_OK;
}

// ---- batched lower Cholesky, one thread per matrix; jnp.linalg.cholesky semantics (all-NaN factor on failure);
// d == 1 is the scalar sqrt branch of the reference ---------------------------------------------------------------
template <typename T> __global__ void cholesky_kernel(int64_t n, int d, const T* __restrict__ A, T* __restrict__ L) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const T* a = A + g * d * d;
  T* l = L + g * d * d;
  bool bad = false;
  for (int j = 0; j < d; ++j) {
    T s = a[j * d + j];
    for (int k = 0; k < j; ++k) s -= l[j * d + k] * l[j * d + k];
    const T ljj = sqrt(s);
    if (!(s > T(0)) && d > 1) bad = true;
    l[j * d + j] = ljj;
    for (int i = j + 1; i < d; ++i) {
      T v = a[i * d + j];
      for (int k = 0; k < j; ++k) v -= l[i * d + k] * l[j * d + k];
      l[i * d + j] = v / ljj;
      l[j * d + i] = T(0);
    }
  }
  if (bad)
    for (int e = 0; e < d * d; ++e) l[e] = T(NAN);
}

template <typename T> static int cholesky_entry(void* stream, int64_t n, int32_t d, const T* A, T* L) {
  const char* who = "auxssm_cholesky";
  if (n < 1 || d < 1 || d > AUXSSM_MAX_DIM || !A || !L) { set_error("%s: invalid argument", who); return AUXSSM_ERR_INVALID; }
  cholesky_kernel<T><<<(unsigned)((n + 63) / 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(n, d, A, L);
  AUXSSM_CHECK_LAUNCH(who);
  return AUXSSM_OK;
}

// ---- delta adaptation (common.py:4-32): delta * exp(rate (acc - target)), clipped -----------------------------------
template <typename T>
__global__ void delta_adaptation_kernel(int64_t n, const T* __restrict__ delta, const T* __restrict__ acc, int acc_stride,
                                        T target, T rate, T lo, T hi, T* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  T v = delta[g] * exp(rate * (acc[g * acc_stride] - target));
  v = v < lo ? lo : v;
  v = v > hi ? hi : v;
  out[g] = v;
}

template <typename T>
static int delta_entry(void* stream, int64_t n, const T* delta, const T* acc, int32_t acc_stride, double target,
                       double rate, double lo, double hi, T* out) {
  const char* who = "auxssm_delta_adaptation";
  if (n < 1 || !delta || !acc || !out) { set_error("%s: invalid argument", who); return AUXSSM_ERR_INVALID; }
  delta_adaptation_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      n, delta, acc, acc_stride, T(target), T(rate), T(lo), T(hi), out);
  AUXSSM_CHECK_LAUNCH(who);
  return AUXSSM_OK;
}

// ---- running statistics of the outer MCMC loop (experiment.py:98-127) -----------------------------------------------
// One thread per chain; `updated` is the int32 accept flag of the sweep (or, with per_step > 1, per-time-step
// flags whose mean is used).  stats[c] = {n, running acceptance mean, windowed acceptance (EMA), ESJD mean}.
template <typename T>
__global__ void mcmc_stats_kernel(int n_chains, int64_t n_per_chain, int per_step, const T* __restrict__ x_old,
                                  const T* __restrict__ x_new, const int32_t* __restrict__ updated, T beta,
                                  T* __restrict__ stats) {
  const int c = blockIdx.x;
  __shared__ T red[32];
  T s = T(0);
  const T* a = x_old + (int64_t)c * n_per_chain;
  const T* b = x_new + (int64_t)c * n_per_chain;
  for (int64_t e = threadIdx.x; e < n_per_chain; e += blockDim.x) {
    const T dlt = b[e] - a[e];
    s += dlt * dlt;
  }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    T tot = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    T acc = T(0);
    for (int k = 0; k < per_step; ++k) acc += T(updated[(int64_t)c * per_step + k] != 0 ? 1 : 0);
    acc /= T(per_step);
    T* st = stats + (int64_t)c * 4;
    const T n = st[0] + T(1);
    st[0] = n;
    st[1] += (acc - st[1]) / n;                 // overall acceptance
    st[2] = (n == T(1)) ? acc : beta * st[2] + (T(1) - beta) * acc;  // windowed acceptance
    st[3] += (tot - st[3]) / n;                 // expected squared jump distance
  }
}

template <typename T>
static int stats_entry(void* stream, int32_t n_chains, int64_t n_per_chain, int32_t per_step, const T* x_old,
                       const T* x_new, const int32_t* updated, double beta, T* stats) {
  const char* who = "auxssm_mcmc_stats";
  if (n_chains < 1 || n_per_chain < 1 || per_step < 1 || !x_old || !x_new || !updated || !stats) {
    set_error("%s: invalid argument", who);
    return AUXSSM_ERR_INVALID;
  }
  mcmc_stats_kernel<T><<<n_chains, 256, 0, static_cast<cudaStream_t>(stream)>>>(n_chains, n_per_chain, per_step, x_old,
                                                                                x_new, updated, T(beta), stats);
  AUXSSM_CHECK_LAUNCH(who);
  return AUXSSM_OK;
}

// ---- Philox4x32-10 fills (layout and device functions: philox.cuh) ----------------------------------------------------
template <typename T>
__global__ void philox_fill_kernel(int mode, uint64_t seed, const uint64_t* __restrict__ ctr, int64_t n, T* __restrict__ out) {
  const uint64_t call = ctr ? ctr[0] : 0;
  constexpr int PER = PhiloxPer<T>::N;
  const int64_t nblk = (n + PER - 1) / PER;
  for (int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; blk < nblk; blk += (int64_t)gridDim.x * blockDim.x) {
    T v[PER];
    philox_block(mode, seed, call, (uint64_t)blk, v);
#pragma unroll
    for (int q = 0; q < PER; ++q)
      if (blk * PER + q < n) out[blk * PER + q] = v[q];
  }
}

__global__ void philox_bump_kernel(uint64_t* ctr, uint64_t by) { ctr[0] += by; }

template <typename T> static int philox_entry(void* stream, int mode, uint64_t seed, uint64_t* ctr, int64_t n, T* out) {
  const char* who = "auxssm_philox_fill";
  if (n < 1 || !out) { set_error("%s: invalid argument", who); return AUXSSM_ERR_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (n / ((sizeof(T) == 4) ? 4 : 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  philox_fill_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(mode, seed, ctr, n, out);
  AUXSSM_CHECK_LAUNCH(who);
  if (ctr) {  // the draw-call counter lives in device memory, so a captured CUDA graph draws fresh numbers on replay
    philox_bump_kernel<<<1, 1, 0, st>>>(ctr, 1);
    AUXSSM_CHECK_LAUNCH(who);
  }
  return AUXSSM_OK;
}

// ---- per-chain mean and biased autocovariances (ess.py:69-79 computes them with an FFT; here the direct sums) ------
// acov[c, k] = (1 / n) sum_{t < n - k} (x[c, t] - mean_c)(x[c, t + k] - mean_c); one CTA per chain, lags strided
// over threads, accumulation in double.
template <typename T>
__global__ void autocov_kernel(int64_t n, const T* __restrict__ x, double* __restrict__ mean, double* __restrict__ acov) {
  extern __shared__ double xc[];
  __shared__ double red[32];
  const int c = blockIdx.x;
  const T* xr = x + (int64_t)c * n;
  double s = 0.0;
  for (int64_t t = threadIdx.x; t < n; t += blockDim.x) s += (double)xr[t];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const double mu = tot / (double)n;
  for (int64_t t = threadIdx.x; t < n; t += blockDim.x) xc[t] = (double)xr[t] - mu;
  __syncthreads();
  if (threadIdx.x == 0) mean[c] = mu;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
    double a0 = 0.0, a1 = 0.0;
    int64_t t = 0;
    for (; t + 1 < n - k; t += 2) { a0 += xc[t] * xc[t + k]; a1 += xc[t + 1] * xc[t + 1 + k]; }
    if (t < n - k) a0 += xc[t] * xc[t + k];
    acov[(int64_t)c * n + k] = (a0 + a1) / (double)n;
  }
}

template <typename T> static int autocov_entry(void* stream, int32_t n_chains, int64_t n, const T* x, double* mean, double* acov) {
  const char* who = "auxssm_autocov";
  if (n_chains < 1 || n < 2 || !x || !mean || !acov) { set_error("%s: invalid argument", who); return AUXSSM_ERR_INVALID; }
  const size_t smem = (size_t)n * sizeof(double);
  int rc = ensure_smem(autocov_kernel<T>, smem, who);
  if (rc) return rc;
  autocov_kernel<T><<<n_chains, 256, smem, static_cast<cudaStream_t>(stream)>>>(n, x, mean, acov);
  AUXSSM_CHECK_LAUNCH(who);
  return AUXSSM_OK;
}

static int bump_entry(void* stream, uint64_t* ctr, uint64_t by) {
  if (!ctr) { set_error("auxssm_philox_advance: null counter"); return AUXSSM_ERR_INVALID; }
  philox_bump_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(ctr, by);
  AUXSSM_CHECK_LAUNCH("auxssm_philox_advance");
  return AUXSSM_OK;
}

}  // namespace auxssm

#define AUXSSM_EXPORT extern "C" __attribute__((visibility("default")))
#define AUXSSM_HOST_SIDE(SFX, TY)                                                                                       \
  AUXSSM_EXPORT int32_t auxssm_csmc_grad_log_target_##SFX(void* stream, const auxssm_csmc_desc* desc, const TY* u,       \
                                                          const TY* m0, const TY* chol_P0, const TY* F, const TY* b,     \
                                                          const TY* chol_Q, const TY* ys, TY* grad) {                    \
    return auxssm::grad_entry<TY>(stream, desc, u, m0, chol_P0, F, b, chol_Q, ys, grad);                                 \
  }                                                                                                                     \
  AUXSSM_EXPORT int32_t auxssm_cholesky_##SFX(void* stream, int64_t n, int32_t d, const TY* A, TY* L) {                  \
    return auxssm::cholesky_entry<TY>(stream, n, d, A, L);                                                               \
  }                                                                                                                     \
  AUXSSM_EXPORT int32_t auxssm_delta_adaptation_##SFX(void* stream, int64_t n, const TY* delta, const TY* acc_rate,      \
                                                      int32_t acc_stride, double target_rate, double adaptation_rate,   \
                                                      double min_delta, double max_delta, TY* delta_out) {              \
    return auxssm::delta_entry<TY>(stream, n, delta, acc_rate, acc_stride, target_rate, adaptation_rate, min_delta,      \
                                   max_delta, delta_out);                                                               \
  }                                                                                                                     \
  AUXSSM_EXPORT int32_t auxssm_mcmc_stats_##SFX(void* stream, int32_t n_chains, int64_t n_per_chain, int32_t per_step,   \
                                                const TY* x_old, const TY* x_new, const int32_t* updated, double beta,  \
                                                TY* stats) {                                                            \
    return auxssm::stats_entry<TY>(stream, n_chains, n_per_chain, per_step, x_old, x_new, updated, beta, stats);         \
  }                                                                                                                     \
  AUXSSM_EXPORT int32_t auxssm_autocov_##SFX(void* stream, int32_t n_chains, int64_t n, const TY* x, double* mean,        \
                                             double* acov) {                                                            \
    return auxssm::autocov_entry<TY>(stream, n_chains, n, x, mean, acov);                                                \
  }                                                                                                                     \
  AUXSSM_EXPORT int32_t auxssm_philox_uniform_##SFX(void* stream, uint64_t seed, uint64_t* counter, int64_t n, TY* out) { \
    return auxssm::philox_entry<TY>(stream, 0, seed, counter, n, out);                                                   \
  }                                                                                                                     \
  AUXSSM_EXPORT int32_t auxssm_philox_normal_##SFX(void* stream, uint64_t seed, uint64_t* counter, int64_t n, TY* out) {  \
    return auxssm::philox_entry<TY>(stream, 1, seed, counter, n, out);                                                   \
  }
AUXSSM_HOST_SIDE(f32, float)
AUXSSM_HOST_SIDE(f64, double)
AUXSSM_EXPORT int32_t auxssm_philox_advance(void* stream, uint64_t* counter, uint64_t by) {
  return auxssm::bump_entry(stream, counter, by);
}

// Parallel-in-time conditional SMC (Corenflos et al. 2022) for the auxiliary sampler with independent proposals.
//
// Replaces, for the closed model family of the fused sweep (include/auxssm.h):
//   aux_samplers/csmc/independent.py:78-118       _get_parallel_kernel (factory: Mt' = N(u_t + s_t^2 grad_t, s_t^2 I),
//                                                 Qt = N(u_t, s_t^2 I) when gradient, G0' = G0 + log M0,
//                                                 Gt' = log Mt(x_t | x_{t-1}) + Gt)
//   aux_samplers/_primitives/csmc/pit/csmc.py:68-114      _csmc (all proposals at once, particle 0 pinned, initial weights)
//   aux_samplers/_primitives/csmc/pit/operator.py:39-149  stitching operator (N x N weights, conditional multinomial
//                                                 resampling of N pairs out of N^2, one pair at the last level)
//   aux_samplers/_primitives/csmc/pit/dc_map.py:73-123    divide and conquer over T padded to a power of two
//
// Design.  The reference copies whole blocks of trajectories at every level (jnp.take over [2^k, N, d] arrays,
// O(T N d log T) traffic); here the particles xs[T, N, d] are written ONCE and a block is represented by the
// reference's own `origins` array (operator.py:97-99): origins[t, n] = the original particle trajectory n visits at
// time t, so x_t^n = xs[t, origins[t, n]] and a stitch only permutes int32 rows.  Every boundary (t - 1, t) is
// stitched exactly once over the log2(T) levels: (T - 1) N^2 weights per chain in total.
//   pit_init_kernel    one CTA per (time step, chain): proposals, initial log-weights, per-step log-sum-exp
//   pit_stitch_kernel  one CTA of 256 threads per (pair of blocks, chain) and level.  The transition density is
//                      evaluated as -|a_j - c_i|^2 / 2 with a_j = Lq^{-1} x_t^j and c_i = Lq^{-1}(F x_{t-1}^i + b)
//                      (O(N d^2) substitutions + O(N^2 d) differences instead of N^2 triangular solves); the N^2
//                      weights live in shared memory; log-sum-exp and CDF use the library's canonical 32-wide order
//                      (xor butterfly / Hillis-Steele scan, recursively over group totals: oracle.csmc.canonical_*),
//                      the N inverse-CDF searches are binary searches over the N^2 shared CDF.
#include "common.cuh"
#include "philox.cuh"

namespace auxssm {

constexpr int kPitThreads = 256;
constexpr int kPitMaxN = 128;
constexpr int kPitMaxD = 32;

template <typename T> struct PitArgs {
  auxssm_csmc_desc d;
  const T *x_star, *m0, *Lp0, *F, *b, *Lq, *ys, *u, *shd, *grad, *normals, *uniforms;
  T* x_out;
  int32_t* anc;
  T* xs;         // [C, T, N, d]
  T* lw;         // [C, T, N]
  int32_t* org;  // [C, T, N]
  RngCtl rng;
  const uint64_t* rng_counter;
  __device__ __forceinline__ uint64_t call0() const { return rng_counter ? rng_counter[0] : rng.call0; }
};

template <typename T> __device__ __forceinline__ T pit_half_log_2pi() { return T(0.91893853320467274178); }

// log potential of one particle (runtime d): FLAT / GAUSSIAN / SV, same formulas as csmc_kernel.cuh:potential
template <typename T> __device__ __forceinline__ T pit_potential(const auxssm_csmc_desc& dsc, const T* y, const T* x) {
  if (dsc.potential == AUXSSM_POT_FLAT) return T(0);
  T acc = T(0);
  if (dsc.potential == AUXSSM_POT_GAUSSIAN) {
    const T sig = T(dsc.sig), ls = log(sig);
    for (int i = 0; i < dsc.d; ++i) {
      const T z = (y[i] - x[i]) / sig;
      acc += T(-0.5) * z * z - ls - pit_half_log_2pi<T>();
    }
  } else {
    for (int i = 0; i < dsc.d; ++i) acc += T(-0.5) * y[i] * y[i] * exp(-x[i]) - T(0.5) * x[i] - pit_half_log_2pi<T>();
  }
  return acc;
}

template <typename T> __device__ __forceinline__ T pit_iso_logpdf(const T* x, const T* m, T s, int d) {
  T acc = T(0);
  const T ls = log(s);
  for (int i = 0; i < d; ++i) {
    const T z = (x[i] - m[i]) / s;
    acc += T(-0.5) * z * z - ls - pit_half_log_2pi<T>();
  }
  return acc;
}

// r <- L^{-1} r (L lower triangular [d, d] row-major)
template <typename T> __device__ __forceinline__ void pit_fwd_subst(const T* L, T* r, int d) {
  for (int i = 0; i < d; ++i) {
    T acc = r[i];
    for (int k = 0; k < i; ++k) acc -= L[i * d + k] * r[k];
    r[i] = acc / L[i * d + i];
  }
}

// ---- canonical reductions over an array in shared memory (oracle.csmc.canonical_sum / canonical_cdf) -----------

template <typename T> __device__ __forceinline__ T pit_block_max(T v, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  T m = red[0];
  for (int q = 1; q < kPitThreads / 32; ++q) m = fmax(m, red[q]);
  return m;
}

// sum of f(e), e < n, in the canonical order; `t1`, `t2` hold the group totals of the two upper levels
template <typename T, typename F> __device__ __forceinline__ T pit_canonical_sum(int n, F f, T* t1, T* t2) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = kPitThreads / 32;
  int g0 = (n + 31) / 32;
  for (int g = w; g < g0; g += nw) {
    const int e = g * 32 + lane;
    T v = (e < n) ? f(e) : T(0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) t1[g] = v;
  }
  __syncthreads();
  T* src = t1;
  T* dst = t2;
  int cur = g0;
  while (cur > 1) {
    const int groups = (cur + 31) / 32;
    for (int g = w; g < groups; g += nw) {
      const int e = g * 32 + lane;
      T v = (e < cur) ? src[e] : T(0);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) dst[g] = v;
    }
    __syncthreads();
    T* sw = src; src = dst; dst = sw;
    cur = groups;
  }
  const T out = src[0];
  __syncthreads();
  return out;
}

// in-place Hillis-Steele scan of every group of 32 entries of a[0..n); group totals (lane 31) go to tot[g]
template <typename T> __device__ __forceinline__ void pit_scan_groups(T* a, int n, T* tot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = kPitThreads / 32;
  const int groups = (n + 31) / 32;
  for (int g = w; g < groups; g += nw) {
    const int e = g * 32 + lane;
    T v = (e < n) ? a[e] : T(0);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const T o = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v = v + o;
    }
    if (e < n) a[e] = v;
    if (lane == 31 && tot != nullptr) tot[g] = v;
  }
  __syncthreads();
}

// a[0..n) <- canonical inclusive CDF (n <= 32^3); t1 >= ceil(n/32) words, t2 >= ceil(n/1024) words
template <typename T> __device__ __forceinline__ void pit_canonical_cdf(T* a, int n, T* t1, T* t2) {
  const int g1 = (n + 31) / 32;
  pit_scan_groups<T>(a, n, t1);
  if (g1 > 1) {
    const int g2 = (g1 + 31) / 32;
    pit_scan_groups<T>(t1, g1, t2);
    if (g2 > 1) {
      pit_scan_groups<T>(t2, g2, nullptr);  // g2 <= 32: one group
      for (int g = threadIdx.x; g < g1; g += kPitThreads)
        if (g >= 32) t1[g] = t1[g] + t2[g / 32 - 1];
      __syncthreads();
    }
    for (int e = threadIdx.x; e < n; e += kPitThreads)
      if (e >= 32) a[e] = a[e] + t1[e / 32 - 1];
    __syncthreads();
  }
}

template <typename T> __device__ __forceinline__ int pit_search_left(const T* cdf, T r, int n) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] < r) lo = mid + 1; else hi = mid;
  }
  return lo < n - 1 ? lo : n - 1;
}

// ---- pass 0: proposals + initial weights (pit/csmc.py:75-92) --------------------------------------------------
template <typename T> __global__ void __launch_bounds__(kPitMaxN) pit_init_kernel(PitArgs<T> a) {
  const auxssm_csmc_desc& dsc = a.d;
  const int t = blockIdx.x, c = blockIdx.y, n = threadIdx.x, N = dsc.N, d = dsc.d, Tn = dsc.T;
  __shared__ T red[8];
  __shared__ T sm[kPitMaxD], su[kPitMaxD];  // proposal mean, u_t
  const T s = a.shd[(int64_t)c * Tn + t];
  const T* ut = a.u + ((int64_t)c * Tn + t) * d;
  if (n < d) {
    su[n] = ut[n];
    sm[n] = dsc.gradient ? ut[n] + s * s * a.grad[((int64_t)c * Tn + t) * d + n] : ut[n];
  }
  __syncthreads();
  T x[kPitMaxD];
  T lw = -INFINITY;
  const bool act = n < N;
  if (act) {
    const int64_t e0 = (((int64_t)c * Tn + t) * N + n) * d;
    for (int i = 0; i < d; ++i) {
      const T eps = a.rng.on ? philox_elem<T>(PHILOX_NORMAL, a.rng.seed, a.call0(), (uint64_t)(e0 + i)) : a.normals[e0 + i];
      x[i] = (n == 0) ? a.x_star[((int64_t)c * Tn + t) * d + i] : sm[i] + s * eps;
      a.xs[e0 + i] = x[i];
    }
    lw = T(0);
    if (dsc.gradient) lw = pit_iso_logpdf<T>(x, su, s, d) - pit_iso_logpdf<T>(x, sm, s, d);  // Qt - Mt, :83-85
    if (t == 0) {                                                                           // G0' = G0 + log M0
      const T* y = a.ys ? a.ys + (int64_t)c * dsc.ys_chain_stride : nullptr;
      T r[kPitMaxD];
      T cst = -T(d) * pit_half_log_2pi<T>();
      for (int i = 0; i < d; ++i) { r[i] = x[i] - a.m0[i]; cst -= log(a.Lp0[i * d + i]); }
      pit_fwd_subst<T>(a.Lp0, r, d);
      T q = T(0);
      for (int i = 0; i < d; ++i) q += r[i] * r[i];
      lw = lw + (pit_potential<T>(dsc, y, x) + (T(-0.5) * q + cst));
    }
    a.org[((int64_t)c * Tn + t) * N + n] = n;
  }
  // per-step normalisation: log_wts -= logsumexp(log_wts, axis=1)   (canonical order, N <= 128: four warps)
  const int lane = n & 31, w = n >> 5, nw = blockDim.x >> 5;
  T m = lw;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (lane == 0) red[w] = m;
  __syncthreads();
  T amax = red[0];
  for (int q = 1; q < nw; ++q) amax = fmax(amax, red[q]);
  if (!isfinite(amax)) amax = T(0);
  __syncthreads();
  T v = act ? exp(lw - amax) : T(0);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) red[w] = v;
  __syncthreads();
  T tot = (lane < nw) ? red[lane] : T(0);
  if (nw > 1) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tot = tot + __shfl_xor_sync(0xffffffffu, tot, off);
  } else {
    tot = red[0];
  }
  const T lse = log(tot) + amax;
  if (act) a.lw[((int64_t)c * Tn + t) * N + n] = lw - lse;
}

// ---- one level of the divide and conquer (dc_map.py:110-121, operator.py:39-115) ------------------------------
template <typename T> struct PitSmem {
  T *W, *A, *Cc, *F, *Lq, *bq, *lwa, *lwb, *pot, *t1, *t2, *red;
  int *li, *ri;
};
template <typename T> __host__ __device__ inline size_t pit_smem_bytes(int N, int d) {
  const size_t n2 = ((size_t)N * N + 31) / 32 * 32;
  const size_t words = n2 + 2 * (size_t)N * (d + 1) + 2 * (size_t)d * d + d + 3 * (size_t)N + n2 / 32 + 32 + 32 + 16;
  return words * sizeof(T) + 2 * (size_t)N * sizeof(int) + 64;
}

template <typename T> __global__ void __launch_bounds__(kPitThreads) pit_stitch_kernel(PitArgs<T> a, int size, int last) {
  extern __shared__ __align__(16) unsigned char pit_raw[];
  const auxssm_csmc_desc& dsc = a.d;
  const int N = dsc.N, d = dsc.d, Tn = dsc.T, c = blockIdx.y, tid = threadIdx.x, DP = d + 1;
  const int a0 = blockIdx.x * 2 * size, t = a0 + size;
  if (t >= Tn) return;  // the right-hand block lies in the padding: pass-through (dc_map.py:101-107)
  const int b1 = (t + size < Tn) ? t + size : Tn;
  const int n2 = N * N, n2p = (n2 + 31) / 32 * 32;
  PitSmem<T> s;
  {
    T* p = reinterpret_cast<T*>(pit_raw);
    s.W = p; p += n2p;
    s.A = p; p += N * DP; s.Cc = p; p += N * DP;
    s.F = p; p += d * d; s.Lq = p; p += d * d; s.bq = p; p += d;
    s.lwa = p; p += N; s.lwb = p; p += N; s.pot = p; p += N;
    s.t1 = p; p += n2p / 32; s.t2 = p; p += 32; s.red = p; p += 32 + 16;
    s.li = reinterpret_cast<int*>(p); s.ri = s.li + N;
  }
  for (int e = tid; e < d * d; e += kPitThreads) { s.F[e] = a.F[e]; s.Lq[e] = a.Lq[e]; }
  if (tid < d) s.bq[tid] = a.b[tid];
  __syncthreads();
  const int64_t rowa = (int64_t)c * Tn + (t - 1), rowb = (int64_t)c * Tn + t;
  // a_j = Lq^{-1} x_t^j (threads 0..N-1),  c_i = Lq^{-1} (F x_{t-1}^i + b) (threads 128..128+N-1)
  if (tid < N) {
    const int j = tid, o = a.org[rowb * N + j];
    T x[kPitMaxD];
    for (int i = 0; i < d; ++i) x[i] = a.xs[(rowb * N + o) * d + i];
    const T* y = a.ys ? a.ys + (int64_t)c * dsc.ys_chain_stride + (int64_t)t * d : nullptr;
    s.pot[j] = pit_potential<T>(dsc, y, x);
    s.lwb[j] = a.lw[rowb * N + j];
    pit_fwd_subst<T>(s.Lq, x, d);
    for (int i = 0; i < d; ++i) s.A[j * DP + i] = x[i];
  } else if (tid >= kPitMaxN && tid < kPitMaxN + N) {
    const int i0 = tid - kPitMaxN, o = a.org[rowa * N + i0];
    T x[kPitMaxD], r[kPitMaxD];
    for (int i = 0; i < d; ++i) x[i] = a.xs[(rowa * N + o) * d + i];
    for (int i = 0; i < d; ++i) {
      T acc = s.bq[i];
      for (int k = 0; k < d; ++k) acc += s.F[i * d + k] * x[k];
      r[i] = acc;
    }
    pit_fwd_subst<T>(s.Lq, r, d);
    for (int i = 0; i < d; ++i) s.Cc[i0 * DP + i] = r[i];
    s.lwa[i0] = a.lw[rowa * N + i0];
  }
  __syncthreads();
  T cst = -T(d) * pit_half_log_2pi<T>();
  for (int i = 0; i < d; ++i) cst -= log(s.Lq[i * d + i]);
  // log-weights (operator.py:145-148): entry i * N + j pairs left trajectory i with right trajectory j
  T tmax = -INFINITY;
  for (int e = tid; e < n2p; e += kPitThreads) {
    T lw = -INFINITY;
    if (e < n2) {
      const int i = e / N, j = e - i * N;
      T q = T(0);
      for (int k = 0; k < d; ++k) {
        const T df = s.A[j * DP + k] - s.Cc[i * DP + k];
        q += df * df;
      }
      lw = ((T(-0.5) * q + cst) + s.pot[j]) + s.lwa[i] + s.lwb[j];
      tmax = fmax(tmax, lw);
    }
    s.W[e] = lw;
  }
  T amax = pit_block_max<T>(tmax, s.red);
  if (!isfinite(amax)) amax = T(0);
  const T* Wc = s.W;
  const T ssum = pit_canonical_sum<T>(n2, [&](int e) { return exp(Wc[e] - amax); }, s.t1, s.t2);
  const T lse = log(ssum) + amax;
  for (int e = tid; e < n2p; e += kPitThreads) s.W[e] = (e < n2) ? exp(s.W[e] - lse) : T(0);
  __syncthreads();
  pit_canonical_cdf<T>(s.W, n2, s.t1, s.t2);
  // N draws out of N^2 (multinomial, pair 0 pinned: resamplings.py:36) or one unpinned draw at the last level
  const T total = s.W[n2 - 1];
  const int ndraw = last ? 1 : N;
  if (tid < ndraw) {
    const int64_t e = rowb * N + tid;
    const T uu = a.rng.on ? philox_elem<T>(PHILOX_UNIFORM, a.rng.seed, a.call0() + 1, (uint64_t)e) : a.uniforms[e];
    int idx = pit_search_left<T>(s.W, total * (T(1) - uu), n2);
    if (!last && tid == 0) idx = 0;
    s.li[tid] = idx / N;
    s.ri[tid] = idx - (idx / N) * N;
  }
  __syncthreads();
  const int lane = tid & 31, w = tid >> 5, nw = kPitThreads / 32;
  if (last) {
    // single trajectory: x_out / ancestors directly (pit/csmc.py:113-114)
    for (int tt = a0 + tid; tt < b1; tt += kPitThreads) {
      const int64_t row = (int64_t)c * Tn + tt;
      const int o = a.org[row * N + (tt < t ? s.li[0] : s.ri[0])];
      a.anc[row] = o;
      for (int i = 0; i < d; ++i) a.x_out[row * d + i] = a.xs[(row * N + o) * d + i];
    }
    return;
  }
  // gather the origins of both halves (operator.py:97-99), one warp per time step; weights reset to 1 / N (:107-108)
  const T lwu = -log(T(N));
  for (int tt = a0 + w; tt < b1; tt += nw) {
    const int64_t row = (int64_t)c * Tn + tt;
    const int* pick = (tt < t) ? s.li : s.ri;
    int v[kPitMaxN / 32];
#pragma unroll
    for (int q = 0; q < kPitMaxN / 32; ++q) {
      const int n = lane + 32 * q;
      v[q] = (n < N) ? a.org[row * N + pick[n]] : 0;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < kPitMaxN / 32; ++q) {
      const int n = lane + 32 * q;
      if (n < N) { a.org[row * N + n] = v[q]; a.lw[row * N + n] = lwu; }
    }
  }
}

static size_t pit_align(size_t b) { return (b + 255) / 256 * 256; }

static int check_pit_desc(const auxssm_csmc_desc* d, const char* who) {
  if (d == nullptr) { set_error("%s: null descriptor", who); return AUXSSM_ERR_INVALID; }
  if (d->T < 2 || d->N < 2 || d->d < 1 || d->n_chains < 1) {
    set_error("%s: need T >= 2, N >= 2, d >= 1, n_chains >= 1 (got %d, %d, %d, %d)", who, d->T, d->N, d->d, d->n_chains);
    return AUXSSM_ERR_INVALID;
  }
  if (d->N > kPitMaxN || d->d > kPitMaxD) {
    set_error("%s: N <= %d and d <= %d (the N^2 stitching weights of one pair live in shared memory); got N = %d, d = %d",
              who, kPitMaxN, kPitMaxD, d->N, d->d);
    return AUXSSM_ERR_UNSUPPORTED;
  }
  if (d->potential < AUXSSM_POT_FLAT || d->potential > AUXSSM_POT_SV || d->gradient < 0 || d->gradient > 1) {
    set_error("%s: bad potential / gradient enum", who);
    return AUXSSM_ERR_INVALID;
  }
  return AUXSSM_OK;
}

template <typename T> static size_t pit_scratch(const auxssm_csmc_desc* d) {
  const size_t rows = (size_t)d->n_chains * d->T;
  return pit_align(rows * d->N * d->d * sizeof(T)) + pit_align(rows * d->N * sizeof(T)) + pit_align(rows * d->N * sizeof(int32_t));
}

template <typename T>
static int pit_entry(void* stream, const auxssm_csmc_desc* d, const T* x_star, const T* m0, const T* Lp0, const T* F,
                     const T* b, const T* Lq, const T* ys, const T* u, const T* shd, const T* grad, const T* normals,
                     const T* uniforms, T* x_out, int32_t* anc, void* scratch, size_t scratch_bytes) {
  const char* who = "auxssm_pit_csmc_sweep";
  int rc = check_pit_desc(d, who);
  if (rc) return rc;
  if (!x_star || !m0 || !Lp0 || !F || !b || !Lq || !u || !shd || !x_out || !anc || (d->gradient && !grad) ||
      (d->potential != AUXSSM_POT_FLAT && !ys) || ((normals == nullptr) != (uniforms == nullptr))) {
    set_error("%s: null data pointer (normals and uniforms must be both given or both NULL)", who);
    return AUXSSM_ERR_INVALID;
  }
  if (scratch == nullptr || scratch_bytes < pit_scratch<T>(d)) {
    set_error("%s: scratch too small (%zu bytes, need %zu)", who, scratch_bytes, pit_scratch<T>(d));
    return AUXSSM_ERR_SCRATCH;
  }
  PitArgs<T> a;
  a.d = *d;
  a.x_star = x_star; a.m0 = m0; a.Lp0 = Lp0; a.F = F; a.b = b; a.Lq = Lq; a.ys = ys; a.u = u; a.shd = shd; a.grad = grad;
  a.normals = normals; a.uniforms = uniforms; a.x_out = x_out; a.anc = anc;
  const size_t rows = (size_t)d->n_chains * d->T;
  unsigned char* p = static_cast<unsigned char*>(scratch);
  a.xs = reinterpret_cast<T*>(p); p += pit_align(rows * d->N * d->d * sizeof(T));
  a.lw = reinterpret_cast<T*>(p); p += pit_align(rows * d->N * sizeof(T));
  a.org = reinterpret_cast<int32_t*>(p);
  a.rng.on = normals == nullptr ? 1 : 0;
  a.rng.seed = d->rng_seed;
  a.rng.call0 = d->rng_counter;
  a.rng_counter = d->rng_counter_ptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = pit_smem_bytes<T>(d->N, d->d);
  rc = ensure_smem(pit_stitch_kernel<T>, smem, who);
  if (rc) return rc;
  const int nthr = (d->N + 31) / 32 * 32;
  pit_init_kernel<T><<<dim3(d->T, d->n_chains), nthr, 0, st>>>(a);
  AUXSSM_CHECK_LAUNCH("auxssm_pit_csmc_sweep (pit_init_kernel)");
  int K = 0;
  while ((1 << K) < d->T) ++K;
  for (int k = 0; k < K; ++k) {
    const int size = 1 << k, pairs = (1 << K) / (2 * size);
    pit_stitch_kernel<T><<<dim3(pairs, d->n_chains), kPitThreads, smem, st>>>(a, size, k == K - 1 ? 1 : 0);
    AUXSSM_CHECK_LAUNCH("auxssm_pit_csmc_sweep (pit_stitch_kernel)");
  }
  return AUXSSM_OK;
}

}  // namespace auxssm

extern "C" __attribute__((visibility("default"))) size_t auxssm_pit_csmc_scratch_bytes(const auxssm_csmc_desc* desc, int32_t elem_size) {
  if (auxssm::check_pit_desc(desc, "auxssm_pit_csmc_scratch_bytes")) return 0;
  return elem_size == 8 ? auxssm::pit_scratch<double>(desc) : auxssm::pit_scratch<float>(desc);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_pit_csmc_sweep_f32(
    void* stream, const auxssm_csmc_desc* desc, const float* x_star, const float* m0, const float* chol_P0, const float* F,
    const float* b, const float* chol_Q, const float* ys, const float* u, const float* sqrt_half_delta, const float* grad,
    const float* normals, const float* uniforms, float* x_out, int32_t* ancestors, void* scratch, size_t scratch_bytes) {
  return auxssm::pit_entry<float>(stream, desc, x_star, m0, chol_P0, F, b, chol_Q, ys, u, sqrt_half_delta, grad, normals,
                                  uniforms, x_out, ancestors, scratch, scratch_bytes);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_pit_csmc_sweep_f64(
    void* stream, const auxssm_csmc_desc* desc, const double* x_star, const double* m0, const double* chol_P0,
    const double* F, const double* b, const double* chol_Q, const double* ys, const double* u, const double* sqrt_half_delta,
    const double* grad, const double* normals, const double* uniforms, double* x_out, int32_t* ancestors, void* scratch,
    size_t scratch_bytes) {
  return auxssm::pit_entry<double>(stream, desc, x_star, m0, chol_P0, F, b, chol_Q, ys, u, sqrt_half_delta, grad, normals,
                                   uniforms, x_out, ancestors, scratch, scratch_bytes);
}

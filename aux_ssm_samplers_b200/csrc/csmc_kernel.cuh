// Conditional SMC particle Gibbs sweep (forward conditional particle filter + backward pass) as ONE
// persistent kernel: one CTA per chain, one thread per particle.
//
// Replaces aux_samplers/_primitives/csmc/csmc.py (:69-107 forward, :110-124 ancestor tracing, :127-149
// backward sampling), resamplings.py:14-37 (conditional multinomial; plus the killing scheme),
// math/utils.py:24-39 (normalize), and the proposal/weight classes of csmc/independent.py:144-268 and
// examples/stochastic_volatility/{auxiliary_csmc.py:17-48, auxiliary_guided_csmc.py:29-167}.
//
// Layout.  Particles of the current step live in shared memory ([N][d], the gather source of the next
// resampling step), each thread keeps its particle in registers; weights are normalised with a
// warp-shuffle log-sum-exp, the CDF is a warp-shuffle inclusive scan (canonical order)
// kept in shared memory for the inverse-CDF binary search.  The particle history (xs, log_ws, As) is
// streamed to HBM once with coalesced stores and re-read once by the backward pass of the same CTA.
// CTAs are persistent: a CTA loops over chains and re-uses ONE history slot, so scratch is
// (#resident CTAs) x history, independent of the number of chains.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "smem_linalg.cuh"

namespace auxssm {

// The backward pass consumes ONE uniform per time step, the same for every thread of the CTA.  A Philox block holds
// 4 (fp32) / 2 (fp64) consecutive elements, so each thread keeps the last block it generated instead of running the
// generator once per step (a quarter / half of the backward pass's generator work).
template <typename T> struct BwdCache {
  uint64_t blk = ~(uint64_t)0;
  T v[PhiloxPer<T>::N];
};

template <typename T> struct CsmcArgs {
  auxssm_csmc_desc d;
  const T *x_star, *m0, *Lp0, *F, *b, *Lq, *ys, *u, *shd, *grad, *K, *cL, *normals, *u_res, *u_kill, *u_bwd;
  T* x_out;
  int32_t* anc;
  T* xs_hist;
  T* lw_hist;
  int32_t* a_hist;
  int hist_per_chain;  // 1: history arrays are [n_chains, ...] outputs; 0: [gridDim.x, ...] scratch slots
  RngCtl rng;          // in-kernel Philox draws when the draw pointers are NULL (philox.cuh)
  const uint64_t* rng_counter;  // optional device-side call counter overriding rng.call0 (CUDA-graph replays)

  // element `e` of the logical draw arrays (same layout as the external ones)
  __device__ __forceinline__ uint64_t call0() const { return rng_counter ? rng_counter[0] : rng.call0; }
  __device__ __forceinline__ uint64_t call_bwd() const { return call0() + (d.resampling == AUXSSM_RESAMPLE_KILLING ? 3 : 2); }
  __device__ __forceinline__ T draw_u_res(int64_t e) const {
    return rng.on ? philox_elem<T>(PHILOX_UNIFORM, rng.seed, call0() + 1, (uint64_t)e) : u_res[e];
  }
  __device__ __forceinline__ T draw_u_kill(int64_t e) const {
    return rng.on ? philox_elem<T>(PHILOX_UNIFORM, rng.seed, call0() + 2, (uint64_t)e) : u_kill[e];
  }
  __device__ __forceinline__ T draw_u_bwd(int64_t e) const {
    return rng.on ? philox_elem<T>(PHILOX_UNIFORM, rng.seed, call_bwd(), (uint64_t)e) : u_bwd[e];
  }
  __device__ __forceinline__ T draw_u_bwd(int64_t e, BwdCache<T>& bc) const {
    if (!rng.on) return u_bwd[e];
    constexpr int PER = PhiloxPer<T>::N;
    const uint64_t blk = (uint64_t)e / PER;
    if (blk != bc.blk) { philox_block(PHILOX_UNIFORM, rng.seed, call_bwd(), blk, bc.v); bc.blk = blk; }
    return philox_pick(bc.v, (int)((uint64_t)e % PER));
  }
  template <int NMAX> __device__ __forceinline__ void draw_normals(int64_t e0, int n, T (&out)[NMAX]) const {
    if (rng.on) {
      philox_run<T, NMAX>(PHILOX_NORMAL, rng.seed, call0(), (uint64_t)e0, n, out);
    } else {
#pragma unroll
      for (int k = 0; k < NMAX; ++k) if (k < n) out[k] = normals[e0 + k];
    }
  }
};

template <typename T> __device__ __forceinline__ T log_2pi_half() { return T(0.91893853320467274178); }

// ---- canonical block reductions (blockDim.x a multiple of 32, at most 32 warps) ---------------------
// Two levels of the same 32-wide pattern: inside each warp, then across the (zero-padded) 32 warp totals,
// which every warp recomputes redundantly with shuffles (no serial loop over warps).
// `red` is a 32-word buffer that must not be rewritten before another barrier has been passed.

template <typename T> __device__ __forceinline__ T block_max(T v, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncwarp();
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  if (lane == 0) red[w] = v;
  __syncthreads();
  T m = (lane < nw) ? red[lane] : -INFINITY;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  return m;
}

// xor butterfly (16, 8, 4, 2, 1) inside each warp, then the same butterfly over the warp totals
// (matches oracle.canonical_sum)
template <typename T> __device__ __forceinline__ T block_sum(T v, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncwarp();
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) red[w] = v;
  __syncthreads();
  T s = (lane < nw) ? red[lane] : T(0);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

// Hillis-Steele inclusive scan (1, 2, 4, 8, 16) inside each warp, the same scan over the warp totals gives
// the warp offsets, one add of the offset (matches oracle.canonical_cdf)
template <typename T> __device__ __forceinline__ T block_cdf(T v, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncwarp();
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const T o = __shfl_up_sync(0xffffffffu, v, off);
    if (lane >= off) v = v + o;
  }
  if (lane == 31) red[w] = v;
  __syncthreads();
  T tot = (lane < nw) ? red[lane] : T(0);
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const T o = __shfl_up_sync(0xffffffffu, tot, off);
    if (lane >= off) tot = tot + o;
  }
  // inclusive scan of the warp totals; the offset of warp w is the inclusive value of warp w - 1
  const T offs = __shfl_sync(0xffffffffu, tot, (w > 0) ? w - 1 : 0);
  return (w > 0) ? v + offs : v;
}

// w = exp(lw - logsumexp(lw)) over the N active threads (math/utils.py:24-39)
// `red` holds 2 x 32 words (one per reduction).
template <typename T> __device__ __forceinline__ T block_normalize(T lw, bool active, T* red) {
  T amax = block_max<T>(active ? lw : -INFINITY, red);
  if (!isfinite(amax)) amax = T(0);
  const T e = active ? exp(lw - amax) : T(0);
  const T s = block_sum<T>(e, red + 32);
  const T lse = log(s) + amax;
  return active ? exp(lw - lse) : T(0);
}

// first j with cdf[j] >= r (searchsorted side='left'), clamped to n - 1
template <typename T> __device__ __forceinline__ int search_left(const T* cdf, T r, int n) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] < r) lo = mid + 1; else hi = mid;
  }
  return lo < n - 1 ? lo : n - 1;
}

// Conditional systematic resampling (resamplings.py:40-86; Chopin & Singh Alg. 4) from the three uniforms
// (U, V, W) of resamplings.py:65.  Every thread evaluates the plan redundantly (a handful of flops, no barrier).
// The grid r_j = (j + uniform) / N is increasing, so searchsorted gives a non-decreasing index vector whose zeros
// are its first n_zero entries: zero_loc[k] = k, and the roll of :84 is by shift = floor(n_zero W).  Particle i
// therefore takes the ancestor of grid point (i + shift) mod N; no index exchange between threads is needed.
// n_zero = #{j : r_j <= cdf[0]} is found by walking down from an over-estimate with the exact predicate.
template <typename T> struct SysPlan {
  T uniform;
  int shift;
  __device__ __forceinline__ T r(int i, int N) const {
    int j = i + shift;
    if (j >= N) j -= N;
    if (j < 0) j += N;
    return (T(j) + uniform) / T(N);
  }
};
template <typename T> __device__ __forceinline__ SysPlan<T> systematic_plan(const T* cdf, int N, T U, T V, T W) {
  SysPlan<T> pl;
  const T w0 = cdf[0];  // the canonical CDF starts with the weight of particle 0 itself
  const T tmp = T(N) * w0, fl = floor(tmp);
  if (tmp <= T(1)) {
    pl.uniform = tmp * U;
  } else {
    const T rem = tmp - fl, p_cond = rem * (fl + T(1)) / tmp;
    pl.uniform = (V < p_cond) ? rem * U : rem + (T(1) - rem) * U;
  }
  int j = (int)floor(tmp - pl.uniform) + 2;
  if (j > N - 1) j = N - 1;
  for (; j >= 0; --j)
    if ((T(j) + pl.uniform) / T(N) <= w0) break;
  const int n_zero = j + 1;
  pl.shift = 0;
  if (n_zero != 1) {
    const int roll = (int)floor(T(n_zero) * W);
    pl.shift = (roll < n_zero) ? roll : -1;  // out-of-range gather of zero_loc (fill value -1), resamplings.py:82
  }
  return pl;
}

// ---- per-particle dense algebra in registers (matrices in shared memory, row stride D) -------------

template <typename T, int D> __device__ __forceinline__ void matvec(const T* M, const T (&x)[D], T (&out)[D]) {
  constexpr int V = (D >= 4) ? 4 : D;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T a0 = T(0), a1 = T(0);
#pragma unroll
    for (int k = 0; k < D; k += V) {
      T m[V];
      ldv<T, V>(m, M + i * D + k);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if ((q & 1) == 0) a0 = fma_(m[q], x[k + q], a0); else a1 = fma_(m[q], x[k + q], a1);
      }
    }
    out[i] = a0 + a1;
  }
}

// out = L eps with L lower triangular
template <typename T, int D> __device__ __forceinline__ void matvec_lower(const T* L, const T (&x)[D], T (&out)[D]) {
  constexpr int V = (D >= 4) ? 4 : D;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T a0 = T(0), a1 = T(0);
#pragma unroll
    for (int k = 0; k < (i / V + 1) * V; k += V) {
      T m[V];
      ldv<T, V>(m, L + i * D + k);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (k + q <= i) {
          if ((q & 1) == 0) a0 = fma_(m[q], x[k + q], a0); else a1 = fma_(m[q], x[k + q], a1);
        }
      }
    }
    out[i] = a0 + a1;
  }
}

// -0.5 |L^{-1} r|^2 ; r is overwritten by L^{-1} r.  invd = 1 / diag(L) (0 beyond the logical dim).
template <typename T, int D> __device__ __forceinline__ T quad_form(const T* L, const T* invd, T (&r)[D]) {
  constexpr int V = (D >= 4) ? 4 : D;
  T q2 = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T a0 = r[i], a1 = T(0);
#pragma unroll
    for (int k = 0; k < ((i + V - 1) / V) * V; k += V) {
      T l[V];
      ldv<T, V>(l, L + i * D + k);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (k + q < i) {
          if ((q & 1) == 0) a0 = fma_(-l[q], r[k + q], a0); else a1 = fma_(-l[q], r[k + q], a1);
        }
      }
    }
    r[i] = (a0 + a1) * invd[i];
    q2 = fma_(r[i], r[i], q2);
  }
  return T(-0.5) * q2;
}

// Shared-memory image of the model (all matrices D x D row-major, zero padded).
template <typename T, int D> struct SModel {
  T *F, *Lq, *Lp0, *K, *cL;                   // D*D each (K, cL: GUIDED only)
  T *b, *m0, *invq, *invp0, *invl, *y, *xs_t; // D each: b, m0, 1/diag factors, y_t, x*_t
  T *u_t, *mean_t, *xsel;                     // D each: auxiliary u_t, proposal mean, selected x_{t+1}
  T cst_q, cst_p0, cst_l;                     // -sum log diag - d/2 log 2pi of chol Q, chol P0, chol Lambda_t
  T s_t;                                      // sqrt(delta_t / 2)
};

template <typename T, int D> __device__ __forceinline__ T potential(const auxssm_csmc_desc& dsc, const T* y, const T (&x)[D]) {
  if (dsc.potential == AUXSSM_POT_FLAT) return T(0);
  T acc = T(0);
  if (dsc.potential == AUXSSM_POT_GAUSSIAN) {
    const T sig = T(dsc.sig), ls = log(sig);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i < dsc.d) {
        const T z = (y[i] - x[i]) / sig;
        acc += T(-0.5) * z * z - ls - log_2pi_half<T>();
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i < dsc.d) {
        // log N(y; 0, exp(x / 2)) = -y^2 exp(-x) / 2 - x / 2 - log(2 pi) / 2: one exp, no log / divide
        acc += T(-0.5) * y[i] * y[i] * exp(-x[i]) - T(0.5) * x[i] - log_2pi_half<T>();
      }
    }
  }
  return acc;
}

// sum_i log N(x_i; m_i, s)
template <typename T, int D>
__device__ __forceinline__ T iso_logpdf(const T (&x)[D], const T* m, T s, int d) {
  T acc = T(0);
  const T ls = log(s);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (i < d) {
      const T z = (x[i] - m[i]) / s;
      acc += T(-0.5) * z * z - ls - log_2pi_half<T>();
    }
  }
  return acc;
}

// log N(xn; F xp + b, Q)   (t >= 1)   or   log N(xn; m0, P0)   (t == 0)
template <typename T, int D>
__device__ __forceinline__ T prior_logpdf(const SModel<T, D>& m, bool first, const T (&xn)[D], const T (&xpred)[D]) {
  T r[D];
#pragma unroll
  for (int i = 0; i < D; ++i) r[i] = xn[i] - xpred[i];
  return first ? quad_form<T, D>(m.Lp0, m.invp0, r) + m.cst_p0 : quad_form<T, D>(m.Lq, m.invq, r) + m.cst_q;
}

// Proposal + log-weight of one particle at time t.  xp = resampled parent (ignored at t = 0), eps = N(0,I).
template <typename T, int D, int PROP>
__device__ __forceinline__ T propose_and_weight(const auxssm_csmc_desc& dsc, const SModel<T, D>& m, bool first,
                                                bool pinned, const T (&xp)[D], const T (&eps)[D], T (&xn)[D]) {
  T xpred[D];
  if (first) {
#pragma unroll
    for (int i = 0; i < D; ++i) xpred[i] = m.m0[i];
  } else {
    matvec<T, D>(m.F, xp, xpred);
#pragma unroll
    for (int i = 0; i < D; ++i) xpred[i] += m.b[i];
  }
  T noise[D];
  if (PROP == AUXSSM_PROP_BOOTSTRAP) {
    matvec_lower<T, D>(first ? m.Lp0 : m.Lq, eps, noise);
#pragma unroll
    for (int i = 0; i < D; ++i) xn[i] = pinned ? m.xs_t[i] : xpred[i] + noise[i];
    return potential<T, D>(dsc, m.y, xn);
  } else if (PROP == AUXSSM_PROP_INDEPENDENT) {
#pragma unroll
    for (int i = 0; i < D; ++i) xn[i] = pinned ? m.xs_t[i] : m.mean_t[i] + m.s_t * eps[i];
    T lw = prior_logpdf<T, D>(m, first, xn, xpred) + potential<T, D>(dsc, m.y, xn);
    // gradient == 1 follows the reference to the letter: GradientAuxiliaryG0 (independent.py:173-190) corrects every
    // particle, GradientAuxiliaryGt (:252-268) sums its correction over ALL particles (no axis=-1), i.e. adds a
    // per-step constant that cancels in every normalisation -- omitted here.  gradient == 2 applies the per-particle
    // correction at every step (the importance weight the proposal calls for).
    if (dsc.gradient == 2 || (dsc.gradient == 1 && first))
      lw += iso_logpdf<T, D>(xn, m.u_t, m.s_t, dsc.d) - iso_logpdf<T, D>(xn, m.mean_t, m.s_t, dsc.d);
    return lw;
  } else {
    // GUIDED: mu = xpred + K (u - xpred); x = mu + cL eps
    T diff[D], mu[D];
#pragma unroll
    for (int i = 0; i < D; ++i) diff[i] = m.u_t[i] - xpred[i];
    matvec<T, D>(m.K, diff, mu);
#pragma unroll
    for (int i = 0; i < D; ++i) mu[i] += xpred[i];
    matvec_lower<T, D>(m.cL, eps, noise);
#pragma unroll
    for (int i = 0; i < D; ++i) xn[i] = pinned ? m.xs_t[i] : mu[i] + noise[i];
    T lw = potential<T, D>(dsc, m.y, xn) + prior_logpdf<T, D>(m, first, xn, xpred) +
           iso_logpdf<T, D>(xn, m.u_t, m.s_t, dsc.d);
    T r[D];
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = xn[i] - mu[i];
    lw -= quad_form<T, D>(m.cL, m.invl, r) + m.cst_l;
    return lw;
  }
}

template <int D> __host__ __device__ inline size_t csmc_smem_elems(int N, int d, int prop) {
  const size_t mats = (prop == AUXSSM_PROP_GUIDED) ? 5 : 3;
  return mats * D * D + 16 * D + 2 * (size_t)N * d + N + 256;
}

// Loads a [d x d] row-major global matrix into a D x D zero-padded shared image.
template <typename T, int D> __device__ __forceinline__ void load_padded(const T* g, T* s, int d) {
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
    const int i = e / D, k = e % D;
    s[e] = (g != nullptr && i < d && k < d) ? g[i * d + k] : T(0);
  }
}

// 1 / diag and the Gaussian normaliser of a lower factor held in shared memory (mvn/base.py:16-58:
// non-finite diagonal entries are dropped from the dimension count and from the log-determinant).
template <typename T, int D> __device__ __forceinline__ T factor_consts(const T* L, T* invd, int d) {
  T cst = T(0);
  for (int i = 0; i < D; ++i) {
    const T lii = (i < d) ? L[i * D + i] : T(0);
    const bool ok = (i < d) && isfinite(lii);
    if (threadIdx.x == 0) invd[i] = ok ? T(1) / lii : T(0);
    if (ok) cst -= log(fabs(lii)) + log_2pi_half<T>();
  }
  return cst;
}

template <typename T, int D, int PROP, int MAXT>
__global__ void __launch_bounds__(MAXT) csmc_sweep_kernel(CsmcArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const auxssm_csmc_desc& dsc = a.d;
  const int N = dsc.N, d = dsc.d, Tn = dsc.T, i = threadIdx.x;
  const bool active = i < N;
  T* p = reinterpret_cast<T*>(smem_raw);
  SModel<T, D> m;
  m.F = p; p += D * D;
  m.Lq = p; p += D * D;
  m.Lp0 = p; p += D * D;
  m.K = m.cL = nullptr;
  if (PROP == AUXSSM_PROP_GUIDED) { m.K = p; p += D * D; m.cL = p; p += D * D; }
  m.b = p; p += D; m.m0 = p; p += D; m.invq = p; p += D; m.invp0 = p; p += D; m.invl = p; p += D;
  m.y = p; p += D; m.xs_t = p; p += D; m.u_t = p; p += D; m.mean_t = p; p += D; m.xsel = p; p += D;
  T* red = p; p += 128;  // 4 x 32: max, sum, cdf, killing-max (never reused without a barrier in between)
  T* misc = p; p += 32;  // misc[0]: max weight (killing); misc[8..]: selected index as int
  int* isel = reinterpret_cast<int*>(misc + 8);
  T* cdf = p; p += N;
  T* part = p; p += (size_t)N * d;   // particles of the current step, [N][d]
  T* stage = p;                      // [N][d] staging for coalesced normals / history loads

  load_padded<T, D>(a.F, m.F, d);
  load_padded<T, D>(a.Lq, m.Lq, d);
  load_padded<T, D>(a.Lp0, m.Lp0, d);
  if (i < D) {
    m.b[i] = (i < d) ? a.b[i] : T(0);
    m.m0[i] = (i < d) ? a.m0[i] : T(0);
    m.invl[i] = T(0); m.y[i] = T(0); m.u_t[i] = T(0); m.mean_t[i] = T(0);
  }
  __syncthreads();
  m.cst_q = factor_consts<T, D>(m.Lq, m.invq, d);
  m.cst_p0 = factor_consts<T, D>(m.Lp0, m.invp0, d);
  m.cst_l = T(0);
  m.s_t = T(1);
  __syncthreads();

  // Time-varying dynamics (Mt.params scanned over t, csmc.py:103) and a separate backward model Pt (csmc.py:47-50):
  // the shared-memory image of (F, b, chol Q) is reloaded at the steps where it changes.
  const bool tv = dsc.dyn_time_varying != 0;
  const T* ptF = dsc.pt_F ? static_cast<const T*>(dsc.pt_F) : a.F;
  const T* ptb = dsc.pt_b ? static_cast<const T*>(dsc.pt_b) : a.b;
  const T* ptLq = dsc.pt_chol_Q ? static_cast<const T*>(dsc.pt_chol_Q) : a.Lq;
  const bool pt_differs = dsc.backward && (dsc.pt_F || dsc.pt_b || dsc.pt_chol_Q);
  auto load_dynamics = [&](const T* Fg, const T* bg, const T* Lg, int idx) {  // all threads; ends with a barrier
    const int64_t o2 = tv ? (int64_t)idx * d * d : 0, o1 = tv ? (int64_t)idx * d : 0;
    __syncthreads();
    load_padded<T, D>(Fg + o2, m.F, d);
    load_padded<T, D>(Lg + o2, m.Lq, d);
    if (i < D) m.b[i] = (i < d) ? bg[o1 + i] : T(0);
    __syncthreads();
    m.cst_q = factor_consts<T, D>(m.Lq, m.invq, d);
    __syncthreads();
  };

  for (int c = blockIdx.x; c < dsc.n_chains; c += gridDim.x) {
    if (pt_differs && !tv && c != (int)blockIdx.x) load_dynamics(a.F, a.b, a.Lq, 0);  // the last chain left Pt behind
    const int64_t hs = a.hist_per_chain ? c : blockIdx.x;  // history slot
    T* xs_h = a.xs_hist + hs * (int64_t)Tn * N * d;
    T* lw_h = a.lw_hist + hs * (int64_t)Tn * N;
    int32_t* a_h = a.a_hist ? a.a_hist + hs * (int64_t)(Tn > 1 ? Tn - 1 : 0) * N : nullptr;
    const T* xstar = a.x_star + (int64_t)c * Tn * d;
    const T* ysc = a.ys ? a.ys + (int64_t)c * dsc.ys_chain_stride : nullptr;
    T w = T(0);
    T xn[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xn[k] = T(0);
    // Software pipeline: the draws and per-step parameters of step t+1 are requested at the top of step t,
    // so no DRAM round trip sits on the (strictly sequential) critical path of the time loop.
    constexpr bool DIRECT = (D <= 4);  // normals read straight into registers (coalesced for small d)
    T ur_n = T(0), uk_n = T(0), uw_n = T(0), eps_n[D], xs_n = T(0), y_n = T(0), u_n = T(0), g_n = T(0), s_n = T(1);
#pragma unroll
    for (int k = 0; k < D; ++k) eps_n[k] = T(0);
    auto prefetch = [&](int tt) {  // everything step tt needs
      const bool go = tt < Tn;
      if (go && active) {
        if (tt > 0 && dsc.resampling == AUXSSM_RESAMPLE_SYSTEMATIC) {
          const int64_t us = ((int64_t)c * (Tn - 1) + (tt - 1)) * 3;  // (U, V, W), broadcast loads
          ur_n = a.draw_u_res(us); uk_n = a.draw_u_res(us + 1); uw_n = a.draw_u_res(us + 2);
        } else if (tt > 0) {
          ur_n = a.draw_u_res(((int64_t)c * (Tn - 1) + (tt - 1)) * N + i);
          if (dsc.resampling == AUXSSM_RESAMPLE_KILLING) uk_n = a.draw_u_kill(((int64_t)c * (Tn - 1) + (tt - 1)) * N + i);
        }
        if (DIRECT) a.template draw_normals<D>((((int64_t)c * Tn + tt) * N + i) * d, d, eps_n);
      }
      if (go && i < D) {
        xs_n = (i < d) ? xstar[(int64_t)tt * d + i] : T(0);
        if (ysc) y_n = (i < d) ? ysc[(int64_t)tt * d + i] : T(0);
        if (PROP != AUXSSM_PROP_BOOTSTRAP) {
          u_n = (i < d) ? a.u[((int64_t)c * Tn + tt) * d + i] : T(0);
          g_n = (a.grad != nullptr && i < d) ? a.grad[((int64_t)c * Tn + tt) * d + i] : T(0);
        }
      }
      if (go && PROP != AUXSSM_PROP_BOOTSTRAP) s_n = a.shd[(int64_t)c * Tn + tt];
    };
    prefetch(0);

    // ------------------------------ forward conditional particle filter ------------------------------
    for (int t = 0; t < Tn; ++t) {
      const bool first = (t == 0);
      // (1) conditional resampling against the CDF of the previous step, (2) gather the parent
      int anc = 0;
      T xp[D];
#pragma unroll
      for (int k = 0; k < D; ++k) xp[k] = T(0);
      // take this step's prefetched values, then request the next step's
      const T ur = ur_n, uk = uk_n, uw = uw_n, xs_c = xs_n, y_c = y_n, u_c = u_n, g_c = g_n, s_c = s_n;
      T eps[D];
#pragma unroll
      for (int k = 0; k < D; ++k) eps[k] = eps_n[k];
      prefetch(t + 1);
      if (!first && active) {
        const T r = (dsc.resampling == AUXSSM_RESAMPLE_SYSTEMATIC) ? systematic_plan<T>(cdf, N, ur, uk, uw).r(i, N)
                                                                   : cdf[N - 1] * (T(1) - ur);
        anc = search_left<T>(cdf, r, N);
        // killing: survive iff u_kill * max_j w_j < w_i  (max weight of the previous step kept in misc[0])
        if (dsc.resampling == AUXSSM_RESAMPLE_KILLING && uk * misc[0] < w) anc = i;
        if (i == 0) anc = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) if (k < d) xp[k] = part[anc * d + k];
      }
      if (tv && !first) load_dynamics(a.F, a.b, a.Lq, t - 1);  // parents are in registers; `part` is not touched
      // (3) per-step parameters -> shared (read after the barrier below)
      if (i < D) {
        m.xs_t[i] = xs_c;
        if (ysc) m.y[i] = y_c;
        if (PROP != AUXSSM_PROP_BOOTSTRAP) {
          m.u_t[i] = u_c;
          m.mean_t[i] = u_c + s_c * s_c * g_c;
        }
      }
      if (PROP != AUXSSM_PROP_BOOTSTRAP) m.s_t = s_c;
      if (!DIRECT) {
        __syncthreads();  // every parent gathered before `stage` / `part` are overwritten
        if (PROP == AUXSSM_PROP_GUIDED && (t <= 1 || dsc.gains_time_stride != 0)) {
          load_padded<T, D>(a.K + (int64_t)t * dsc.gains_time_stride, m.K, d);
          load_padded<T, D>(a.cL + (int64_t)t * dsc.gains_time_stride, m.cL, d);
        }
        // normals of this step, staged with coalesced loads (or generated in place)
        const int64_t e0 = ((int64_t)c * Tn + t) * N * d;
        if (a.rng.on) {
          for (int e = i; e < N * d; e += blockDim.x) stage[e] = philox_elem<T>(PHILOX_NORMAL, a.rng.seed, a.call0(), (uint64_t)(e0 + e));
        } else {
          for (int e = i; e < N * d; e += blockDim.x) stage[e] = a.normals[e0 + e];
        }
      } else if (PROP == AUXSSM_PROP_GUIDED && (t <= 1 || dsc.gains_time_stride != 0)) {
        load_padded<T, D>(a.K + (int64_t)t * dsc.gains_time_stride, m.K, d);
        load_padded<T, D>(a.cL + (int64_t)t * dsc.gains_time_stride, m.cL, d);
      }
      __syncthreads();  // parents gathered, parameters / staged normals visible
      if (PROP == AUXSSM_PROP_GUIDED && (t <= 1 || dsc.gains_time_stride != 0)) {
        m.cst_l = factor_consts<T, D>(m.cL, m.invl, d);
        __syncthreads();
      }
      if (!DIRECT) {
#pragma unroll
        for (int k = 0; k < D; ++k) eps[k] = (active && k < d) ? stage[i * d + k] : T(0);
      }
      // (4) propose, pin particle 0 to the reference trajectory, (5) weight
      T lw = propose_and_weight<T, D, PROP>(dsc, m, first, i == 0, xp, eps, xn);
      if (active) {
#pragma unroll
        for (int k = 0; k < D; ++k) if (k < d) part[i * d + k] = xn[k];
        lw_h[(int64_t)t * N + i] = lw;
        if (!first && a_h) a_h[(int64_t)(t - 1) * N + i] = anc;
      }
      __syncthreads();
      // (6) history: coalesced copy of the [N][d] block
      T* xg = xs_h + (int64_t)t * N * d;
      for (int e = i; e < N * d; e += blockDim.x) xg[e] = part[e];
      // (7) normalise, CDF
      w = block_normalize<T>(lw, active, red);
      const T cd = block_cdf<T>(w, red + 64);
      if (active) cdf[i] = cd;
      if (dsc.resampling == AUXSSM_RESAMPLE_KILLING) {
        const T wmax = block_max<T>(w, red + 96);
        if (i == 0) misc[0] = wmax;
      }
      __syncthreads();
    }

    // ------------------------------------ backward pass ----------------------------------------------
    T* xo = a.x_out + (int64_t)c * Tn * d;
    int32_t* ao = a.anc + (int64_t)c * Tn;
    const int64_t ub = (int64_t)c * Tn;
    BwdCache<T> bwdc;
    // B_T ~ Cat(w_T): index = #{j : cdf[j] < r}, clamped (csmc.py:111 / :131)
    {
      const T r = cdf[N - 1] * (T(1) - a.draw_u_bwd(ub, bwdc));
      int B = __syncthreads_count(active && cdf[i] < r);
      if (B > N - 1) B = N - 1;
      if (i == B) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          if (k < d) { xo[(int64_t)(Tn - 1) * d + k] = xn[k]; }
          m.xsel[k] = (k < d) ? xn[k] : T(0);
        }
        ao[Tn - 1] = B;
      }
      if (i == 0) isel[0] = B;
      __syncthreads();
    }
    if (dsc.backward) {
      T ub_n = (Tn > 1) ? a.draw_u_bwd(ub + 1, bwdc) : T(0);
      if (pt_differs && !tv && Tn > 1) load_dynamics(ptF, ptb, ptLq, 0);
      for (int t = Tn - 2, k2 = 1; t >= 0; --t, ++k2) {
        T xt[D], xpred[D], xsel[D];
        T lwh = T(0);
        const T ub_c = ub_n;
        if (t > 0) ub_n = a.draw_u_bwd(ub + k2 + 1, bwdc);
        if (tv && (pt_differs || t < Tn - 2)) load_dynamics(ptF, ptb, ptLq, t);  // Pt(x_{t+1} | x_t): entry t
        {
          // history of step t, staged with coalesced loads
          const T* xg = xs_h + (int64_t)t * N * d;
          for (int e = i; e < N * d; e += blockDim.x) stage[e] = xg[e];
          if (active) lwh = lw_h[(int64_t)t * N + i];
          __syncthreads();
#pragma unroll
          for (int k = 0; k < D; ++k) xt[k] = (active && k < d) ? stage[i * d + k] : T(0);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) xsel[k] = m.xsel[k];
        matvec<T, D>(m.F, xt, xpred);
#pragma unroll
        for (int k = 0; k < D; ++k) xpred[k] += m.b[k];
        T lw = prior_logpdf<T, D>(m, false, xsel, xpred) + lwh;
        const T wb = block_normalize<T>(lw, active, red);
        const T cd = block_cdf<T>(wb, red + 64);
        if (active) cdf[i] = cd;
        __syncthreads();
        const T r = cdf[N - 1] * (T(1) - ub_c);
        int B = __syncthreads_count(active && cdf[i] < r);
        if (B > N - 1) B = N - 1;
        __syncthreads();  // everyone has read m.xsel / cdf before they change
        if (i == B) {
#pragma unroll
          for (int k = 0; k < D; ++k) {
            if (k < d) xo[(int64_t)t * d + k] = xt[k];
            m.xsel[k] = xt[k];
          }
          ao[t] = B;
        }
        __syncthreads();
      }
    } else {
      // ancestor tracing: B_{t-1} = A_t[B_t]  (csmc.py:116-121); a serial pointer chase by one thread
      if (i == 0) {
        int B = isel[0];
        for (int t = Tn - 2; t >= 0; --t) {
          B = a_h[(int64_t)t * N + B];
          ao[t] = B;
          for (int k = 0; k < d; ++k) xo[(int64_t)t * d + k] = xs_h[((int64_t)t * N + B) * d + k];
        }
      }
      __syncthreads();
    }
    __syncthreads();
  }
}

}  // namespace auxssm

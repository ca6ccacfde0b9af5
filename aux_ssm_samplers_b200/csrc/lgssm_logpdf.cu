// Trajectory log-densities of a linear-Gaussian state-space model.
//
// Replaces aux_samplers/_primitives/kalman/base.py of the reference: posterior_logpdf:73-96,
// prior_logpdf:100-134, log_likelihood:138-166, with the Cholesky-parametrised Gaussian density of
// aux_samplers/_primitives/math/mvn/base.py:16-58 and its nansum conventions (a NaN term -- e.g. a
// missing observation -- contributes 0).
//
// One CTA per (unit, time tile); the per-step terms are summed in time order inside the CTA and the
// per-tile partial sums are reduced in a fixed order by a second tiny kernel (deterministic).
// Cholesky factors of time-invariant Q / R are computed once per CTA.
#include "common.cuh"
#include "smem_linalg.cuh"

namespace auxssm {

template <typename T> struct LogpdfArgs {
  auxssm_lgssm_desc d;
  uint32_t which;
  const T *ys, *xs, *ell, *m0, *P0, *Fs, *Qs, *bs, *Hs, *Rs, *cs;
  T* out;
  T* partial;  // [units][n_tiles]
  int n_tiles, tile_len;
};

template <typename T> __device__ __forceinline__ T half_log_2pi_() { return T(0.91893853320467274178); }

// log N(r; 0, L L^T) for the residual in `r` (overwritten), mvn/base.py:16-58: non-finite diagonal
// entries of L are dropped from the dimension count and the log-determinant.  Uniform result.
// scalar_norm: the dy = 1 branch of base.py:160-162 is jax.scipy.stats.norm.logpdf, which keeps the -log(scale) term of
// a non-finite scale (an infinite variance gives -inf there, where the multivariate form drops the entry).
template <typename T, int D>
__device__ __forceinline__ T gauss_logpdf(const T* L, T* r, T* red, int n, const T* dg, bool scalar_norm = false) {
  constexpr int LD = Tile<D>::LD;
  fwd_solve<T, D>(L, nullptr, r, n, 0, dg);
  T part = T(0);
  const int i = threadIdx.x;
  if (i < n) {
    const T lii = L[i * LD + i];
    if (finite_(lii) || scalar_norm) part = T(-0.5) * r[i] * r[i] - log(fabs(lii)) - half_log_2pi_<T>();
  }
  return cta_sum<T, D>(part, red);
}

template <int D> __host__ __device__ constexpr size_t logpdf_smem_elems() {
  return (size_t)4 * D * Tile<D>::LD + 7 * D + 2 * CholWs<D>::WORDS + 16;
}

template <typename T, int D>
__global__ void __launch_bounds__(Tile<D>::NT, Occ<T, D>::MINB) lgssm_logpdf_kernel(LogpdfArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* p = reinterpret_cast<T*>(smem_raw);
  constexpr int NT = Tile<D>::NT, LD = Tile<D>::LD;
  for (size_t e = threadIdx.x; e < logpdf_smem_elems<D>(); e += NT) p[e] = T(0);
  __syncthreads();
  T* Ft = p; p += D * LD;
  T* LQ = p; p += D * LD;
  T* Ht = p; p += D * LD;
  T* LR = p; p += D * LD;
  T* xprev = p; p += D;
  T* xcur = p; p += D;
  T* r = p; p += D;
  T* bv = p; p += D;
  T* cv = p; p += D;
  T* yv = p; p += D;
  T* red = p; p += D;
  T* dgQ = p; p += CholWs<D>::WORDS;
  T* dgR = p; p += CholWs<D>::WORDS;
  const int unit = blockIdx.x, tile = blockIdx.y;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx, dy = a.d.dy, Tn = a.d.T;
  const bool hid = (a.d.flags & AUXSSM_FLAG_H_IDENTITY) != 0;
  const bool do_prior = (a.which & AUXSSM_LOGPDF_PRIOR) != 0, do_lik = (a.which & AUXSSM_LOGPDF_LOGLIK) != 0;
  const int t0 = tile * a.tile_len;
  int t1 = t0 + a.tile_len;
  if (t1 > Tn) t1 = Tn;
  const int i = threadIdx.x;
  T total = T(0);
  bool first_dyn = true, first_obs = true;
  for (int t = t0; t < t1; ++t) {
    cta_sync<D>();
    if (i < dx) {
      const T* xg = a.xs + unit_off(a.d.xs, c, b);
      xcur[i] = xg[(int64_t)t * a.d.xs.time + i];
      if (t > 0 && t == t0) xprev[i] = xg[(int64_t)(t - 1) * a.d.xs.time + i];
    }
    cta_sync<D>();
    if (do_prior) {
      if (t == 0) {
        load_mat<T, D, false>(a.P0 + unit_off(a.d.P0, c, b), LQ, dx, dx);
        if (i < dx) r[i] = xcur[i] - a.m0[unit_off(a.d.m0, c, b) + i];
        chol_inplace<T, D>(LQ, dgQ, dx);
        T term = gauss_logpdf<T, D>(LQ, r, red, dx, dgQ);
        if (!(term == term)) term = T(0);
        total += term;
        first_dyn = true;  // LQ now holds chol(P0): force a reload of Q below
      } else {
        const int tm1 = t - 1;
        if (first_dyn || a.d.Fs.time != 0)
          load_mat<T, D, true>(a.Fs + unit_off(a.d.Fs, c, b) + (int64_t)tm1 * a.d.Fs.time, Ft, dx, dx);
        if (first_dyn || a.d.bs.time != 0) {
          if (i < dx) bv[i] = a.bs[unit_off(a.d.bs, c, b) + (int64_t)tm1 * a.d.bs.time + i];
        }
        if (first_dyn || a.d.Qs.time != 0) {
          cta_sync<D>();
          load_mat<T, D, false>(a.Qs + unit_off(a.d.Qs, c, b) + (int64_t)tm1 * a.d.Qs.time, LQ, dx, dx);
          chol_inplace<T, D>(LQ, dgQ, dx);
        }
        first_dyn = false;
        mv_tn<T, D, 1>(Ft, xprev, r, bv, dx, dx);  // r = F x_{t-1} + b
        if (i < dx) r[i] = xcur[i] - r[i];
        T term = gauss_logpdf<T, D>(LQ, r, red, dx, dgQ);
        if (!(term == term)) term = T(0);
        total += term;
      }
    }
    if (do_lik) {
      if (!hid && (first_obs || a.d.Hs.time != 0)) {
        cta_sync<D>();
        load_mat<T, D, true>(a.Hs + unit_off(a.d.Hs, c, b) + (int64_t)t * a.d.Hs.time, Ht, dy, dx);
      }
      if (first_obs || a.d.cs.time != 0) {
        if (i < dy) cv[i] = a.cs ? a.cs[unit_off(a.d.cs, c, b) + (int64_t)t * a.d.cs.time + i] : T(0);
      }
      if (first_obs || a.d.Rs.time != 0) {
        cta_sync<D>();
        const T* gr = a.Rs + unit_off(a.d.Rs, c, b) + (int64_t)t * a.d.Rs.time;
        if (a.d.flags & AUXSSM_FLAG_R_DIAGONAL) {
          AUXSSM_FOR_ROWS(ri, rj, dy, dy) LR[ri * LD + rj] = (ri == rj) ? gr[ri] : T(0);
        } else {
          load_mat<T, D, false>(gr, LR, dy, dy);
        }
        chol_inplace<T, D>(LR, dgR, dy);
      }
      first_obs = false;
      if (i < dy) yv[i] = a.ys[unit_off(a.d.ys, c, b) + (int64_t)t * a.d.ys.time + i];
      if (hid) {
        cta_sync<D>();
        if (i < dy) r[i] = yv[i] - xcur[i] - cv[i];
      } else {
        mv_tn<T, D, 1>(Ht, xcur, r, cv, dx, dy);  // r = H x + c
        if (i < dy) r[i] = yv[i] - r[i];
      }
      T term = gauss_logpdf<T, D>(LR, r, red, dy, dgR, dy == 1);
      if (!(term == term)) term = T(0);
      total += term;
    }
    cta_sync<D>();
    if (i < dx) xprev[i] = xcur[i];
  }
  if (threadIdx.x == 0) a.partial[(int64_t)unit * a.n_tiles + tile] = total;
}

// ---- fast path: time-invariant dynamics (and, for the likelihood part, H = I with diagonal R) -------
// prep: one CTA per unit factors Q (and P0) once and evaluates the t = 0 terms; main: ONE THREAD PER
// TIME STEP, residual and forward substitution in registers, F / chol(Q) read as shared broadcasts.

template <typename T> struct LogpdfFastArgs {
  LogpdfArgs<T> a;
  T* prep;  // [units][dx*dx (chol Q) + dx (1/diag) + 2]   (+0: -sum log diag - dx/2 log 2pi, +1: t = 0 prior term)
};

template <typename T, int D>
__global__ void __launch_bounds__(Tile<D>::NT, Occ<T, D>::MINB) logpdf_prep_kernel(LogpdfFastArgs<T> fa) {
  const LogpdfArgs<T>& a = fa.a;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* p = reinterpret_cast<T*>(smem_raw);
  constexpr int NT = Tile<D>::NT, LD = Tile<D>::LD;
  for (int e = threadIdx.x; e < D * LD + 2 * D + CholWs<D>::WORDS; e += NT) p[e] = T(0);
  __syncthreads();
  T* L = p; p += D * LD;
  T* r = p; p += D;
  T* red = p; p += D;
  T* dg = p; p += CholWs<D>::WORDS;
  const int unit = blockIdx.x, c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx, i = threadIdx.x;
  T* out = fa.prep + (int64_t)unit * (dx * dx + dx + 2);
  // t = 0 term: log N(x_0; m0, P0)
  load_mat<T, D, false>(a.P0 + unit_off(a.d.P0, c, b), L, dx, dx);
  if (i < dx) r[i] = a.xs[unit_off(a.d.xs, c, b) + i] - a.m0[unit_off(a.d.m0, c, b) + i];
  chol_inplace<T, D>(L, dg, dx);
  T term = gauss_logpdf<T, D>(L, r, red, dx, dg);
  if (!(term == term)) term = T(0);
  cta_sync<D>();
  if (a.d.T > 1) {
    load_mat<T, D, false>(a.Qs + unit_off(a.d.Qs, c, b), L, dx, dx);
    chol_inplace<T, D>(L, dg, dx);
    store_mat<T, D>(out, L, dx, dx);
    T part = T(0);
    if (i < dx) {
      const T lii = L[i * LD + i];
      out[dx * dx + i] = finite_(lii) ? T(1) / lii : T(0);
      if (finite_(lii)) part = -log(fabs(lii)) - half_log_2pi_<T>();
    }
    const T cst = cta_sum<T, D>(part, red);
    if (i == 0) out[dx * dx + dx] = cst;
  }
  if (i == 0) out[dx * dx + dx + 1] = term;
}

template <typename T, int D, int NTB>
__global__ void __launch_bounds__(NTB) logpdf_fast_kernel(LogpdfFastArgs<T> fa) {
  const LogpdfArgs<T>& a = fa.a;
  constexpr int V = (D >= 4) ? 4 : D;
  __shared__ __align__(16) T sF[D * D];
  __shared__ __align__(16) T sL[D * D];
  __shared__ T sb[D], sinv[D], sc[D];
  __shared__ T sred[NTB / 32];
  const int unit = blockIdx.x, c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx, Tn = a.d.T;
  const bool do_prior = (a.which & AUXSSM_LOGPDF_PRIOR) != 0, do_lik = (a.which & AUXSSM_LOGPDF_LOGLIK) != 0;
  const T* prep = fa.prep + (int64_t)unit * (dx * dx + dx + 2);
  for (int e = threadIdx.x; e < D * D; e += NTB) {
    const int i = e / D, k = e % D;  // D is a compile-time power of two
    const bool in = (i < dx && k < dx);
    sF[e] = (do_prior && in && Tn > 1) ? a.Fs[unit_off(a.d.Fs, c, b) + i * dx + k] : T(0);
    sL[e] = (do_prior && in && Tn > 1) ? prep[i * dx + k] : T(0);
  }
  if (threadIdx.x < D) {
    const int i = threadIdx.x;
    sb[i] = (do_prior && i < dx && Tn > 1) ? a.bs[unit_off(a.d.bs, c, b) + i] : T(0);
    sinv[i] = (do_prior && i < dx && Tn > 1) ? prep[dx * dx + i] : T(0);
    sc[i] = (do_lik && a.cs != nullptr && i < dx) ? a.cs[unit_off(a.d.cs, c, b) + i] : T(0);
  }
  __syncthreads();
  // AUXSSM_FLAG_F_IS_DIAGONAL is a hint: the whole F is in shared memory anyway, so it is verified here and the
  // diagonal shortcut is only taken when it holds (a wrong hint costs time, never correctness).
  bool f_diag = (a.d.flags & AUXSSM_FLAG_F_IS_DIAGONAL) != 0;
  if (f_diag) {
    int off = 0;
    for (int e = threadIdx.x; e < D * D; e += NTB)
      if ((e / D) != (e % D) && sF[e] != T(0)) off = 1;
    f_diag = __syncthreads_or(off) == 0;
  }
  const T cst = (do_prior && Tn > 1) ? prep[dx * dx + dx] : T(0);
  const int t = blockIdx.y * NTB + threadIdx.x;
  T total = T(0);
  if (t < Tn) {
    const T* xg = a.xs + unit_off(a.d.xs, c, b) + (int64_t)t * a.d.xs.time;
    T x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = (i < dx) ? xg[i] : T(0);
    if (do_prior && t > 0) {
      const T* xp = xg - a.d.xs.time;
      T z[D];
#pragma unroll
      for (int i = 0; i < D; ++i) z[i] = (i < dx) ? xp[i] : T(0);  // z <- x_{t-1}
      T r[D];
      if (f_diag) {  // verified above: F x is one multiply per entry
#pragma unroll
        for (int i = 0; i < D; ++i) r[i] = x[i] - fma_(sF[i * D + i], z[i], sb[i]);
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          T acc0 = sb[i], acc1 = T(0);
#pragma unroll
          for (int k = 0; k < D; k += V) {
            T f[V];
            ldv<T, V>(f, sF + i * D + k);
#pragma unroll
            for (int q = 0; q < V; ++q) {
              if ((q & 1) == 0) acc0 = fma_(f[q], z[k + q], acc0); else acc1 = fma_(f[q], z[k + q], acc1);
            }
          }
          r[i] = x[i] - (acc0 + acc1);
        }
      }
      T quad = T(0);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        T acc0 = r[i], acc1 = T(0);
#pragma unroll
        for (int k = 0; k < ((i + V - 1) / V) * V; k += V) {
          T l[V];
          ldv<T, V>(l, sL + i * D + k);
#pragma unroll
          for (int q = 0; q < V; ++q) {
            if (k + q < i) {
              if ((q & 1) == 0) acc0 = fma_(-l[q], z[k + q], acc0); else acc1 = fma_(-l[q], z[k + q], acc1);
            }
          }
        }
        z[i] = (acc0 + acc1) * sinv[i];  // z now holds chol(Q)^{-1} r (entries < i already overwritten)
        quad = fma_(z[i], z[i], quad);
      }
      T term = T(-0.5) * quad + cst;
      if (!(term == term)) term = T(0);
      total += term;
    }
    if (do_lik) {  // H = I, diagonal R
      const T* yg = a.ys + unit_off(a.d.ys, c, b) + (int64_t)t * a.d.ys.time;
      const T* rg = a.Rs + unit_off(a.d.Rs, c, b) + (int64_t)t * a.d.Rs.time;
      T term = T(0);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if (i < dx) {
          const T rr = rg[i];
          const T lii = sqrt(rr);  // chol of the diagonal matrix; NaN / inf entries follow mvn/base.py:49-57
          const T e = (yg[i] - x[i] - sc[i]) / lii;
          if (finite_(lii) || dx == 1) term += T(-0.5) * e * e - log(lii) - half_log_2pi_<T>();  // dx = 1: norm.logpdf
          else term += T(-0.5) * e * e;
        }
      }
      if (!(term == term)) term = T(0);
      total += term;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    T sum = T(0);
    for (int q = 0; q < NTB / 32; ++q) sum += sred[q];
    if (blockIdx.y == 0 && do_prior) sum += prep[dx * dx + dx + 1];
    a.partial[(int64_t)unit * a.n_tiles + blockIdx.y] = sum;
  }
}

// log p(y | x) alone, H = I, diagonal R_t, time-invariant c (base.py:138-166): the only density the aux-Kalman sweep
// evaluates when the prior terms cancel.  One WARP per time step, lane = state entry: consecutive steps are
// consecutive rows of ys / xs / Rs, so every load instruction of a warp covers one contiguous row (the thread-per-step
// mapping of logpdf_fast_kernel read 32 rows 4 bytes at a time: 0.28 ms per call for 378 MB at d = 30).  The step term
// is warp-reduced before the NaN -> 0 rule is applied to it, as the per-step form does.
template <typename T, int NTB>
__global__ void __launch_bounds__(NTB) loglik_rows_kernel(LogpdfArgs<T> a) {
  __shared__ T sred[NTB / 32];
  const int unit = blockIdx.x, c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx, Tn = a.d.T, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t_lo = blockIdx.y * a.tile_len, t_hi = (t_lo + a.tile_len < Tn) ? t_lo + a.tile_len : Tn;
  const T* xg = a.xs + unit_off(a.d.xs, c, b);
  const T* yg = a.ys + unit_off(a.d.ys, c, b);
  const T* rg = a.Rs + unit_off(a.d.Rs, c, b);
  T cc[2] = {T(0), T(0)};  // dx <= 64: at most two entries per lane
  if (a.cs != nullptr) {
    const T* cg = a.cs + unit_off(a.d.cs, c, b);
    if (lane < dx) cc[0] = cg[lane];
    if (lane + 32 < dx) cc[1] = cg[lane + 32];
  }
  T total = T(0);
#pragma unroll 4
  for (int t = t_lo + w; t < t_hi; t += NTB / 32) {
    T term = T(0);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = lane + 32 * q;
      if (i < dx) {
        const T rr = rg[(int64_t)t * a.d.Rs.time + i];
        const T y = yg[(int64_t)t * a.d.ys.time + i], x = xg[(int64_t)t * a.d.xs.time + i];
        const T lii = sqrt(rr);  // chol of the diagonal matrix; NaN / inf entries follow mvn/base.py:49-57
        const T e = (y - x - cc[q]) / lii;
        if (finite_(lii) || dx == 1) term += T(-0.5) * e * e - log(lii) - half_log_2pi_<T>();  // dx = 1: norm.logpdf
        else term += T(-0.5) * e * e;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) term += __shfl_xor_sync(0xffffffffu, term, off);
    if (!(term == term)) term = T(0);
    total += term;
  }
  if (lane == 0) sred[w] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    T sum = T(0);
    for (int q = 0; q < NTB / 32; ++q) sum += sred[q];
    a.partial[(int64_t)unit * a.n_tiles + blockIdx.y] = sum;
  }
}

template <typename T>
__global__ void logpdf_reduce_kernel(const T* partial, const T* ell, T* out, int n_chains, int per_chain,
                                     int minus_ell) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  T s = T(0);
  for (int k = 0; k < per_chain; ++k) s += partial[(int64_t)c * per_chain + k];
  if (minus_ell) s -= ell[c];
  out[c] = s;
}

struct LogpdfPlan {
  int n_tiles, tile_len;
  bool fast;
  size_t off_prep, total;
};

constexpr int kFastThreads = 128;

static bool logpdf_fast_ok(const auxssm_lgssm_desc* d, size_t es, uint32_t which) {
  const int dmax = d->dx > d->dy ? d->dx : d->dy;
  if (dmax > 64 || (dmax > 32 && es > 4)) return false;  // d = 64: fp32 only (F and chol(Q) in 32 KB static smem)
  if ((which & AUXSSM_LOGPDF_PRIOR) && d->T > 1 && (d->Fs.time != 0 || d->Qs.time != 0 || d->bs.time != 0))
    return false;
  if (which & AUXSSM_LOGPDF_LOGLIK) {
    if (!(d->flags & AUXSSM_FLAG_H_IDENTITY) || !(d->flags & AUXSSM_FLAG_R_DIAGONAL)) return false;
    if (d->cs.time != 0) return false;
  }
  return true;
}

static LogpdfPlan make_logpdf_plan(const auxssm_lgssm_desc* d, size_t es, uint32_t which) {
  LogpdfPlan p;
  const long units = (long)d->n_chains * d->n_batch;
  p.fast = logpdf_fast_ok(d, es, which);
  if (p.fast) {
    p.tile_len = kFastThreads;
    p.n_tiles = (d->T + kFastThreads - 1) / kFastThreads;
    p.off_prep = ((size_t)units * p.n_tiles * es + 255) / 256 * 256;
    p.total = p.off_prep + ((size_t)units * ((size_t)d->dx * d->dx + d->dx + 2) * es + 255) / 256 * 256;
    return p;
  }
  p.off_prep = 0;
  long want = (8L * device_sm_count() + units - 1) / units;
  long cap = d->T / 8;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  p.tile_len = (int)((d->T + want - 1) / want);
  p.n_tiles = (d->T + p.tile_len - 1) / p.tile_len;
  p.total = ((size_t)units * p.n_tiles * es + 255) / 256 * 256;
  return p;
}

template <typename T, int D> static int launch_logpdf_fast(cudaStream_t st, const LogpdfArgs<T>& a, T* prep) {
  const int units = a.d.n_chains * a.d.n_batch;
  LogpdfFastArgs<T> fa;
  fa.a = a;
  fa.prep = prep;
  if (a.which & AUXSSM_LOGPDF_PRIOR) {
    const size_t smem = ((size_t)D * Tile<D>::LD + 2 * D + CholWs<D>::WORDS) * sizeof(T);
    logpdf_prep_kernel<T, D><<<units, Tile<D>::NT, smem, st>>>(fa);
    AUXSSM_CHECK_LAUNCH("auxssm_lgssm_logpdf (prep)");
  }
  logpdf_fast_kernel<T, D, kFastThreads><<<dim3(units, a.n_tiles), kFastThreads, 0, st>>>(fa);
  AUXSSM_CHECK_LAUNCH("auxssm_lgssm_logpdf (fast)");
  return AUXSSM_OK;
}

template <typename T, int D> static int launch_logpdf(cudaStream_t st, const LogpdfArgs<T>& a) {
  const int units = a.d.n_chains * a.d.n_batch;
  const size_t smem = logpdf_smem_elems<D>() * sizeof(T);
  int rc = ensure_smem(lgssm_logpdf_kernel<T, D>, smem, "auxssm_lgssm_logpdf");
  if (rc) return rc;
  lgssm_logpdf_kernel<T, D><<<dim3(units, a.n_tiles), Tile<D>::NT, smem, st>>>(a);
  AUXSSM_CHECK_LAUNCH("auxssm_lgssm_logpdf");
  return AUXSSM_OK;
}

template <typename T>
static int logpdf_entry(void* stream, const auxssm_lgssm_desc* d, uint32_t which, const T* ys, const T* xs,
                        const T* ell, const T* m0, const T* P0, const T* Fs, const T* Qs, const T* bs, const T* Hs,
                        const T* Rs, const T* cs, T* out, void* scratch, size_t scratch_bytes) {
  const char* who = "auxssm_lgssm_logpdf";
  int rc = check_desc(d, who);
  if (rc) return rc;
  const bool hid = (d->flags & AUXSSM_FLAG_H_IDENTITY) != 0;
  const bool do_prior = (which & AUXSSM_LOGPDF_PRIOR) != 0, do_lik = (which & AUXSSM_LOGPDF_LOGLIK) != 0;
  if (!xs || !out || (do_prior && (!m0 || !P0 || (d->T > 1 && (!Fs || !Qs || !bs)))) ||
      (do_lik && (!ys || !Rs || (!hid && !Hs))) || ((which & AUXSSM_LOGPDF_MINUS_ELL) && !ell)) {
    set_error("%s: null data pointer", who);
    return AUXSSM_ERR_INVALID;
  }
  const LogpdfPlan p = make_logpdf_plan(d, sizeof(T), which);
  if (scratch_bytes < p.total || scratch == nullptr) {
    set_error("%s: scratch too small (%zu < %zu bytes)", who, scratch_bytes, p.total);
    return AUXSSM_ERR_SCRATCH;
  }
  LogpdfArgs<T> a;
  a.d = *d;
  a.which = which;
  a.ys = ys; a.xs = xs; a.ell = ell; a.m0 = m0; a.P0 = P0; a.Fs = Fs; a.Qs = Qs; a.bs = bs; a.Hs = Hs; a.Rs = Rs;
  a.cs = cs; a.out = out;
  a.partial = static_cast<T*>(scratch);
  a.n_tiles = p.n_tiles;
  a.tile_len = p.tile_len;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dmax = d->dx > d->dy ? d->dx : d->dy;
  T* prep = reinterpret_cast<T*>(static_cast<unsigned char*>(scratch) + p.off_prep);
  if (p.fast && !do_prior && do_lik && dmax <= 64) {
    loglik_rows_kernel<T, kFastThreads><<<dim3(d->n_chains * d->n_batch, a.n_tiles), kFastThreads, 0, st>>>(a);
    AUXSSM_CHECK_LAUNCH("auxssm_lgssm_logpdf (likelihood rows)");
  } else
  switch (padded_dim(dmax) * (p.fast ? -1 : 1)) {
    case -4: rc = launch_logpdf_fast<T, 4>(st, a, prep); break;
    case -8: rc = launch_logpdf_fast<T, 8>(st, a, prep); break;
    case -16: rc = launch_logpdf_fast<T, 16>(st, a, prep); break;
    case -32: rc = launch_logpdf_fast<T, 32>(st, a, prep); break;
    case -64:
      if constexpr (sizeof(T) == 4) { rc = launch_logpdf_fast<T, 64>(st, a, prep); break; }
      set_error("%s: internal dispatch error", who); return AUXSSM_ERR_UNSUPPORTED;
    case 4: rc = launch_logpdf<T, 4>(st, a); break;
    case 8: rc = launch_logpdf<T, 8>(st, a); break;
    case 16: rc = launch_logpdf<T, 16>(st, a); break;
    case 32: rc = launch_logpdf<T, 32>(st, a); break;
    case 64: rc = launch_logpdf<T, 64>(st, a); break;
    default: set_error("%s: unsupported dimension %d", who, dmax); return AUXSSM_ERR_UNSUPPORTED;
  }
  if (rc) return rc;
  logpdf_reduce_kernel<T><<<(d->n_chains + 127) / 128, 128, 0, st>>>(
      a.partial, ell, out, d->n_chains, d->n_batch * p.n_tiles, (which & AUXSSM_LOGPDF_MINUS_ELL) ? 1 : 0);
  AUXSSM_CHECK_LAUNCH("auxssm_lgssm_logpdf (reduce)");
  return AUXSSM_OK;
}

}  // namespace auxssm

// This translation unit carries the fp32 entry point (and the dtype-independent scratch query); the fp64
// instantiations are compiled by lgssm_logpdf_f64.cu, which includes this file with AUXSSM_ONLY_F64 defined, so
// that the two halves build in parallel.
#ifndef AUXSSM_ONLY_F64
extern "C" __attribute__((visibility("default"))) size_t auxssm_lgssm_logpdf_scratch_bytes(const auxssm_lgssm_desc* desc, int32_t elem_size) {
  if (auxssm::check_desc(desc, "auxssm_lgssm_logpdf_scratch_bytes")) return 0;
  const size_t a = auxssm::make_logpdf_plan(desc, (size_t)elem_size, AUXSSM_LOGPDF_POSTERIOR).total;
  const size_t b = auxssm::make_logpdf_plan(desc, (size_t)elem_size, AUXSSM_LOGPDF_PRIOR).total;
  const size_t c = auxssm::make_logpdf_plan(desc, (size_t)elem_size, AUXSSM_LOGPDF_LOGLIK).total;
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_lgssm_logpdf_f32(void* stream, const auxssm_lgssm_desc* desc, uint32_t which,
                                           const float* ys, const float* xs, const float* ell, const float* m0,
                                           const float* P0, const float* Fs, const float* Qs, const float* bs,
                                           const float* Hs, const float* Rs, const float* cs, float* out,
                                           void* scratch, size_t scratch_bytes) {
  return auxssm::logpdf_entry<float>(stream, desc, which, ys, xs, ell, m0, P0, Fs, Qs, bs, Hs, Rs, cs, out, scratch,
                                     scratch_bytes);
}
#else
extern "C" __attribute__((visibility("default"))) int32_t auxssm_lgssm_logpdf_f64(void* stream, const auxssm_lgssm_desc* desc, uint32_t which,
                                           const double* ys, const double* xs, const double* ell, const double* m0,
                                           const double* P0, const double* Fs, const double* Qs, const double* bs,
                                           const double* Hs, const double* Rs, const double* cs, double* out,
                                           void* scratch, size_t scratch_bytes) {
  return auxssm::logpdf_entry<double>(stream, desc, which, ys, xs, ell, m0, P0, Fs, Qs, bs, Hs, Rs, cs, out, scratch,
                                      scratch_bytes);
}
#endif

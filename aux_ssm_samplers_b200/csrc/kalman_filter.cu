// Kalman filtering for (batches of) linear-Gaussian state-space models.
//
// Replaces aux_samplers/_primitives/kalman/filtering.py of the reference:
//   filtering:18-46, _sequential_filtering:66-79, sequential_update:84-130,
//   sequential_predict:135-139, _parallel_filtering:49-63, _filtering_init_one:197-250,
//   _filtering_op_impl:164-183.
//
// Layout on the GPU: one CTA per (chain, inner-batch unit, time chunk); every d x d operand
// lives in shared memory (see smem_linalg.cuh).
//
//  * parallel == 0  (or one time chunk): the CTA runs the predict/update recursion over all T.
//  * parallel == 1 with n > 1 time chunks -- a chunked prefix scan over the reference's
//    associative 5-tuple elements (A, b, C, eta, J):
//      pass A  chunk 0 runs the recursion directly (its aggregate's (b, C) IS the filtering
//              marginal); chunks 1..n-2 build their elements (_filtering_init_one) and left-fold
//              them with the reference's combine operator (_filtering_op_impl);
//      pass B  one CTA per unit scans the (b, C) prefix across chunk aggregates;
//      pass C  chunks 1..n-1 run the recursion from their incoming prefix, which also yields
//              the log-likelihood increments the reference obtains from a second vmapped
//              sequential_predict_update pass (filtering.py:60-62).
//
// Missing observations (non-finite y entries) are handled by the mathematical row mask the
// reference emulates with infinite variances (filtering.py:86-101): masked rows of H, c are
// zeroed, masked rows/cols of R become unit rows, and masked dimensions are excluded from the
// Gaussian normaliser; a fully missing y_t is a pure prediction step (filtering.py:127-130).
#include "common.cuh"
#include "smem_linalg.cuh"

namespace auxssm {

template <typename T> struct FilterArgs {
  auxssm_lgssm_desc d;
  const T *ys, *m0, *P0, *Fs, *Qs, *bs, *Hs, *Rs, *cs;
  T *ms, *Ps;
  T* ellpart;  // [units][n_chunks]
  T* prefix;   // [units][n_chunks][dx + dx*dx]   filtering marginal at the END of chunk j
  T* agg;      // [units][n_chunks][3 dx^2 + 2 dx] chunk aggregates (A, C, J, b, eta)
  int* gate;   // number of non-zero off-diagonal entries of F (common.cuh); nullptr when no hint was given
  int n_chunks, chunk_len;
};

template <typename T, int D> struct Carve {
  T* p;
  __device__ T* mat() { T* r = p; p += D * Tile<D>::LD; return r; }
  __device__ T* vec() { T* r = p; p += D; return r; }
};

template <typename T, int D> struct StepBuf {
  T *P, *W, *S, *U, *Ft, *Q, *Ht, *R;
  T *m, *bv, *y, *c, *v, *dg;
  T* rd;  // diag(R_t) when AUXSSM_FLAG_R_DIAGONAL
  int* mask;
  T* red;
  bool h_clean, r_clean, c_clean;
};

template <typename T> __device__ __forceinline__ T half_log_2pi() { return T(0.91893853320467274178); }

// ---- staging ---------------------------------------------------------------------------------

template <typename T, int D>
__device__ __forceinline__ void stage_dyn(const FilterArgs<T>& a, int c, int b, int tm1, bool first, T* Ft, T* Fp,
                                          T* Q, T* bv) {
  const int dx = a.d.dx;
  cta_sync<D>();
  if (first || a.d.Fs.time != 0) {
    const T* g = a.Fs + unit_off(a.d.Fs, c, b) + (int64_t)tm1 * a.d.Fs.time;
    if (Ft) load_mat<T, D, true>(g, Ft, dx, dx);
    if (Fp) load_mat<T, D, false>(g, Fp, dx, dx);
  }
  if (first || a.d.Qs.time != 0)
    load_mat<T, D, false>(a.Qs + unit_off(a.d.Qs, c, b) + (int64_t)tm1 * a.d.Qs.time, Q, dx, dx);
  if (first || a.d.bs.time != 0) {
    const T* g = a.bs + unit_off(a.d.bs, c, b) + (int64_t)tm1 * a.d.bs.time;
    if (threadIdx.x < dx) bv[threadIdx.x] = g[threadIdx.x];
  }
  cta_sync<D>();
}

// Stages y_t, H_t (transposed into Ht and/or plain into Hp), R_t, c_t with the missing-row mask
// applied.  Returns the number of observed (finite) entries of y_t.
template <typename T, int D>
__device__ __forceinline__ int stage_obs(const FilterArgs<T>& a, int c, int b, int t, bool first, StepBuf<T, D>& s,
                                         T* Hp, bool explicit_identity) {
  constexpr int LD = Tile<D>::LD;
  const int dx = a.d.dx, dy = a.d.dy;
  const bool hid = (a.d.flags & AUXSSM_FLAG_H_IDENTITY) != 0;
  cta_sync<D>();
  if (threadIdx.x < dy) {
    const T yv = a.ys[unit_off(a.d.ys, c, b) + (int64_t)t * a.d.ys.time + threadIdx.x];
    s.y[threadIdx.x] = yv;
    s.mask[threadIdx.x] = finite_(yv) ? 0 : 1;
  }
  cta_sync<D>();
  int n_mask = 0;
  for (int i = 0; i < dy; ++i) n_mask += s.mask[i];
  const bool any_mask = n_mask > 0;
  if (first || any_mask || !s.h_clean || (!hid && a.d.Hs.time != 0)) {
    if (hid) {
      if (explicit_identity) {
        AUXSSM_FOR_ROWS(i, j, dy, dx) {
          const T v = (i == j && !s.mask[i]) ? T(1) : T(0);
          if (s.Ht) s.Ht[j * LD + i] = v;
          if (Hp) Hp[i * LD + j] = v;
        }
      }
    } else {
      const T* g = a.Hs + unit_off(a.d.Hs, c, b) + (int64_t)t * a.d.Hs.time;
      AUXSSM_FOR_ROWS(i, j, dy, dx) {
        const T v = s.mask[i] ? T(0) : g[i * dx + j];
        if (s.Ht) s.Ht[j * LD + i] = v;
        if (Hp) Hp[i * LD + j] = v;
      }
    }
    s.h_clean = !any_mask;
  }
  const bool rdiag = (a.d.flags & AUXSSM_FLAG_R_DIAGONAL) != 0;
  if (rdiag && (first || any_mask || !s.r_clean || a.d.Rs.time != 0)) {
    const T* g = a.Rs + unit_off(a.d.Rs, c, b) + (int64_t)t * a.d.Rs.time;
    if (threadIdx.x < dy) s.rd[threadIdx.x] = s.mask[threadIdx.x] ? T(1) : g[threadIdx.x];
    if (s.R != nullptr) {
      cta_sync<D>();
      AUXSSM_FOR_ROWS(i, j, dy, dy) s.R[i * LD + j] = (i == j) ? s.rd[i] : T(0);
    }
    s.r_clean = !any_mask;
  }
  if (!rdiag && s.R != nullptr && (first || any_mask || !s.r_clean || a.d.Rs.time != 0)) {
    const T* g = a.Rs + unit_off(a.d.Rs, c, b) + (int64_t)t * a.d.Rs.time;
    AUXSSM_FOR_ROWS(i, j, dy, dy) {
      T v = g[i * dy + j];
      if (s.mask[i] || s.mask[j]) v = (i == j) ? T(1) : T(0);
      s.R[i * LD + j] = v;
    }
    s.r_clean = !any_mask;
  }
  if (first || any_mask || !s.c_clean || a.d.cs.time != 0) {
    if (threadIdx.x < dy) {
      T v = T(0);
      if (a.cs != nullptr) v = a.cs[unit_off(a.d.cs, c, b) + (int64_t)t * a.d.cs.time + threadIdx.x];
      s.c[threadIdx.x] = s.mask[threadIdx.x] ? T(0) : v;
    }
    s.c_clean = !any_mask;
  }
  cta_sync<D>();
  return dy - n_mask;
}

// ---- the sequential recursion ------------------------------------------------------------------

// (m, P) <- (F m + b, F P F^T + Q), symmetrised.  filtering.py:135-139.
template <typename T, int D> __device__ __forceinline__ void kalman_predict(StepBuf<T, D>& s, int dx) {
  mm_tn<T, D, 0>(s.P, s.Ft, s.W, nullptr, dx);   // W = P F^T   (P symmetric)
  mm_tn<T, D, 1>(s.Ft, s.W, s.P, s.Q, dx);       // P = F W + Q
  mv_tn<T, D, 1>(s.Ft, s.m, s.m, s.bv, dx, dx);  // m = F m + b
}

// Measurement update with the masked observation staged in s; returns the log-likelihood
// increment (uniform over the CTA).  filtering.py:84-130, with  Z = L^{-1} H P  so that
// G S G^T = Z^T Z  and  G (y - yhat) = Z^T L^{-1} (y - yhat).
// R_t kept in registers with the row mapping of AUXSSM_FOR_ROWS (H = I path): element q of thread
// (warp w, lane) is R[w + (q / CPL) * NW][lane + 32 * (q % CPL)].
template <int D> struct RowRegs {
  static constexpr int NW = Tile<D>::NT / 32, RPW = (D + NW - 1) / NW, CPL = (D + 31) / 32, N = RPW * CPL;
};

template <typename T, int D>
__device__ __forceinline__ void load_R_regs(const FilterArgs<T>& a, int c, int b, int t, T (&Rr)[RowRegs<D>::N]) {
  using RR = RowRegs<D>;
  const int dy = a.d.dy, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* __restrict__ g = a.Rs + unit_off(a.d.Rs, c, b) + (int64_t)t * a.d.Rs.time;
#pragma unroll
  for (int qi = 0; qi < RR::RPW; ++qi) {
    const int i = w + qi * RR::NW;
#pragma unroll
    for (int qj = 0; qj < RR::CPL; ++qj) {
      const int j = lane + 32 * qj;
      Rr[qi * RR::CPL + qj] = (i < dy && j < dy) ? g[i * dy + j] : T(0);
    }
  }
}

// Measurement update with the masked observation staged in s; returns the log-likelihood
// increment (uniform over the CTA).  filtering.py:84-130, with  Z = L^{-1} H P  so that
// G S G^T = Z^T Z  and  G (y - yhat) = Z^T L^{-1} (y - yhat).  With H = I, R_t comes from registers.
template <typename T, int D, bool HID>
__device__ __forceinline__ T kalman_update(StepBuf<T, D>& s, int dx, int dy, int n_obs,
                                           const T (&Rr)[RowRegs<D>::N], bool rdiag) {
  constexpr int LD = Tile<D>::LD;
  using RR = RowRegs<D>;
  const int i = threadIdx.x;
  if (HID) {
    cta_sync<D>();
    if (i < dy) s.v[i] = s.mask[i] ? T(0) : s.y[i] - s.m[i] - s.c[i];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool any_mask = n_obs < dy;
#pragma unroll
    for (int qi = 0; qi < RR::RPW; ++qi) {
      const int r = w + qi * RR::NW;
#pragma unroll
      for (int qj = 0; qj < RR::CPL; ++qj) {
        const int q = lane + 32 * qj;
        if (r < dy && q < dx) {
          const T p = s.P[r * LD + q];
          T u = p, sv = p + (rdiag ? ((r == q) ? s.rd[r] : T(0)) : Rr[qi * RR::CPL + qj]);
          if (any_mask) {
            if (s.mask[r]) u = T(0);
            if (s.mask[r] || s.mask[q]) sv = (r == q) ? T(1) : T(0);
          }
          s.U[r * LD + q] = u;
          s.S[r * LD + q] = sv;
        }
      }
    }
    cta_sync<D>();
  } else {
    mv_tn<T, D, 0>(s.Ht, s.m, s.v, nullptr, dx, dy);  // v = H m
    if (i < dy) s.v[i] = s.mask[i] ? T(0) : s.y[i] - s.v[i] - s.c[i];
    mm_tn<T, D, 0>(s.Ht, s.P, s.U, nullptr, dx);      // U = H P            (dy x dx)
    transpose_to<T, D>(s.U, s.W, dy, dx);             // W = U^T = P H^T    (dx x dy)
    mm_tn<T, D, 1>(s.W, s.Ht, s.S, s.R, dx);          // S = H P H^T + R
  }
  chol_inplace<T, D>(s.S, s.dg, dy);
  fwd_solve<T, D>(s.S, s.U, s.v, dy, dx, s.dg);       // U <- Z = L^{-1} U ; v <- L^{-1} (y - yhat)
  mm_tn<T, D, 2>(s.U, s.U, s.P, s.P, dy);             // P -= Z^T Z
  mv_tn<T, D, 1>(s.U, s.v, s.m, s.m, dy, dx);         // m += Z^T v
  T part = T(0);
  if (i < dy && !s.mask[i]) part = T(-0.5) * s.v[i] * s.v[i] - log(s.dg[i]);
  T ell = cta_sum<T, D>(part, s.red) - T(n_obs) * half_log_2pi<T>();
  if (!(ell == ell)) ell = T(0);  // nan_to_num(ell_inc, nan=0.), filtering.py:125
  return ell;
}

template <typename T, int D> __device__ __forceinline__ void store_state(const FilterArgs<T>& a, int c, int b, int t,
                                                                         const StepBuf<T, D>& s) {
  const int dx = a.d.dx;
  cta_sync<D>();
  if (a.Ps == nullptr) return;  // ell-only call
  T* gm = a.ms + unit_off(a.d.ms, c, b) + (int64_t)t * a.d.ms.time;
  if (threadIdx.x < dx) gm[threadIdx.x] = s.m[threadIdx.x];
  store_mat<T, D>(a.Ps + unit_off(a.d.Ps, c, b) + (int64_t)t * a.d.Ps.time, s.P, dx, dx);
}

template <typename T, int D> __device__ __forceinline__ StepBuf<T, D> carve_step(Carve<T, D>& cv, bool hid) {
  StepBuf<T, D> s;
  s.P = cv.mat(); s.W = cv.mat(); s.S = cv.mat();
  s.Ft = cv.mat(); s.Q = cv.mat();
  s.R = hid ? nullptr : cv.mat();  // with H = I, R_t lives in registers (load_R_regs)
  // with H = I the update never needs W (P F^T is dead after the predict), so U shares its storage
  s.U = hid ? s.W : cv.mat();
  s.Ht = hid ? nullptr : cv.mat();
  s.m = cv.vec(); s.bv = cv.vec(); s.y = cv.vec(); s.c = cv.vec(); s.v = cv.vec(); s.red = cv.vec();
  s.dg = cv.p; cv.p += CholWs<D>::WORDS;
  s.rd = cv.vec();
  s.mask = reinterpret_cast<int*>(cv.p);
  cv.p += (D * sizeof(int) + sizeof(T) - 1) / sizeof(T);
  s.h_clean = s.r_clean = s.c_clean = false;
  return s;
}

template <int D> __host__ __device__ constexpr size_t step_smem_elems(bool hid) {
  return (size_t)(hid ? 5 : 8) * D * Tile<D>::LD + 11 * D + D /* mask, generous */ + 16;
}

// Runs t = t_begin .. t_end-1 of the recursion starting from the state in s (valid for t_begin-1),
// or from the prior (m0, P0) with the t = 0 update when t_begin == 0.  Returns the summed ell.
template <typename T, int D, bool HID>
__device__ __forceinline__ T run_recursion(const FilterArgs<T>& a, int c, int b, int t_begin, int t_end,
                                           StepBuf<T, D>& s) {
  const int dx = a.d.dx, dy = a.d.dy;
  const bool rdiag = (a.d.flags & AUXSSM_FLAG_R_DIAGONAL) != 0;
  T ell = T(0);
  T Rr[RowRegs<D>::N];
#pragma unroll
  for (int q = 0; q < RowRegs<D>::N; ++q) Rr[q] = T(0);
  bool first_dyn = true, first_obs = true;
  int t = t_begin;
  if (t == 0) {
    cta_sync<D>();
    if (threadIdx.x < dx) s.m[threadIdx.x] = a.m0[unit_off(a.d.m0, c, b) + threadIdx.x];
    load_mat<T, D, false>(a.P0 + unit_off(a.d.P0, c, b), s.P, dx, dx);
    if (HID && !rdiag) load_R_regs<T, D>(a, c, b, 0, Rr);
    const int n_obs = stage_obs<T, D>(a, c, b, 0, true, s, nullptr, false);
    first_obs = false;
    if (n_obs > 0) {
      ell += kalman_update<T, D, HID>(s, dx, dy, n_obs, Rr, rdiag);
      symmetrize<T, D>(s.P, dx);
    }
    store_state<T, D>(a, c, b, 0, s);
    t = 1;
  }
  for (; t < t_end; ++t) {
    // R_t is requested first so that its DRAM latency hides behind the two predict products
    if (HID && !rdiag && (first_obs || a.d.Rs.time != 0)) load_R_regs<T, D>(a, c, b, t, Rr);
    stage_dyn<T, D>(a, c, b, t - 1, first_dyn, s.Ft, nullptr, s.Q, s.bv);
    first_dyn = false;
    const int n_obs = stage_obs<T, D>(a, c, b, t, first_obs, s, nullptr, false);
    first_obs = false;
    kalman_predict<T, D>(s, dx);
    if (n_obs > 0) ell += kalman_update<T, D, HID>(s, dx, dy, n_obs, Rr, rdiag);
    symmetrize<T, D>(s.P, dx);
    store_state<T, D>(a, c, b, t, s);
  }
  return ell;
}

// ---- fast recursion for H = I with DIAGONAL R_t -------------------------------------------------------
// With independent scalar observations y_i = x_i + c_i + N(0, r_i) the measurement update of
// filtering.py:84-130 factorises into d rank-one updates (sequential processing of observations):
//     s = P_ii + r_i,  k = P[:, i] / s,  m += k (y_i - m_i - c_i),  P -= k k^T s,
//     ell += log N(y_i; m_i + c_i, s)
// which is algebraically the same posterior and the same joint log-likelihood (chain rule) without
// any Cholesky factor or triangular solve: d^3 FMAs, no sqrt/divide chain.  Each thread keeps its
// TM x TN tile of P in registers across the whole time loop; row i of the current P is exchanged
// through a double-buffered shared vector (one barrier per scalar update).  Missing entries of y_t are
// simply skipped.  When F is time-invariant and diagonal (AR(1)-type dynamics, e.g. the stochastic
// volatility model) the predict step is done on the register tile as well: P <- f_a f_b P_ab + Q_ab.
template <typename T, int D>
__device__ __forceinline__ T run_recursion_scalar(const FilterArgs<T>& a, int c, int b, int t_begin, int t_end,
                                                  StepBuf<T, D>& s) {
  using Tl = Tile<D>;
  constexpr int LD = Tl::LD, TM = Tl::TM, TN = Tl::TN, NT = Tl::NT, TX = D / TN, TILES = (D / TM) * TX;
  const int d = a.d.dx;
  const int tile = threadIdx.x;
  const bool has = tile < TILES;
  const int ty = has ? tile / TX : 0, tx = has ? tile % TX : 0;
  T* rowbuf = s.W;          // 2 x (D + 4) words of the (otherwise unused) W buffer
  T* svbuf = s.S;           // per-dimension innovation variance and squared innovation (S is unused here)
  T* i2buf = s.S + D;
  T p[TM][TN], mreg[TM];
  T ell = T(0);
  int parity = 0;

  auto tile_load = [&]() {
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      if (has) ldv<T, TN>(p[r], s.P + (ty * TM + r) * LD + tx * TN);
      mreg[r] = (has && tx == 0) ? s.m[ty * TM + r] : T(0);
    }
  };
  auto tile_store = [&]() {
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      if (has) stv<T, TN>(s.P + (ty * TM + r) * LD + tx * TN, p[r]);
      if (has && tx == 0) s.m[ty * TM + r] = mreg[r];
    }
  };
  // all d scalar updates of one time step; returns the log-likelihood increment
  auto scalar_updates = [&]() -> T {
    for (int i0 = 0; i0 < d; i0 += TM) {
#pragma unroll
      for (int q = 0; q < TM; ++q) {
        const int i = i0 + q;
        if (i < d && !s.mask[i]) {  // uniform over the CTA
          T* rb = rowbuf + parity * (D + 4);
          parity ^= 1;
          if (has && ty == i0 / TM) {
            stv<T, TN>(rb + tx * TN, p[q]);
            if (tx == 0) rb[D] = mreg[q];
          }
          cta_sync<D>();
          const T sv = rb[i] + s.rd[i];
          const T inv = T(1) / sv;
          const T innov = s.y[i] - rb[D] - s.c[i];
          if (has) {
            T ra[TM], rc[TN];
            ldv<T, TM>(ra, rb + ty * TM);
            ldv<T, TN>(rc, rb + tx * TN);
#pragma unroll
            for (int r = 0; r < TM; ++r) {
              const T kr = ra[r] * inv;
#pragma unroll
              for (int cc = 0; cc < TN; ++cc) p[r][cc] = fma_(-kr, rc[cc], p[r][cc]);
              if (tx == 0) mreg[r] = fma_(kr, innov, mreg[r]);
            }
          }
          if (threadIdx.x == 0) { svbuf[i] = sv; i2buf[i] = innov * innov * inv; }
        }
      }
    }
    cta_sync<D>();
    T part = T(0);
    const int i = threadIdx.x;
    if (i < d && !s.mask[i]) part = T(-0.5) * i2buf[i] - T(0.5) * log(svbuf[i]) - half_log_2pi<T>();
    T e = cta_sum<T, D>(part, s.red);
    if (!(e == e)) e = T(0);
    return e;
  };

  // is F time-invariant and diagonal?  (decided once per CTA)
  bool fdiag = false;
  int t = t_begin;
  bool first_dyn = true, first_obs = true;
  if (t == 0) {
    cta_sync<D>();
    if (threadIdx.x < d) s.m[threadIdx.x] = a.m0[unit_off(a.d.m0, c, b) + threadIdx.x];
    load_mat<T, D, false>(a.P0 + unit_off(a.d.P0, c, b), s.P, d, d);
    const int n_obs = stage_obs<T, D>(a, c, b, 0, true, s, nullptr, false);
    first_obs = false;
    tile_load();
    if (n_obs > 0) ell += scalar_updates();
    tile_store();
    store_state<T, D>(a, c, b, 0, s);
    t = 1;
  } else {
    cta_sync<D>();
    tile_load();
  }
  for (; t < t_end; ++t) {
    stage_dyn<T, D>(a, c, b, t - 1, first_dyn, s.Ft, nullptr, s.Q, s.bv);
    if (first_dyn) {
      int off = 0;
      if (a.d.Fs.time == 0) {
        AUXSSM_FOR_ROWS(ri, rj, d, d) if (ri != rj && s.Ft[ri * LD + rj] != T(0)) off = 1;
      } else {
        off = 1;
      }
      fdiag = (Tl::NT == 32) ? (__any_sync(0xffffffffu, off) == 0) : (__syncthreads_or(off) == 0);
    }
    first_dyn = false;
    const int n_obs = stage_obs<T, D>(a, c, b, t, first_obs, s, nullptr, false);
    first_obs = false;
    if (fdiag) {
      // predict on the register tile: P_ab <- f_a f_b P_ab + Q_ab ; m_a <- f_a m_a + b_a
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        if (has) {
          const int ra = ty * TM + r;
          const T fa = s.Ft[ra * LD + ra];
          T qv[TN];
          ldv<T, TN>(qv, s.Q + ra * LD + tx * TN);
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) {
            const int cb = tx * TN + cc;
            p[r][cc] = fma_(fa * s.Ft[cb * LD + cb], p[r][cc], qv[cc]);
          }
          if (tx == 0) mreg[r] = fma_(fa, mreg[r], s.bv[ra]);
        }
      }
    } else {
      tile_store();
      kalman_predict<T, D>(s, d);   // dense F: two products in shared memory (uses W)
      symmetrize<T, D>(s.P, d);
      tile_load();
      cta_sync<D>();
    }
    if (n_obs > 0) ell += scalar_updates();
    cta_sync<D>();
    tile_store();
    store_state<T, D>(a, c, b, t, s);
  }
  cta_sync<D>();
  return ell;
}

__device__ __forceinline__ void chunk_range(int j, int chunk_len, int T, int& t0, int& t1) {
  // chunk j covers scan elements t in [1 + j*len, 1 + (j+1)*len) ; chunk 0 additionally owns t = 0
  t0 = (j == 0) ? 0 : 1 + j * chunk_len;
  t1 = 1 + (j + 1) * chunk_len;
  if (t1 > T) t1 = T;
}

// pass A (chunk 0) and pass C (chunks >= 1): the recursion over one chunk.
template <typename T, int D, bool HID>
__global__ void __launch_bounds__(Tile<D>::NT, Occ<T, D>::MINB) filter_recursion_kernel(FilterArgs<T> a, int chunk_first) {
  if (a.gate != nullptr && *a.gate == 0) return;  // the hint held: the structured kernel does this launch's work
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Carve<T, D> cv{reinterpret_cast<T*>(smem_raw)};
  constexpr int NT = Tile<D>::NT;
  const size_t total = step_smem_elems<D>(HID);
  for (size_t e = threadIdx.x; e < total; e += NT) cv.p[e] = T(0);
  __syncthreads();
  StepBuf<T, D> s = carve_step<T, D>(cv, HID);
  const int j = chunk_first + blockIdx.y;
  const int unit = blockIdx.x;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx;
  int t0, t1;
  chunk_range(j, a.chunk_len, a.d.T, t0, t1);
  const int psz = dx + dx * dx;
  if (j > 0) {
    const T* pf = a.prefix + ((int64_t)unit * a.n_chunks + (j - 1)) * psz;
    if (threadIdx.x < dx) s.m[threadIdx.x] = pf[threadIdx.x];
    load_mat<T, D, false>(pf + dx, s.P, dx, dx);
  }
  const bool scalar_path = HID && (a.d.flags & AUXSSM_FLAG_R_DIAGONAL) != 0;
  const T ell = scalar_path ? run_recursion_scalar<T, D>(a, c, b, t0, t1, s) : run_recursion<T, D, HID>(a, c, b, t0, t1, s);
  cta_sync<D>();
  if (j + 1 < a.n_chunks && j == 0) {
    T* pf = a.prefix + ((int64_t)unit * a.n_chunks + j) * psz;
    if (threadIdx.x < dx) pf[threadIdx.x] = s.m[threadIdx.x];
    store_mat<T, D>(pf + dx, s.P, dx, dx);
  }
  if (threadIdx.x == 0) a.ellpart[(int64_t)unit * a.n_chunks + j] = ell;
}

// ---- one warp per chain: H = I, diagonal R_t, time-invariant diagonal F, Q, b ---------------------------
// The structured case of every auxiliary Kalman sampler whose latent dynamics are AR(1)-type (e.g. the
// stochastic-volatility model).  The whole recursion lives in the registers of ONE warp: lane (ty, tx) of
// a 4 x 8 lane grid owns a (D/4) x (D/8) tile of P, of Q and of f f^T; the predict step is 1 FMA per
// element, the measurement update is d rank-one updates (see run_recursion_scalar) with row i exchanged
// through a double-buffered shared vector and __syncwarp only (no block barrier anywhere).  y_{t+1} and
// diag(R_{t+1}) are prefetched into registers one step ahead so DRAM latency never sits on the critical
// path; P_t / m_t are stored straight from the register tile (a lane's four columns of a row are 16 contiguous bytes of
// the compact row-major P_t); only odd d / unaligned bases and the chunk hand-over go through the shared-memory stage.
template <typename T, int D>
__global__ void __launch_bounds__(128, 2)
filter_warp_kernel(FilterArgs<T> a, int chunk_first) {
  // Four chains per CTA, one per warp: warps are bound to the four SM sub-partitions by their index
  // inside the CTA, so single-warp CTAs would all queue on one scheduler.
  constexpr int GY = 4, GX = 8, TM = D / GY, TN = D / GX, LDS_ = D + 4, WPC = 4;
  __shared__ __align__(16) T stage_[WPC][D * LDS_];
  __shared__ __align__(16) T ev_[WPC][D];
  __shared__ __align__(16) T rdv_[WPC][D];
  __shared__ T svbuf_[WPC][D], mvec_[WPC][D];
  __shared__ __align__(16) T zbuf_[WPC][4 * (D + 4)];  // Z = L_b^{-1} P[I, :] of the current block (see below)
  const int wq = threadIdx.x >> 5;
  T* stage = stage_[wq];
  T* rowbuf = stage_[wq];  // the row-exchange buffers alias the output stage (disjoint in time, __syncwarp between)
  T *ev = ev_[wq], *rdv = rdv_[wq], *svbuf = svbuf_[wq], *mvec = mvec_[wq];
  const int lane = threadIdx.x & 31, ty = lane / GX, tx = lane % GX;
  const int j = chunk_first + blockIdx.y, unit = blockIdx.x * WPC + wq;
  if (unit >= a.d.n_chains * a.d.n_batch) return;  // whole warp; only __syncwarp is used below
  if (a.gate != nullptr && *a.gate != 0) return;   // F is not diagonal after all: the general kernel runs instead
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int d = a.d.dx, Tn = a.d.T;
  int t0, t1;
  chunk_range(j, a.chunk_len, Tn, t0, t1);

  // time-invariant parameters -> registers
  const T* Fg = a.Fs ? a.Fs + unit_off(a.d.Fs, c, b) : nullptr;
  const T* Qg = a.Qs ? a.Qs + unit_off(a.d.Qs, c, b) : nullptr;
  const T* bg = a.bs ? a.bs + unit_off(a.d.bs, c, b) : nullptr;
  T p[TM][TN], q[TM][TN], mreg[TM], fa[TM], fc[TN], ba[TM];
  // MROW (the blocked exchange path, see EXCH below): the filtered mean lives ONE ENTRY PER LANE (lane = row) and is
  // updated from the shared image of Z with 4 FMAs per block; the register-tile form below it (every lane carrying the
  // TM entries of its row group, only the tx = 0 copies ever read) cost 4 TM FMAs per block, 17 % of the block.
  constexpr bool MROW = (TM % 4 == 0) && (TM % 2 == 0) && (TN % 2 == 0) && (TM > TN) && (D <= 32);
  T m_own = T(0), f_own = T(0), b_own = T(0);
#pragma unroll
  for (int cc = 0; cc < TN; ++cc) {
    const int cb = tx * TN + cc;
    fc[cc] = (Fg && cb < d) ? Fg[cb * d + cb] : T(0);
  }
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int ra = ty * TM + r;
    fa[r] = (Fg && ra < d) ? Fg[ra * d + ra] : T(0);
    ba[r] = (bg && ra < d) ? bg[ra] : T(0);
#pragma unroll
    for (int cc = 0; cc < TN; ++cc) {
      const int cb = tx * TN + cc;
      const bool in = ra < d && cb < d;
      q[r][cc] = (Qg && in) ? Qg[ra * d + cb] : T(0);
    }
  }
  // initial state: prior (chunk 0) or the incoming prefix
  const int psz = d + d * d;
  const T* m_src = (j == 0) ? a.m0 + unit_off(a.d.m0, c, b) : a.prefix + ((int64_t)unit * a.n_chunks + (j - 1)) * psz;
  const T* P_src = (j == 0) ? a.P0 + unit_off(a.d.P0, c, b) : m_src + d;
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int ra = ty * TM + r;
    mreg[r] = (!MROW && tx == 0 && ra < d) ? m_src[ra] : T(0);
#pragma unroll
    for (int cc = 0; cc < TN; ++cc) {
      const int cb = tx * TN + cc;
      p[r][cc] = (ra < d && cb < d) ? P_src[ra * d + cb] : T(0);
    }
  }
  if (MROW && lane < d) {
    m_own = m_src[lane];
    f_own = Fg ? Fg[lane * d + lane] : T(0);
    b_own = bg ? bg[lane] : T(0);
  }
  const T* yg = a.ys + unit_off(a.d.ys, c, b);
  const T* rg = a.Rs + unit_off(a.d.Rs, c, b);
  const T* cg = a.cs ? a.cs + unit_off(a.d.cs, c, b) : nullptr;
  const bool want_out = a.Ps != nullptr;
  T* msg = want_out ? a.ms + unit_off(a.d.ms, c, b) : nullptr;
  T* Psg = want_out ? a.Ps + unit_off(a.d.Ps, c, b) : nullptr;
  const bool pairs_ok = (d & 1) == 0 && (a.d.Ps.time & 1) == 0 &&
                        (reinterpret_cast<uintptr_t>(Psg) % (2 * sizeof(T))) == 0;
  // prefetch the first step's observation
  T y_n = T(0), r_n = T(1), c_n = T(0);
  if (lane < d) {
    y_n = yg[(int64_t)t0 * a.d.ys.time + lane];
    r_n = rg[(int64_t)t0 * a.d.Rs.time + lane];
    if (cg) c_n = cg[(int64_t)t0 * a.d.cs.time + lane];
  }
  if (D > 32) { /* lanes cover at most 32 dimensions: this kernel is instantiated for D <= 32 only */ }
  T ell = T(0);
  int parity = 0;
  for (int t = t0; t < t1; ++t) {
    // ---- observation of this step -> shared; prefetch the next one ----
    const T y_c = y_n, r_c = r_n, c_c = c_n;
    if (t + 1 < t1 && lane < d) {
      y_n = yg[(int64_t)(t + 1) * a.d.ys.time + lane];
      r_n = rg[(int64_t)(t + 1) * a.d.Rs.time + lane];
      if (cg) c_n = cg[(int64_t)(t + 1) * a.d.cs.time + lane];
    }
    const bool obs = (lane < d) && finite_(y_c);
    const unsigned obs_mask = __ballot_sync(0xffffffffu, obs);
    if (lane < D) { ev[lane] = y_c - c_c; rdv[lane] = r_c; }
    // ---- predict (not for the very first time step of the model) ----
    if (t > 0) {
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        if constexpr (TN % 2 == 0) {
#pragma unroll
          for (int c2 = 0; c2 < TN / 2; ++c2) {  // packed pairs: two FMAs per issue slot (smem_linalg.cuh)
            const Pair<T> ff = pair_mul_s(fa[r], pair_make(fc[2 * c2], fc[2 * c2 + 1]));
            const Pair<T> pp = pair_fma(ff, pair_make(p[r][2 * c2], p[r][2 * c2 + 1]),
                                        pair_make(q[r][2 * c2], q[r][2 * c2 + 1]));
            pair_get(pp, p[r][2 * c2], p[r][2 * c2 + 1]);
          }
        } else {
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) p[r][cc] = fma_(fa[r] * fc[cc], p[r][cc], q[r][cc]);
        }
        if (!MROW) mreg[r] = fma_(fa[r], mreg[r], ba[r]);
      }
      if (MROW) m_own = fma_(f_own, m_own, b_own);
    }
    __syncwarp();
    // ---- measurement update: blocks of four scalar observations ----
    // Four conditionally independent scalar observations are absorbed at once: the 4 x 4 innovation
    // covariance S_b = P[I, I] + diag(r_I) is factorised redundantly by every lane in registers, Z = L_b^{-1} P[I, :]
    // by forward substitution, P -= Z^T Z, m += Z^T L_b^{-1}(y_I - m_I).  Same algebra as d rank-one updates, but a
    // quarter of the sequentially dependent stages (the kernel is bound by that latency, not by flops).
    // A missing entry becomes a virtual unit observation of nothing (zero row, unit pivot).
    T quad = T(0);
    if constexpr (TM % 4 == 0) {
      constexpr int NB = 4, RS = D + 4;
      // Blocks are walked lane-row group by lane-row group: the group loop stays rolled (its body, BPG blocks,
      // is ~1k instructions; unrolling all D / 4 blocks overflowed the instruction cache), the register rows a
      // block publishes are then compile-time indices hb * NB + aa.
      constexpr int BPG = TM / NB;
#pragma unroll 1  // (full unrolling re-measured after the round-2 cuts: 3.5 k SASS instructions, 4.03 vs 3.98 ms per call)
      for (int grp = 0; grp < GY; ++grp) {
#pragma unroll
      for (int hb = 0; hb < BPG; ++hb) {
        const int i0 = (grp * BPG + hb) * NB;
        const unsigned bits = (obs_mask >> i0) & 0xFu;
        if (bits != 0u) {  // warp-uniform
          T* rb = rowbuf + parity * (NB * RS);
          parity ^= 1;
          if (ty == grp) {
#pragma unroll
            for (int aa = 0; aa < NB; ++aa) {
              stv<T, TN>(rb + aa * RS + tx * TN, p[hb * NB + aa]);
              if (!MROW && tx == 0) rb[aa * RS + D] = mreg[hb * NB + aa];
            }
          }
          if (MROW && lane >= i0 && lane < i0 + NB) rb[(lane - i0) * RS + D] = m_own;
          __syncwarp();
          // EXCH: every lane used to run the 4 x 4 substitution twice, on its TM row entries and on its TN column
          // entries of the same 4 x D matrix Z (P is symmetric).  With TM = 8 that was 2/3 of the substitution work and
          // a quarter of the kernel's FMA-pipe time; now only the TN column entries are solved and the row entries are
          // re-read from a 4 x D shared image of Z that the ty = 0 lanes publish.
          constexpr bool EXCH = (TM % 2 == 0) && (TN % 2 == 0) && (TM > TN);
          T sr[NB][NB], rdl[NB], evl[NB], RA[NB][TM], RC[NB][TN], v[NB];
          ldv<T, NB>(rdl, rdv + i0);
          ldv<T, NB>(evl, ev + i0);
#pragma unroll
          for (int aa = 0; aa < NB; ++aa) {
            ldv<T, NB>(sr[aa], rb + aa * RS + i0);
            if (!EXCH) ldv<T, TM>(RA[aa], rb + aa * RS + ty * TM);
            ldv<T, TN>(RC[aa], rb + aa * RS + tx * TN);
            v[aa] = evl[aa] - rb[aa * RS + D];
          }
          T L[NB][NB], inv[NB];
          if (bits == 0xFu) {  // all four observed (the common case): S_b = P[I, I] + diag(r_I)
#pragma unroll
            for (int aa = 0; aa < NB; ++aa)
#pragma unroll
              for (int bb = 0; bb <= aa; ++bb) L[aa][bb] = (aa == bb) ? sr[aa][aa] + rdl[aa] : sr[aa][bb];
          } else {             // masked S_b: off-diagonals vanish, diagonal becomes 1, rows of P[I, :] are zeroed
#pragma unroll
            for (int aa = 0; aa < NB; ++aa) {
              const T oka = ((bits >> aa) & 1u) ? T(1) : T(0);
#pragma unroll
              for (int bb = 0; bb <= aa; ++bb) {
                const T okb = ((bits >> bb) & 1u) ? T(1) : T(0);
                L[aa][bb] = (aa == bb) ? (oka != T(0) ? sr[aa][aa] + rdl[aa] : T(1)) : oka * okb * sr[aa][bb];
              }
              if (!EXCH) {
#pragma unroll
                for (int r = 0; r < TM; ++r) RA[aa][r] *= oka;
              }
#pragma unroll
              for (int cc = 0; cc < TN; ++cc) RC[aa][cc] *= oka;
              v[aa] *= oka;
            }
          }
          // 4 x 4 Cholesky in registers
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            T piv = L[k][k];
#pragma unroll
            for (int j = 0; j < k; ++j) piv = fma_(-L[k][j], L[k][j], piv);
            if (lane == 0) svbuf[i0 + k] = piv;   // conditional innovation variance of observation i0 + k
            T lkk;
            pivot_fast_<T>(piv, lkk, inv[k]);
            L[k][k] = lkk;
#pragma unroll
            for (int r = k + 1; r < NB; ++r) {
              T acc = L[r][k];
#pragma unroll
              for (int j = 0; j < k; ++j) acc = fma_(-L[r][j], L[k][j], acc);
              L[r][k] = acc * inv[k];
            }
          }
          // forward substitution on the 4-vectors held by this lane
          const T n10 = -L[1][0], n20 = -L[2][0], n21 = -L[2][1], n30 = -L[3][0], n31 = -L[3][1], n32 = -L[3][2];
          auto fsub = [&](T& x0, T& x1, T& x2, T& x3) {
            x0 = x0 * inv[0];
            x1 = fma_(n10, x0, x1) * inv[1];
            x2 = fma_(n21, x1, fma_(n20, x0, x2)) * inv[2];
            x3 = fma_(n32, x2, fma_(n31, x1, fma_(n30, x0, x3))) * inv[3];
          };
          // the same substitution on two right-hand sides at once (packed pairs, bit-identical per half)
          auto fsub2 = [&](T& a0, T& b0, T& a1, T& b1, T& a2, T& b2, T& a3, T& b3) {
            Pair<T> x0 = pair_mul_s(inv[0], pair_make(a0, b0));
            Pair<T> x1 = pair_mul_s(inv[1], pair_fma_s(n10, x0, pair_make(a1, b1)));
            Pair<T> x2 = pair_mul_s(inv[2], pair_fma_s(n21, x1, pair_fma_s(n20, x0, pair_make(a2, b2))));
            Pair<T> x3 = pair_mul_s(inv[3], pair_fma_s(n32, x2, pair_fma_s(n31, x1, pair_fma_s(n30, x0, pair_make(a3, b3)))));
            pair_get(x0, a0, b0); pair_get(x1, a1, b1); pair_get(x2, a2, b2); pair_get(x3, a3, b3);
          };
          if constexpr (EXCH) {
#pragma unroll
            for (int cc = 0; cc < TN; cc += 2)
              fsub2(RC[0][cc], RC[0][cc + 1], RC[1][cc], RC[1][cc + 1], RC[2][cc], RC[2][cc + 1], RC[3][cc], RC[3][cc + 1]);
            T* zb = zbuf_[wq];
            if (ty == 0) {
#pragma unroll
              for (int aa = 0; aa < NB; ++aa) stv<T, TN>(zb + aa * RS + tx * TN, RC[aa]);
            }
            __syncwarp();
#pragma unroll
            for (int aa = 0; aa < NB; ++aa) ldv<T, TM>(RA[aa], zb + aa * RS + ty * TM);
          } else if constexpr (TM % 2 == 0 && TN % 2 == 0) {
#pragma unroll
            for (int r = 0; r < TM; r += 2)
              fsub2(RA[0][r], RA[0][r + 1], RA[1][r], RA[1][r + 1], RA[2][r], RA[2][r + 1], RA[3][r], RA[3][r + 1]);
#pragma unroll
            for (int cc = 0; cc < TN; cc += 2)
              fsub2(RC[0][cc], RC[0][cc + 1], RC[1][cc], RC[1][cc + 1], RC[2][cc], RC[2][cc + 1], RC[3][cc], RC[3][cc + 1]);
          } else {
#pragma unroll
            for (int r = 0; r < TM; ++r) fsub(RA[0][r], RA[1][r], RA[2][r], RA[3][r]);
#pragma unroll
            for (int cc = 0; cc < TN; ++cc) fsub(RC[0][cc], RC[1][cc], RC[2][cc], RC[3][cc]);
          }
          fsub(v[0], v[1], v[2], v[3]);
#pragma unroll
          for (int aa = 0; aa < NB; ++aa) {
            if constexpr (TN % 2 == 0) {
              Pair<T> nrc[TN / 2];  // -Z[aa, columns of this lane], negated once per block row
#pragma unroll
              for (int c2 = 0; c2 < TN / 2; ++c2) nrc[c2] = pair_make(-RC[aa][2 * c2], -RC[aa][2 * c2 + 1]);
#pragma unroll
              for (int r = 0; r < TM; ++r) {
#pragma unroll
                for (int c2 = 0; c2 < TN / 2; ++c2) {
                  const Pair<T> pp = pair_fma_s(RA[aa][r], nrc[c2], pair_make(p[r][2 * c2], p[r][2 * c2 + 1]));
                  pair_get(pp, p[r][2 * c2], p[r][2 * c2 + 1]);
                }
                if (!MROW) mreg[r] = fma_(RA[aa][r], v[aa], mreg[r]);  // only the tx == 0 copies are ever read
              }
              if (MROW) m_own = fma_(zbuf_[wq][aa * RS + (lane < D ? lane : 0)], v[aa], m_own);  // m_row += Z[aa][row] v[aa]
            } else {
#pragma unroll
              for (int r = 0; r < TM; ++r) {
#pragma unroll
                for (int cc = 0; cc < TN; ++cc) p[r][cc] = fma_(-RA[aa][r], RC[aa][cc], p[r][cc]);
                mreg[r] = fma_(RA[aa][r], v[aa], mreg[r]);  // only the tx == 0 copies are ever read
              }
            }
            quad = fma_(v[aa], v[aa], quad);
          }
        }
      }
      }
    } else {
#pragma unroll
      for (int i0 = 0; i0 < D; i0 += TM) {
#pragma unroll
        for (int qq = 0; qq < TM; ++qq) {
          const int i = i0 + qq;
          if (i < d && ((obs_mask >> i) & 1u)) {  // warp-uniform
            T* rb = rowbuf + parity * (D + 4);
            parity ^= 1;
            if (ty == i0 / TM) {
              stv<T, TN>(rb + tx * TN, p[qq]);
              if (tx == 0) rb[D] = mreg[qq];
            }
            __syncwarp();
            const T sv = rb[i] + rdv[i];
            const T inv = T(1) / sv;
            const T innov = ev[i] - rb[D];
            T ra[TM], rc[TN];
            ldv<T, TM>(ra, rb + ty * TM);
            ldv<T, TN>(rc, rb + tx * TN);
#pragma unroll
            for (int r = 0; r < TM; ++r) {
              const T kr = ra[r] * inv;
#pragma unroll
              for (int cc = 0; cc < TN; ++cc) p[r][cc] = fma_(-kr, rc[cc], p[r][cc]);
              mreg[r] = fma_(kr, innov, mreg[r]);  // only the tx == 0 copies are ever read
            }
            quad = fma_(innov * innov, inv, quad);
            if (lane == 0) svbuf[i] = sv;
          }
        }
      }
    }
    // ---- log-likelihood increment of this step (nan_to_num -> 0, filtering.py:125) ----
    __syncwarp();
    T part = ((lane < d) && ((obs_mask >> lane) & 1u)) ? T(-0.5) * log(svbuf[lane]) - half_log_2pi<T>() : T(0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    T inc = part - T(0.5) * quad;
    if (!(inc == inc)) inc = T(0);
    if (obs_mask != 0u) ell += inc;
    // ---- outputs: tile -> shared -> coalesced rows ----
    // ell-only call (ms == Ps == NULL: the reverse proposal of kalman/generic.py:67 needs the likelihood only):
    // nothing is staged or stored, except the chunk-0 state a later chunk starts from.
    const bool need_stage = (t + 1 == t1 && j == 0 && j + 1 < a.n_chunks);  // the state a later chunk starts from
    if (!want_out && !need_stage) continue;
    if constexpr (MROW && TN == 4) {
      // Direct stores from the register tile: a lane's four columns of a row are 16 contiguous bytes of the compact
      // row-major P_t (8-byte aligned when d is even), so the tile needs no trip through shared memory -- 8 STS.128 +
      // 8 LDS.128 (64 shared-memory wavefronts) and two warp barriers fewer per step, the same 16 STG.64 per lane.
      if (want_out && pairs_ok && !need_stage) {
        T* Pt = Psg + (int64_t)t * a.d.Ps.time;
        const int c0 = tx * TN;
#pragma unroll
        for (int r = 0; r < TM; ++r) {
          const int row = ty * TM + r;
          if (row < d) {
            const T va[2] = {p[r][0], p[r][1]}, vb[2] = {p[r][2], p[r][3]};
            if (c0 < d) stv<T, 2>(Pt + row * d + c0, va);
            if (c0 + 2 < d) stv<T, 2>(Pt + row * d + c0 + 2, vb);
          }
        }
        if (lane < d) msg[(int64_t)t * a.d.ms.time + lane] = m_own;
        continue;
      }
    }
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      stv<T, TN>(stage + (ty * TM + r) * LDS_ + tx * TN, p[r]);
      if (!MROW && tx == 0) mvec[ty * TM + r] = mreg[r];
    }
    if (MROW && lane < D) mvec[lane] = m_own;
    __syncwarp();
    if (!want_out) continue;
    T* Pt = Psg + (int64_t)t * a.d.Ps.time;
    if (pairs_ok) {
      // a lane owns four consecutive columns of a row (16-byte shared load, two 2-element global stores)
      constexpr int LPR = D / 4, RPP = 32 / LPR;
      const int lr = lane / LPR, c4 = (lane % LPR) * 4;
#pragma unroll 4
      for (int r0 = 0; r0 < d; r0 += RPP) {
        const int r = r0 + lr;
        if (r < d) {
          T v[4];
          ldv<T, 4>(v, stage + r * LDS_ + c4);
          const T va[2] = {v[0], v[1]}, vb[2] = {v[2], v[3]};
          if (c4 < d) stv<T, 2>(Pt + r * d + c4, va);
          if (c4 + 2 < d) stv<T, 2>(Pt + r * d + c4 + 2, vb);
        }
      }
    } else if (lane < d) {
#pragma unroll 6
      for (int r = 0; r < d; ++r) Pt[r * d + lane] = stage[r * LDS_ + lane];
    }
    if (lane < d) msg[(int64_t)t * a.d.ms.time + lane] = mvec[lane];
    __syncwarp();
  }
  // chunk outputs
  if (j == 0 && j + 1 < a.n_chunks) {
    T* pf = a.prefix + ((int64_t)unit * a.n_chunks + j) * psz;
    if (lane < d) pf[lane] = mvec[lane];
    for (int r = 0; r < d; ++r)
      if (lane < d) pf[d + r * d + lane] = stage[r * LDS_ + lane];
  }
  if (lane == 0) a.ellpart[(int64_t)unit * a.n_chunks + j] = ell;
}

}  // namespace auxssm
#include "kalman_filter_tile.cuh"
namespace auxssm {

// ---- parallel-in-time: elements and the associative combine ------------------------------------

template <typename T, int D> struct ScanBuf {
  T *A1, *C1, *J1;        // running aggregate (A plain)
  T *A2t, *C2, *J2;       // current element (A transposed)
  T *B1, *B2, *B3, *B4, *B5, *B6, *B7;
  T *b1, *e1, *b2, *e2, *w, *tmp;
  int* perm;
  int* ired;
};

template <int D> __host__ __device__ constexpr size_t scan_smem_elems() {
  return (size_t)13 * D * Tile<D>::LD + 6 * D + 11 * D + D + 2 * D /* perm, ired */ + 32;
}

// Element (A, b, C, eta, J) of time step t (t >= 1, not the first element of the scan), with A
// stored TRANSPOSED in sc.A2t.  _filtering_init_one (filtering.py:197-250) with m = 0, P = 0, i.e.
// m^- = b, P^- = Q.  Uses B1..B7 and the StepBuf vectors as scratch.
template <typename T, int D>
__device__ __forceinline__ void build_element(const FilterArgs<T>& a, int c, int b, int t, StepBuf<T, D>& s,
                                              ScanBuf<T, D>& sc) {
  constexpr int LD = Tile<D>::LD, NT = Tile<D>::NT;
  const int dx = a.d.dx, dy = a.d.dy, i = threadIdx.x;
  T* F = sc.B1; T* Ft = sc.B2; T* Ht = sc.B3; T* Hp = sc.B4; T* S = sc.B5; T* Ut = sc.B6; T* Z = sc.B7;
  // stage (always reloaded: the scratch buffers are recycled by the combine)
  stage_dyn<T, D>(a, c, b, t - 1, true, Ft, F, sc.C2, sc.b2);  // C2 <- Q (= P^-), b2 <- b (= m^-)
  s.Ht = Ht; s.R = S;
  for (int e = i; e < D * LD; e += NT) { Ht[e] = T(0); Hp[e] = T(0); S[e] = T(0); }
  const int n_obs = stage_obs<T, D>(a, c, b, t, true, s, Hp, true);
  if (n_obs == 0) {  // _passthrough, filtering.py:239-248
    copy_mat<T, D>(Ft, sc.A2t, dx, dx);
    for (int e = i; e < D * LD; e += NT) sc.J2[e] = T(0);
    if (i < dx) sc.e2[i] = T(0);
    symmetrize<T, D>(sc.C2, dx);
    return;
  }
  mv_tn<T, D, 0>(Ht, sc.b2, s.v, nullptr, dx, dy);                 // v = H b
  if (i < dy) s.v[i] = s.mask[i] ? T(0) : s.y[i] - s.v[i] - s.c[i];  // y - H b - c
  mm_tn<T, D, 0>(sc.C2, Ht, Ut, nullptr, dx);                      // Ut = Q H^T          (dx x dy)
  mm_tn<T, D, 1>(Ht, Ut, S, S, dx);                                // S = H Q H^T + R
  transpose_to<T, D>(Ut, Z, dx, dy);                               // Z <- U = H Q        (dy x dx)
  chol_inplace<T, D>(S, s.dg, dy);
  fwd_solve<T, D>(S, Z, s.v, dy, dx, s.dg);                              // Z = L^{-1} H Q ; v = L^{-1}(y - Hb - c)
  fwd_solve<T, D>(S, Hp, nullptr, dy, dx, s.dg);                         // Hp <- Zh = L^{-1} H
  mm_tn<T, D, 2>(Z, Z, sc.C2, sc.C2, dy);                          // C = Q - Z^T Z
  mv_tn<T, D, 1>(Z, s.v, sc.b2, sc.b2, dy, dx);                    // b = b + Z^T v
  symmetrize<T, D>(sc.C2, dx);
  mm_tn<T, D, 0>(Hp, Z, Ut, nullptr, dy);                          // Ut <- (K H)^T = Zh^T Z
  mm_tn<T, D, 2>(F, Ut, sc.A2t, Ft, dx);                           // A^T = F^T - F^T (K H)^T
  mm_tn<T, D, 0>(Hp, Hp, Ut, nullptr, dy);                         // N = Zh^T Zh = H^T S^{-1} H
  mm_tn<T, D, 0>(Ut, F, Z, nullptr, dx);                           // Z <- N F
  mm_tn<T, D, 0>(F, Z, sc.J2, nullptr, dx);                        // J = F^T N F
  symmetrize<T, D>(sc.J2, dx);
  mv_tn<T, D, 0>(Hp, s.v, sc.tmp, nullptr, dy, dx);                // tmp = Zh^T v = H^T S^{-1}(y - Hb - c)
  mv_tn<T, D, 0>(F, sc.tmp, sc.e2, nullptr, dx, dx);               // eta = F^T tmp
}

// (A1, b1, C1, eta1, J1) <- (A1, ...) (x) (A2, b2, C2, eta2, J2).   _filtering_op_impl,
// filtering.py:164-183, with one LU of M = I + C1 J2 serving every solve:
//   V = M^{-1} A1, Dm = M^{-1} C1, w = M^{-1}(b1 + C1 eta2);
//   A = A2 V; b = A2 w + b2; C = A2 Dm A2^T + C2; eta = V^T (eta2 - J2 b1) + eta1; J = V^T J2 A1 + J1.
// `full` = false skips A, eta, J (enough when only the (b, C) prefix is needed).
template <typename T, int D>
__device__ __forceinline__ void combine(ScanBuf<T, D>& sc, T* red, int dx, bool full) {
  constexpr int LD = Tile<D>::LD;
  const int i = threadIdx.x;
  T* M = sc.B1; T* V = sc.B2; T* Dm = sc.B3; T* An = sc.B4; T* W = sc.B5;
  mm_tn<T, D, 0>(sc.C1, sc.J2, M, nullptr, dx);  // C1 J2
  if (i < dx) M[i * LD + i] += T(1);
  mv_tn<T, D, 1>(sc.C1, sc.e2, sc.w, sc.b1, dx, dx);  // w = b1 + C1 eta2
  lu_inplace<T, D>(M, sc.perm, red, sc.ired, dx);
  lu_solve<T, D>(M, sc.perm, sc.C1, Dm, sc.w, sc.tmp, dx, dx);  // Dm = M^{-1} C1 ; tmp = M^{-1} w
  if (full) {
    lu_solve<T, D>(M, sc.perm, sc.A1, V, nullptr, nullptr, dx, dx);  // V = M^{-1} A1
    mm_tn<T, D, 0>(sc.J2, sc.A1, W, nullptr, dx);                    // W = J2 A1
    mm_tn<T, D, 1>(V, W, sc.J1, sc.J1, dx);                          // J = V^T W + J1
    symmetrize<T, D>(sc.J1, dx);
    mv_tn<T, D, 2>(sc.J2, sc.b1, sc.w, sc.e2, dx, dx);               // w = eta2 - J2 b1
    mv_tn<T, D, 1>(V, sc.w, sc.e1, sc.e1, dx, dx);                   // eta = V^T w + eta1
    mm_tn<T, D, 0>(sc.A2t, V, An, nullptr, dx);                      // A = A2 V
  }
  mv_tn<T, D, 1>(sc.A2t, sc.tmp, sc.b1, sc.b2, dx, dx);              // b = A2 M^{-1}(b1 + C1 eta2) + b2
  mm_tn<T, D, 0>(Dm, sc.A2t, W, nullptr, dx);                        // W = Dm^T A2^T
  mm_tn<T, D, 1>(sc.A2t, W, sc.C1, sc.C2, dx);                       // C = A2 W + C2
  symmetrize<T, D>(sc.C1, dx);
  if (full) { T* t = sc.A1; sc.A1 = An; sc.B4 = t; }
}

template <typename T, int D> __device__ __forceinline__ ScanBuf<T, D> carve_scan(Carve<T, D>& cv) {
  ScanBuf<T, D> sc;
  sc.A1 = cv.mat(); sc.C1 = cv.mat(); sc.J1 = cv.mat();
  sc.A2t = cv.mat(); sc.C2 = cv.mat(); sc.J2 = cv.mat();
  sc.B1 = cv.mat(); sc.B2 = cv.mat(); sc.B3 = cv.mat(); sc.B4 = cv.mat(); sc.B5 = cv.mat();
  sc.B6 = cv.mat(); sc.B7 = cv.mat();
  sc.b1 = cv.vec(); sc.e1 = cv.vec(); sc.b2 = cv.vec(); sc.e2 = cv.vec(); sc.w = cv.vec(); sc.tmp = cv.vec();
  return sc;
}

template <typename T, int D>
__device__ __forceinline__ void carve_scan_tail(Carve<T, D>& cv, StepBuf<T, D>& s, ScanBuf<T, D>& sc) {
  s.P = s.W = s.S = s.U = s.Ft = s.Q = s.Ht = s.R = nullptr;
  s.m = cv.vec(); s.bv = cv.vec(); s.y = cv.vec(); s.c = cv.vec(); s.v = cv.vec(); s.red = cv.vec();
  s.dg = cv.p; cv.p += CholWs<D>::WORDS;
  s.rd = cv.vec();
  s.mask = reinterpret_cast<int*>(cv.p);
  cv.p += (D * sizeof(int) + sizeof(T) - 1) / sizeof(T);
  sc.perm = reinterpret_cast<int*>(cv.p);
  cv.p += (D * sizeof(int) + sizeof(T) - 1) / sizeof(T);
  sc.ired = reinterpret_cast<int*>(cv.p);
  cv.p += (D * sizeof(int) + sizeof(T) - 1) / sizeof(T);
  s.h_clean = s.r_clean = s.c_clean = false;
}

// pass A for chunks 1..n-2: aggregate = elem(t0) (x) elem(t0+1) (x) ... (x) elem(t1-1)
template <typename T, int D>
__global__ void __launch_bounds__(Tile<D>::NT, Occ<T, D>::MINB) filter_aggregate_kernel(FilterArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Carve<T, D> cv{reinterpret_cast<T*>(smem_raw)};
  constexpr int NT = Tile<D>::NT;
  const size_t total = scan_smem_elems<D>();
  for (size_t e = threadIdx.x; e < total; e += NT) cv.p[e] = T(0);
  __syncthreads();
  ScanBuf<T, D> sc = carve_scan<T, D>(cv);
  StepBuf<T, D> s;
  carve_scan_tail<T, D>(cv, s, sc);
  const int j = 1 + blockIdx.y;
  const int unit = blockIdx.x;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx;
  int t0, t1;
  chunk_range(j, a.chunk_len, a.d.T, t0, t1);
  for (int t = t0; t < t1; ++t) {
    build_element<T, D>(a, c, b, t, s, sc);
    if (t == t0) {
      transpose_to<T, D>(sc.A2t, sc.A1, dx, dx);
      copy_mat<T, D>(sc.C2, sc.C1, dx, dx);
      copy_mat<T, D>(sc.J2, sc.J1, dx, dx);
      if (threadIdx.x < dx) { sc.b1[threadIdx.x] = sc.b2[threadIdx.x]; sc.e1[threadIdx.x] = sc.e2[threadIdx.x]; }
      cta_sync<D>();
    } else {
      combine<T, D>(sc, s.red, dx, true);
    }
  }
  cta_sync<D>();
  const int esz = 3 * dx * dx + 2 * dx;
  T* g = a.agg + ((int64_t)unit * a.n_chunks + j) * esz;
  store_mat<T, D>(g, sc.A1, dx, dx);
  store_mat<T, D>(g + dx * dx, sc.C1, dx, dx);
  store_mat<T, D>(g + 2 * dx * dx, sc.J1, dx, dx);
  if (threadIdx.x < dx) {
    g[3 * dx * dx + threadIdx.x] = sc.b1[threadIdx.x];
    g[3 * dx * dx + dx + threadIdx.x] = sc.e1[threadIdx.x];
  }
}

// pass B: prefix_j = prefix_{j-1} (x) aggregate_j for j = 1..n-2, (b, C) components only.
template <typename T, int D>
__global__ void __launch_bounds__(Tile<D>::NT, Occ<T, D>::MINB) filter_prefix_kernel(FilterArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Carve<T, D> cv{reinterpret_cast<T*>(smem_raw)};
  constexpr int NT = Tile<D>::NT;
  const size_t total = scan_smem_elems<D>();
  for (size_t e = threadIdx.x; e < total; e += NT) cv.p[e] = T(0);
  __syncthreads();
  ScanBuf<T, D> sc = carve_scan<T, D>(cv);
  StepBuf<T, D> s;
  carve_scan_tail<T, D>(cv, s, sc);
  const int unit = blockIdx.x;
  const int dx = a.d.dx;
  const int psz = dx + dx * dx, esz = 3 * dx * dx + 2 * dx;
  T* pf = a.prefix + (int64_t)unit * a.n_chunks * psz;
  if (threadIdx.x < dx) sc.b1[threadIdx.x] = pf[threadIdx.x];
  load_mat<T, D, false>(pf + dx, sc.C1, dx, dx);
  for (int j = 1; j + 1 < a.n_chunks; ++j) {
    const T* g = a.agg + ((int64_t)unit * a.n_chunks + j) * esz;
    cta_sync<D>();
    load_mat<T, D, true>(g, sc.A2t, dx, dx);
    load_mat<T, D, false>(g + dx * dx, sc.C2, dx, dx);
    load_mat<T, D, false>(g + 2 * dx * dx, sc.J2, dx, dx);
    if (threadIdx.x < dx) {
      sc.b2[threadIdx.x] = g[3 * dx * dx + threadIdx.x];
      sc.e2[threadIdx.x] = g[3 * dx * dx + dx + threadIdx.x];
    }
    cta_sync<D>();
    combine<T, D>(sc, s.red, dx, false);
    T* po = pf + (int64_t)j * psz;
    if (threadIdx.x < dx) po[threadIdx.x] = sc.b1[threadIdx.x];
    store_mat<T, D>(po + dx, sc.C1, dx, dx);
  }
}

// ell[c] = sum over inner batch and chunks, in a fixed order (deterministic).
template <typename T> __global__ void ell_reduce_kernel(const T* ellpart, T* ell, int n_chains, int per_chain) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  T s = T(0);
  for (int k = 0; k < per_chain; ++k) s += ellpart[(int64_t)c * per_chain + k];
  ell[c] = s;
}

// ---- host side -----------------------------------------------------------------------------------

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct FilterPlan {
  int n_chunks, chunk_len;
  size_t off_ell, off_prefix, off_agg, total;
};

static FilterPlan make_plan(const auxssm_lgssm_desc* d, size_t es) {
  FilterPlan p;
  p.n_chunks = resolve_time_chunks(d);
  const int n_elem = d->T > 1 ? d->T - 1 : 1;
  p.chunk_len = (n_elem + p.n_chunks - 1) / p.n_chunks;
  p.n_chunks = (n_elem + p.chunk_len - 1) / p.chunk_len;  // drop empty trailing chunks
  const size_t units = (size_t)d->n_chains * d->n_batch;
  size_t off = kGateBytes;  // [0, kGateBytes): the diagonal-F gate word
  p.off_ell = off; off = align_up(off + units * p.n_chunks * es, 256);
  p.off_prefix = off;
  if (p.n_chunks > 1) off = align_up(off + units * p.n_chunks * ((size_t)d->dx + (size_t)d->dx * d->dx) * es, 256);
  p.off_agg = off;
  if (p.n_chunks > 2) off = align_up(off + units * p.n_chunks * (3 * (size_t)d->dx * d->dx + 2 * (size_t)d->dx) * es, 256);
  p.total = off;
  return p;
}

template <typename T, int D>
static int launch_filter(cudaStream_t st, const FilterArgs<T>& a) {
  const bool hid = (a.d.flags & AUXSSM_FLAG_H_IDENTITY) != 0;
  const int units = a.d.n_chains * a.d.n_batch;
  constexpr int NT = Tile<D>::NT;
  const size_t smem_step = step_smem_elems<D>(hid) * sizeof(T);
  auto rec = hid ? filter_recursion_kernel<T, D, true> : filter_recursion_kernel<T, D, false>;
  int rc = ensure_smem(rec, smem_step, "auxssm_kalman_filter");
  if (rc) return rc;
  // structured fast path: H = I, diagonal R, time-invariant diagonal F / Q / b -> one warp per chain
  const uint32_t need = AUXSSM_FLAG_H_IDENTITY | AUXSSM_FLAG_R_DIAGONAL | AUXSSM_FLAG_F_IS_DIAGONAL;
  const bool warp_path = D >= 8 && D <= 32 && (a.d.flags & need) == need &&
                         (a.d.T == 1 || (a.d.Fs.time == 0 && a.d.Qs.time == 0 && a.d.bs.time == 0));
  // same structure at D = 64: four warps per chain (kalman_filter_tile.cuh)
  const bool tile_path = D == 64 && (a.d.flags & need) == need &&
                         (a.d.T == 1 || (a.d.Fs.time == 0 && a.d.Qs.time == 0 && a.d.bs.time == 0));
  constexpr int TD = (D == 64) ? 64 : 64, TNW = 4;
  const size_t smem_tile = filter_tile_smem_elems<TD, TNW>() * sizeof(T);
  FilterArgs<T> g = a;  // arguments of the general kernel behind a structured one (gated the other way round)
  const bool structured = warp_path || tile_path;
  FilterArgs<T> as = a;
  if (structured && a.d.T > 1) {
    rc = launch_count_offdiag<T>(st, a.Fs, a.d, a.gate);
    if (rc) return rc;
  } else {
    as.gate = nullptr;
    g.gate = nullptr;
  }
  auto launch_rec = [&](int chunk_first, int n_y) -> int {
    if (warp_path) {
      filter_warp_kernel<T, (D >= 8 && D <= 32) ? D : 8><<<dim3((units + 3) / 4, n_y), 128, 0, st>>>(as, chunk_first);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_filter (warp recursion)");
    } else if (tile_path) {
      filter_tile_kernel<T, TD, TNW><<<dim3(units, n_y), 128, smem_tile, st>>>(as, chunk_first);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_filter (tile recursion)");
    }
    if (!structured || g.gate != nullptr) {
      rec<<<dim3(units, n_y), NT, smem_step, st>>>(g, chunk_first);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_filter (recursion)");
    }
    return AUXSSM_OK;
  };
  rc = launch_rec(0, 1);
  if (rc) return rc;
  if (a.n_chunks > 1) {
    const size_t smem_scan = scan_smem_elems<D>() * sizeof(T);
    if (a.n_chunks > 2) {
      rc = ensure_smem(filter_aggregate_kernel<T, D>, smem_scan, "auxssm_kalman_filter (parallel)");
      if (rc) return rc;
      rc = ensure_smem(filter_prefix_kernel<T, D>, smem_scan, "auxssm_kalman_filter (parallel)");
      if (rc) return rc;
      filter_aggregate_kernel<T, D><<<dim3(units, a.n_chunks - 2), NT, smem_scan, st>>>(a);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_filter (aggregate)");
      filter_prefix_kernel<T, D><<<units, NT, smem_scan, st>>>(a);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_filter (prefix)");
    }
    rc = launch_rec(1, a.n_chunks - 1);
    if (rc) return rc;
  }
  return AUXSSM_OK;
}

template <typename T>
static int filter_entry(void* stream, const auxssm_lgssm_desc* d, const T* ys, const T* m0, const T* P0, const T* Fs,
                        const T* Qs, const T* bs, const T* Hs, const T* Rs, const T* cs, T* ms, T* Ps, T* ell,
                        void* scratch, size_t scratch_bytes) {
  const char* who = "auxssm_kalman_filter";
  int rc = check_desc(d, who);
  if (rc) return rc;
  const bool hid = (d->flags & AUXSSM_FLAG_H_IDENTITY) != 0;
  if (!ys || !m0 || !P0 || !Rs || !ell || (!hid && !Hs) || (d->T > 1 && (!Fs || !Qs || !bs))) {
    set_error("%s: null data pointer", who);
    return AUXSSM_ERR_INVALID;
  }
  if ((ms == nullptr) != (Ps == nullptr)) {
    set_error("%s: ms and Ps must both be given or both be NULL (ell-only call)", who);
    return AUXSSM_ERR_INVALID;
  }
  if (hid && Hs != nullptr) { set_error("%s: Hs must be NULL with AUXSSM_FLAG_H_IDENTITY", who); return AUXSSM_ERR_INVALID; }
  const FilterPlan p = make_plan(d, sizeof(T));
  if (scratch_bytes < p.total || (p.total > 0 && scratch == nullptr)) {
    set_error("%s: scratch too small (%zu < %zu bytes)", who, scratch_bytes, p.total);
    return AUXSSM_ERR_SCRATCH;
  }
  FilterArgs<T> a;
  a.d = *d;
  a.ys = ys; a.m0 = m0; a.P0 = P0; a.Fs = Fs; a.Qs = Qs; a.bs = bs; a.Hs = Hs; a.Rs = Rs; a.cs = cs;
  a.ms = ms; a.Ps = Ps;
  unsigned char* sb = static_cast<unsigned char*>(scratch);
  a.ellpart = reinterpret_cast<T*>(sb + p.off_ell);
  a.prefix = reinterpret_cast<T*>(sb + p.off_prefix);
  a.agg = reinterpret_cast<T*>(sb + p.off_agg);
  a.gate = reinterpret_cast<int*>(sb);
  a.n_chunks = p.n_chunks;
  a.chunk_len = p.chunk_len;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dmax = d->dx > d->dy ? d->dx : d->dy;
  switch (padded_dim(dmax)) {
    case 4: rc = launch_filter<T, 4>(st, a); break;
    case 8: rc = launch_filter<T, 8>(st, a); break;
    case 16: rc = launch_filter<T, 16>(st, a); break;
    case 32: rc = launch_filter<T, 32>(st, a); break;
    case 64: rc = launch_filter<T, 64>(st, a); break;
    default: set_error("%s: unsupported dimension %d", who, dmax); return AUXSSM_ERR_UNSUPPORTED;
  }
  if (rc) return rc;
  ell_reduce_kernel<T><<<(d->n_chains + 127) / 128, 128, 0, st>>>(a.ellpart, ell, d->n_chains,
                                                                  d->n_batch * p.n_chunks);
  AUXSSM_CHECK_LAUNCH("auxssm_kalman_filter (ell reduce)");
  return AUXSSM_OK;
}

}  // namespace auxssm

// This translation unit carries the fp32 entry point (and the dtype-independent scratch query); the fp64
// instantiations are compiled by kalman_filter_f64.cu, which includes this file with AUXSSM_ONLY_F64 defined, so
// that the two halves build in parallel.
#ifndef AUXSSM_ONLY_F64
extern "C" __attribute__((visibility("default"))) size_t auxssm_kalman_filter_scratch_bytes(const auxssm_lgssm_desc* desc, int32_t elem_size) {
  if (auxssm::check_desc(desc, "auxssm_kalman_filter_scratch_bytes")) return 0;
  return auxssm::make_plan(desc, (size_t)elem_size).total;
}

extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_filter_f32(void* stream, const auxssm_lgssm_desc* desc, const float* ys,
                                            const float* m0, const float* P0, const float* Fs, const float* Qs,
                                            const float* bs, const float* Hs, const float* Rs, const float* cs,
                                            float* ms, float* Ps, float* ell, void* scratch, size_t scratch_bytes) {
  return auxssm::filter_entry<float>(stream, desc, ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs, ms, Ps, ell, scratch,
                                     scratch_bytes);
}
#else
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_filter_f64(void* stream, const auxssm_lgssm_desc* desc, const double* ys,
                                            const double* m0, const double* P0, const double* Fs, const double* Qs,
                                            const double* bs, const double* Hs, const double* Rs, const double* cs,
                                            double* ms, double* Ps, double* ell, void* scratch,
                                            size_t scratch_bytes) {
  return auxssm::filter_entry<double>(stream, desc, ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs, ms, Ps, ell, scratch,
                                      scratch_bytes);
}
#endif

// One auxiliary-Kalman MH sweep as ONE kernel for narrow states (d <= 4): the few-chain / long-T regime
// (BASELINE configs[0]) where a sweep issued as ~40 separate launches is launch-bound.
//
// Replaces, for a latent-Gaussian target  pi(x) ~ LGSSM prior(m0, P0, F_t, Q_t, b_t) x prod_t exp(g(x_t; y_t)):
//   aux_samplers/kalman/generic.py:53-106       kernel / do_one / _get_alpha
//   _primitives/kalman/filtering.py:49-63        _parallel_filtering  (lax.associative_scan over (A, b, C, eta, J))
//     :197-250 _filtering_init_one, :164-183 _filtering_op_impl, :84-147 sequential_update / predict_update (ell)
//   _primitives/kalman/sampling.py:11-40,52-136  reverse associative scan over (G, e), mean_and_chol, last step
//   _primitives/kalman/base.py:73-166            posterior_logpdf / prior_logpdf / log_likelihood
//   examples/stochastic_volatility/auxiliary_kalman.py:28-54   first / second order observation factories
//
// PARALLEL-IN-TIME DESIGN (the one BASELINE.json's north_star describes).  One thread owns one time step and
// builds its scan element in registers.  Scans are single-pass:
//   * inside a warp   Kogge-Stone over lanes with warp shuffles, the reference's associative operator applied
//                     to the register-resident elements (no shared memory);
//   * inside a CTA    the 32 warp aggregates are scanned by warp 0 the same way and applied as a prefix;
//   * across CTAs     (a chain whose T exceeds one tile is split over several CTAs) DECOUPLED LOOK-BACK: a tile
//                     publishes its aggregate with status AGG, walks back over its predecessors combining
//                     aggregates until it meets a status-PREFIX tile, then publishes its own inclusive prefix.
//                     Payloads are multi-word, so each lives in global memory next to a status word, written
//                     payload -> __threadfence -> status.
// The forward filter scan, the REVERSE scan of the sampler's affine maps (G_t, e_t) (look-AHEAD over later tiles),
// and the second forward scan of the reverse proposal all use the same routine.  The chain-wide sums of the MH ratio
// are combined by the last tile to arrive (ticket counter), which publishes the accept decision the other tiles
// wait for.  All CTAs of a launch must be co-resident (the host checks n_chains x n_tiles against the device).
#include "common.cuh"

namespace auxssm {

constexpr int kStatusAgg = 1, kStatusPrefix = 2;
constexpr double kHalfLog2Pi = 0.91893853320467274178;

// ---- tiny dense algebra in registers (D <= 4, runtime d <= D; entries beyond d are kept at identity / zero) ----------
template <typename T, int D> struct Mat { T a[D][D]; };
template <typename T, int D> struct Vec { T a[D]; };

template <typename T, int D> __device__ __forceinline__ Mat<T, D> m_zero() {
  Mat<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) r.a[i][j] = T(0);
  return r;
}
template <typename T, int D> __device__ __forceinline__ Mat<T, D> m_eye() {
  Mat<T, D> r = m_zero<T, D>();
#pragma unroll
  for (int i = 0; i < D; ++i) r.a[i][i] = T(1);
  return r;
}
template <typename T, int D> __device__ __forceinline__ Vec<T, D> v_zero() {
  Vec<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i) r.a[i] = T(0);
  return r;
}
template <typename T, int D> __device__ __forceinline__ Mat<T, D> mm(const Mat<T, D>& x, const Mat<T, D>& y) {
  Mat<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) s += x.a[i][k] * y.a[k][j];
      r.a[i][j] = s;
    }
  return r;
}
template <typename T, int D> __device__ __forceinline__ Mat<T, D> mmT(const Mat<T, D>& x, const Mat<T, D>& y) {  // x y^T
  Mat<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) s += x.a[i][k] * y.a[j][k];
      r.a[i][j] = s;
    }
  return r;
}
template <typename T, int D> __device__ __forceinline__ Mat<T, D> mTm(const Mat<T, D>& x, const Mat<T, D>& y) {  // x^T y
  Mat<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) s += x.a[k][i] * y.a[k][j];
      r.a[i][j] = s;
    }
  return r;
}
template <typename T, int D> __device__ __forceinline__ Vec<T, D> mv(const Mat<T, D>& x, const Vec<T, D>& v) {
  Vec<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T s = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) s += x.a[i][k] * v.a[k];
    r.a[i] = s;
  }
  return r;
}
template <typename T, int D> __device__ __forceinline__ Vec<T, D> mTv(const Mat<T, D>& x, const Vec<T, D>& v) {
  Vec<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T s = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) s += x.a[k][i] * v.a[k];
    r.a[i] = s;
  }
  return r;
}
template <typename T, int D> __device__ __forceinline__ Mat<T, D> m_add(const Mat<T, D>& x, const Mat<T, D>& y) {
  Mat<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) r.a[i][j] = x.a[i][j] + y.a[i][j];
  return r;
}
template <typename T, int D> __device__ __forceinline__ Mat<T, D> m_sub(const Mat<T, D>& x, const Mat<T, D>& y) {
  Mat<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) r.a[i][j] = x.a[i][j] - y.a[i][j];
  return r;
}
template <typename T, int D> __device__ __forceinline__ Mat<T, D> m_sym(const Mat<T, D>& x) {
  Mat<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) r.a[i][j] = T(0.5) * (x.a[i][j] + x.a[j][i]);
  return r;
}
template <typename T, int D> __device__ __forceinline__ Vec<T, D> v_add(const Vec<T, D>& x, const Vec<T, D>& y) {
  Vec<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i) r.a[i] = x.a[i] + y.a[i];
  return r;
}
template <typename T, int D> __device__ __forceinline__ Vec<T, D> v_sub(const Vec<T, D>& x, const Vec<T, D>& y) {
  Vec<T, D> r;
#pragma unroll
  for (int i = 0; i < D; ++i) r.a[i] = x.a[i] - y.a[i];
  return r;
}

// Lower Cholesky factor of the leading d x d block (identity on the padding); `ok` = every pivot positive.
template <typename T, int D> __device__ __forceinline__ Mat<T, D> chol(const Mat<T, D>& s, int d, bool& ok) {
  Mat<T, D> l = m_eye<T, D>();
  ok = true;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (j < d) {
      T v = s.a[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= l.a[j][k] * l.a[j][k];
      if (!(v > T(0))) ok = false;
      const T ljj = sqrt(v);
      l.a[j][j] = ljj;
#pragma unroll
      for (int i = j + 1; i < D; ++i) {
        if (i < d) {
          T w = s.a[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) w -= l.a[i][k] * l.a[j][k];
          l.a[i][j] = w / ljj;
        }
      }
    }
  }
  return l;
}
// z = L^{-1} v (L lower, unit padding)
template <typename T, int D> __device__ __forceinline__ Vec<T, D> tri_solve(const Mat<T, D>& l, const Vec<T, D>& v) {
  Vec<T, D> z;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T s = v.a[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= l.a[i][k] * z.a[k];
    z.a[i] = s / l.a[i][i];
  }
  return z;
}
// inverse of a general matrix by Gauss-Jordan with partial pivoting (padding block is the identity)
template <typename T, int D> __device__ __forceinline__ Mat<T, D> m_inv(Mat<T, D> m) {
  Mat<T, D> r = m_eye<T, D>();
#pragma unroll
  for (int c = 0; c < D; ++c) {
    int p = c;
    T best = fabs(m.a[c][c]);
#pragma unroll
    for (int i = c + 1; i < D; ++i) {
      const T v = fabs(m.a[i][c]);
      if (v > best) { best = v; p = i; }
    }
#pragma unroll
    for (int i = c + 1; i < D; ++i) {
      if (i == p) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
          T t = m.a[c][j]; m.a[c][j] = m.a[i][j]; m.a[i][j] = t;
          t = r.a[c][j]; r.a[c][j] = r.a[i][j]; r.a[i][j] = t;
        }
      }
    }
    const T inv = T(1) / m.a[c][c];
#pragma unroll
    for (int j = 0; j < D; ++j) { m.a[c][j] *= inv; r.a[c][j] *= inv; }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i != c) {
        const T f = m.a[i][c];
#pragma unroll
        for (int j = 0; j < D; ++j) { m.a[i][j] -= f * m.a[c][j]; r.a[i][j] -= f * r.a[c][j]; }
      }
    }
  }
  return r;
}
// S^{-1} for SPD S through its Cholesky factor
template <typename T, int D> __device__ __forceinline__ Mat<T, D> spd_inv(const Mat<T, D>& s, int d) {
  bool ok;
  const Mat<T, D> l = chol<T, D>(s, d, ok);
  Mat<T, D> li = m_zero<T, D>();  // L^{-1}
#pragma unroll
  for (int c = 0; c < D; ++c) {
    Vec<T, D> e = v_zero<T, D>();
    e.a[c] = T(1);
    const Vec<T, D> z = tri_solve<T, D>(l, e);
#pragma unroll
    for (int i = 0; i < D; ++i) li.a[i][c] = z.a[i];
  }
  return mTm<T, D>(li, li);
}
// log N(v; 0, L L^T) over the leading d entries; NaN -> 0 like nan_to_num / nansum in the reference
template <typename T, int D> __device__ __forceinline__ T gauss_logpdf(const Mat<T, D>& l, const Vec<T, D>& v, int d) {
  const Vec<T, D> z = tri_solve<T, D>(l, v);
  T q = T(0), ld = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (i < d) { q += z.a[i] * z.a[i]; ld += log(l.a[i][i]); }
  }
  const T out = T(-0.5) * q - ld - T(d) * T(kHalfLog2Pi);
  return (out == out) ? out : T(0);
}

// ---- scan elements ------------------------------------------------------------------------------------------------
template <typename T, int D> struct FiltElem { Mat<T, D> A, C, J; Vec<T, D> b, eta; };
template <typename T, int D> struct SampElem { Mat<T, D> G; Vec<T, D> e; };

// _filtering_op_impl (filtering.py:164-183): elem1 (earlier) (x) elem2 (later)
template <typename T, int D>
__device__ __forceinline__ FiltElem<T, D> combine(const FiltElem<T, D>& e1, const FiltElem<T, D>& e2) {
  const Mat<T, D> M = m_add<T, D>(m_eye<T, D>(), mm<T, D>(e1.C, e2.J));  // I + C1 J2  (I + J2 C1 is its transpose)
  const Mat<T, D> Mi = m_inv<T, D>(M);
  const Mat<T, D> X = mm<T, D>(e2.A, Mi);            // A2 (I + C1 J2)^{-1}
  Mat<T, D> Y;                                       // A1^T (I + J2 C1)^{-1} = A1^T Mi^T
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) s += e1.A.a[k][i] * Mi.a[j][k];
      Y.a[i][j] = s;
    }
  FiltElem<T, D> r;
  r.A = mm<T, D>(X, e1.A);
  r.b = v_add<T, D>(mv<T, D>(X, v_add<T, D>(e1.b, mv<T, D>(e1.C, e2.eta))), e2.b);
  r.C = m_sym<T, D>(m_add<T, D>(mm<T, D>(X, mmT<T, D>(e1.C, e2.A)), e2.C));
  r.eta = v_add<T, D>(mv<T, D>(Y, v_sub<T, D>(e2.eta, mv<T, D>(e2.J, e1.b))), e1.eta);
  r.J = m_sym<T, D>(m_add<T, D>(mm<T, D>(Y, mm<T, D>(e2.J, e1.A)), e1.J));
  return r;
}
// _sampling_op_impl (sampling.py:52-55) in the reverse scan: elem1 = the later (already accumulated) map
template <typename T, int D>
__device__ __forceinline__ SampElem<T, D> combine(const SampElem<T, D>& e1, const SampElem<T, D>& e2) {
  SampElem<T, D> r;
  r.G = mm<T, D>(e2.G, e1.G);
  r.e = v_add<T, D>(mv<T, D>(e2.G, e1.e), e2.e);
  return r;
}

template <typename E> struct Words { static constexpr int N = sizeof(E) / 4; };
template <typename E> __device__ __forceinline__ E shfl_elem(const E& e, int src_lane) {
  E r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&e);
  uint32_t* dptr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int w = 0; w < Words<E>::N; ++w) dptr[w] = __shfl_sync(0xffffffffu, s[w], src_lane);
  return r;
}

// Look-back state of one scan of one chain: status[n_tiles] (zeroed before the launch), agg / prefix payloads.
template <typename E> struct LookBack {
  int* status;
  E* agg;
  E* prefix;
};

// Inclusive scan of `mine` over the tile (position `pos` in [0, n_valid) in scan order; REVERSE scans from the
// last thread / last tile down).  Returns the inclusive value in `mine`; `excl_valid` / `excl` receive the
// inclusive value of the position before this thread's (from the previous lane, warp or tile).
template <typename E, bool REVERSE>
__device__ __forceinline__ void tile_scan(E& mine, bool valid, int tile, int n_tiles, LookBack<E> lb, E* warp_buf,
                                          E* bcast, bool& excl_valid, E& excl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // scan order inside the CTA: thread index ascending (forward) or descending (reverse)
  const int slane = REVERSE ? 31 - lane : lane, swarp = REVERSE ? nw - 1 - warp : warp;
  // 1. warp level (Kogge-Stone with shuffles): position p takes from p - off
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int src = REVERSE ? lane + off : lane - off;
    const E other = shfl_elem<E>(mine, src & 31);
    const bool ovalid = __shfl_sync(0xffffffffu, valid ? 1 : 0, src & 31) != 0;
    if (slane >= off && valid && ovalid) mine = combine(other, mine);
  }
  // 2. CTA level: the aggregate of a warp is the value of its last position in scan order
  const int last_lane = REVERSE ? 0 : 31;
  __syncthreads();
  if (lane == last_lane) warp_buf[swarp] = mine;
  __shared__ int warp_valid[32];
  if (lane == last_lane) warp_valid[swarp] = valid ? 1 : 0;
  __syncthreads();
  if (warp == 0) {
    E w = warp_buf[lane < nw ? lane : 0];
    bool wv = lane < nw && warp_valid[lane < nw ? lane : 0] != 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const E other = shfl_elem<E>(w, (lane - off) & 31);
      const bool ov = __shfl_sync(0xffffffffu, wv ? 1 : 0, (lane - off) & 31) != 0;
      if (lane >= off && wv && ov) w = combine(other, w);
    }
    if (lane < nw) warp_buf[lane] = w;  // inclusive over warps (in scan order)
  }
  __syncthreads();
  // NOTE: a partially valid warp has its invalid positions at the END of the scan order (forward) -- they never feed
  // a valid one; in reverse order the caller makes invalid positions the identity and passes valid = true.
  bool have_prev = false;
  E prev;  // inclusive value of everything before this thread's warp inside the tile
  if (swarp > 0) { prev = warp_buf[swarp - 1]; have_prev = true; }
  // 3. across tiles: decoupled look-back
  const int stile = REVERSE ? n_tiles - 1 - tile : tile;  // position of this tile in scan order
  bool have_tile_prefix = false;
  if (n_tiles > 1) {
    if (threadIdx.x == 0) {
      const E agg = warp_buf[nw - 1];
      if (stile == 0) {
        lb.prefix[tile] = agg;
        __threadfence();
        atomicExch(&lb.status[tile], kStatusPrefix);
      } else {
        lb.agg[tile] = agg;
        __threadfence();
        atomicExch(&lb.status[tile], kStatusAgg);
        // walk back over the predecessors in scan order
        E acc;
        bool have = false;
        for (int s = stile - 1; s >= 0; --s) {
          const int p = REVERSE ? n_tiles - 1 - s : s;
          int st;
          while ((st = atomicAdd(&lb.status[p], 0)) == 0) { __nanosleep(20); }
          __threadfence();
          const E got = (st == kStatusPrefix) ? lb.prefix[p] : lb.agg[p];
          acc = have ? combine(got, acc) : got;
          have = true;
          if (st == kStatusPrefix) break;
        }
        *bcast = acc;  // exclusive prefix of this tile
        lb.prefix[tile] = combine(acc, agg);
        __threadfence();
        atomicExch(&lb.status[tile], kStatusPrefix);
      }
    }
    __syncthreads();
    have_tile_prefix = stile > 0;
  }
  // 4. the value just before this thread (for the caller) and the final inclusive value
  // previous position inside the warp
  const E up = shfl_elem<E>(mine, (REVERSE ? lane + 1 : lane - 1) & 31);
  E before;
  bool have_before = false;
  if (slane > 0) { before = up; have_before = true; }
  if (have_prev) {
    mine = combine(prev, mine);
    before = have_before ? combine(prev, before) : prev;
    have_before = true;
  }
  if (have_tile_prefix) {
    const E tp = *bcast;
    mine = combine(tp, mine);
    before = have_before ? combine(tp, before) : tp;
    have_before = true;
  }
  excl_valid = have_before;
  excl = before;
  __syncthreads();
}

// ---- the sweep ------------------------------------------------------------------------------------------------------
template <typename T> struct SweepArgs {
  auxssm_sweep_desc d;
  const T *x, *m0, *P0, *Fs, *Qs, *bs, *ys, *n_aux, *n_eps, *uniforms, *delta;
  T* x_out;
  int32_t* accepted;
  T* alpha_out;
  // scratch
  int* status;        // [3 scans][n_chains][n_tiles] + [n_chains] tickets + [n_chains] decisions (zeroed per launch)
  void* payload;      // look-back payloads
  T* partial;         // [n_chains][n_tiles][kNSums]
  T* xprop;           // [n_chains][T][d]
  int n_tiles, tile_len;
};
constexpr int kNSums = 9;

template <typename T, int D> __device__ __forceinline__ Mat<T, D> load_mat_g(const T* g, int d) {
  Mat<T, D> r = m_zero<T, D>();
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (i < d && j < d) r.a[i][j] = g[i * d + j];
  return r;
}
template <typename T, int D> __device__ __forceinline__ Vec<T, D> load_vec_g(const T* g, int d) {
  Vec<T, D> r = v_zero<T, D>();
#pragma unroll
  for (int i = 0; i < D; ++i)
    if (i < d) r.a[i] = g[i];
  return r;
}

// gradient and (diagonal) Hessian of the potential g(x; y) at x  (model.py:56-82 for SV)
template <typename T> __device__ __forceinline__ void pot_derivs(int potential, T sig, T y, T x, T& g, T& h, T& val) {
  if (potential == AUXSSM_POT_SV) {
    const T w = T(0.5) * y * y * exp(-x);
    g = w - T(0.5);
    h = -w;
    val = -w - T(0.5) * x - T(kHalfLog2Pi);
  } else if (potential == AUXSSM_POT_GAUSSIAN) {
    const T z = (y - x) / sig;
    g = z / sig;
    h = T(-1) / (sig * sig);
    val = T(-0.5) * z * z - log(sig) - T(kHalfLog2Pi);
  } else {
    g = h = val = T(0);
  }
  if (!(val == val)) val = T(0);  // nan_to_num(val), model.py:76
}

// auxiliary observation (aux_y, r = diag R) of auxiliary_kalman.py:28-48 at linearisation point xl
template <typename T, int D>
__device__ __forceinline__ void obs_factory(const auxssm_sweep_desc& dsc, const Vec<T, D>& xl, const Vec<T, D>& u,
                                            const Vec<T, D>& ydat, T delta, Vec<T, D>& ay, Vec<T, D>& r, int d) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    ay.a[i] = T(0);
    r.a[i] = T(1);
    if (i < d) {
      T g, h, val;
      pot_derivs<T>(dsc.potential, T(dsc.sig), ydat.a[i], xl.a[i], g, h, val);
      if (dsc.order == 1) {
        if (!(g == g)) g = T(0);  // nan_to_num(grad), auxiliary_kalman.py:30
        ay.a[i] = u.a[i] + T(0.5) * delta * g;
        r.a[i] = T(0.5) * delta;
      } else {
        const T omega = T(1) / (-h + T(2) / delta);
        ay.a[i] = omega * (T(2) * u.a[i] / delta + g - h * xl.a[i]);
        r.a[i] = omega;
      }
    }
  }
}

// sequential_update with H = I, diagonal R (filtering.py:84-130): returns ell increment, updates (m, P)
template <typename T, int D>
__device__ __forceinline__ T kf_update(Vec<T, D>& m, Mat<T, D>& P, const Vec<T, D>& y, const Vec<T, D>& r, int d) {
  Mat<T, D> S = P;
#pragma unroll
  for (int i = 0; i < D; ++i) S.a[i][i] += (i < d) ? r.a[i] : T(1);
  bool ok;
  const Mat<T, D> L = chol<T, D>(S, d, ok);
  const Vec<T, D> diff = v_sub<T, D>(y, m);
  const T ell = gauss_logpdf<T, D>(L, diff, d);
  const Mat<T, D> K = mm<T, D>(P, spd_inv<T, D>(S, d));
  m = v_add<T, D>(m, mv<T, D>(K, diff));
  P = m_sym<T, D>(m_sub<T, D>(P, mm<T, D>(K, mmT<T, D>(S, K))));
  return ell;
}

// scan element of time t >= 1 (_filtering_init_one, filtering.py:197-250, H = I, c = 0, diagonal R, no missing data)
template <typename T, int D>
__device__ __forceinline__ FiltElem<T, D> build_elem(const Mat<T, D>& F, const Mat<T, D>& Q, const Vec<T, D>& b,
                                                     const Vec<T, D>& y, const Vec<T, D>& r, const Vec<T, D>& m,
                                                     const Mat<T, D>& P, int d) {
  const Vec<T, D> mp = v_add<T, D>(mv<T, D>(F, m), b);
  const Mat<T, D> Pp = m_add<T, D>(mm<T, D>(F, mmT<T, D>(P, F)), Q);
  Mat<T, D> S = Pp;
#pragma unroll
  for (int i = 0; i < D; ++i) S.a[i][i] += (i < d) ? r.a[i] : T(1);
  const Mat<T, D> Si = spd_inv<T, D>(S, d);
  const Mat<T, D> K = mm<T, D>(Pp, Si);
  FiltElem<T, D> e;
  e.A = m_sub<T, D>(F, mm<T, D>(K, F));
  e.b = v_add<T, D>(mp, mv<T, D>(K, v_sub<T, D>(y, mp)));
  e.C = m_sym<T, D>(m_sub<T, D>(Pp, mm<T, D>(K, mmT<T, D>(S, K))));
  const Mat<T, D> tmp = mTm<T, D>(F, Si);  // F^T S^{-1}
  e.eta = mv<T, D>(tmp, v_sub<T, D>(y, b));
  e.J = m_sym<T, D>(mm<T, D>(tmp, F));
  // the padding block must stay inert under the combine: A, C, J zero there (b, eta already zero)
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (i >= d || j >= d) { e.A.a[i][j] = T(0); e.C.a[i][j] = T(0); e.J.a[i][j] = T(0); }
  return e;
}

template <typename T, int D> __device__ __forceinline__ T block_sum_all(T v, T* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = T(0);
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

template <typename T, int D>
__global__ void __launch_bounds__(256) kalman_sweep_fused_kernel(SweepArgs<T> a) {
  using FE = FiltElem<T, D>;
  using SE = SampElem<T, D>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FE* wbufF = reinterpret_cast<FE*>(smem_raw);            // 32 warp aggregates (also reused for SampElem)
  FE* bcastF = wbufF + 32;
  SE* wbufS = reinterpret_cast<SE*>(smem_raw);
  SE* bcastS = reinterpret_cast<SE*>(bcastF);
  T* red = reinterpret_cast<T*>(bcastF + 1);              // 32 words
  T* xnb = red + 32;                                      // [blockDim + 1][D] neighbour exchange of x'
  const auxssm_sweep_desc& dsc = a.d;
  const int d = dsc.lg.dx, Tn = dsc.lg.T, n_tiles = a.n_tiles, TL = a.tile_len;
  const int c = blockIdx.x / n_tiles, tile = blockIdx.x % n_tiles, j = threadIdx.x;
  const int t = 1 + tile * TL + j;                        // this thread's time step (t = 0 rides with thread (0, 0))
  const bool valid = j < TL && t < Tn;
  const bool first = (tile == 0 && j == 0);
  const T delta = a.delta[dsc.delta_stride ? c : 0];
  const T shd = sqrt(T(0.5) * delta);

  auto dynF = [&](int tm) { return load_mat_g<T, D>(a.Fs + (int64_t)c * dsc.lg.Fs.chain + (int64_t)tm * dsc.lg.Fs.time, d); };
  auto dynQ = [&](int tm) { return load_mat_g<T, D>(a.Qs + (int64_t)c * dsc.lg.Qs.chain + (int64_t)tm * dsc.lg.Qs.time, d); };
  auto dynb = [&](int tm) { return load_vec_g<T, D>(a.bs + (int64_t)c * dsc.lg.bs.chain + (int64_t)tm * dsc.lg.bs.time, d); };
  auto state = [&](const T* base, int tt) { return load_vec_g<T, D>(base + ((int64_t)c * Tn + tt) * d, d); };
  auto ydata = [&](int tt) {
    return a.ys ? load_vec_g<T, D>(a.ys + (int64_t)c * dsc.ys_chain_stride + (int64_t)tt * d, d) : v_zero<T, D>();
  };
  const Vec<T, D> m0 = load_vec_g<T, D>(a.m0 + (int64_t)c * dsc.lg.m0.chain, d);
  const Mat<T, D> P0 = load_mat_g<T, D>(a.P0 + (int64_t)c * dsc.lg.P0.chain, d);

  // per-time-step inputs: x_t, u_t = x_t + sqrt(delta/2) n_aux (generic.py:61), data, dynamics t-1 -> t
  const int tc = valid ? t : 0;
  Vec<T, D> xt = state(a.x, tc), ut, ydat = ydata(tc);
  {
    const Vec<T, D> na = state(a.n_aux, tc);
#pragma unroll
    for (int i = 0; i < D; ++i) ut.a[i] = xt.a[i] + shd * na.a[i];
  }
  const Mat<T, D> F = valid ? dynF(t - 1) : m_zero<T, D>();
  const Mat<T, D> Q = valid ? dynQ(t - 1) : m_eye<T, D>();
  const Vec<T, D> b = valid ? dynb(t - 1) : v_zero<T, D>();
  // t = 0 inputs (thread (0, 0) only uses them)
  const Vec<T, D> x0 = state(a.x, 0), yd0 = ydata(0);
  Vec<T, D> u0;
  {
    const Vec<T, D> na = state(a.n_aux, 0);
#pragma unroll
    for (int i = 0; i < D; ++i) u0.a[i] = x0.a[i] + shd * na.a[i];
  }

  const int64_t sidx = (int64_t)c * n_tiles;
  const int64_t scan_stride = (int64_t)gridDim.x;  // = n_chains * n_tiles
  // look-back state of scan `which` (0: forward filter, 1: reverse sampler, 2: forward filter of the reverse proposal)
#define AUXSSM_LB(E, which)                                                                                        \
  LookBack<E> {                                                                                                    \
    a.status + (which) * scan_stride + sidx,                                                                       \
        reinterpret_cast<E*>(static_cast<unsigned char*>(a.payload) + (size_t)(which) * 2 * scan_stride * sizeof(FE)) + sidx, \
        reinterpret_cast<E*>(static_cast<unsigned char*>(a.payload) + (size_t)(which) * 2 * scan_stride * sizeof(FE)) +      \
            scan_stride + sidx                                                                                     \
  }

  T sums[kNSums];
#pragma unroll
  for (int k = 0; k < kNSums; ++k) sums[k] = T(0);
  enum { S_ELL_F = 0, S_ELL_R, S_PRIOR_XP, S_PRIOR_X, S_LIK_F, S_LIK_R, S_POT_XP, S_POT_X, S_CORR };

  // one filter pass around linearisation trajectory xl (xl_t per thread, xl_0 for t = 0): returns ell contributions in
  // `ell_sum`, the filtering marginals of this thread's t in (mt, Pt) and of t = 0 in (m0f, P0f); (ay, r) are the
  // auxiliary observations of this thread's t, (ay0, r0) those of t = 0
  auto filter_pass = [&](int which, const Vec<T, D>& xl, const Vec<T, D>& xl0, Vec<T, D>& ay, Vec<T, D>& r,
                         Vec<T, D>& ay0, Vec<T, D>& r0, Vec<T, D>& mt, Mat<T, D>& Pt, Vec<T, D>& m0f, Mat<T, D>& P0f,
                         T& ell_sum) {
    obs_factory<T, D>(dsc, xl, ut, ydat, delta, ay, r, d);
    obs_factory<T, D>(dsc, xl0, u0, yd0, delta, ay0, r0, d);
    m0f = m0;
    P0f = P0;
    const T ell0 = kf_update<T, D>(m0f, P0f, ay0, r0, d);   // filtering.py:52 (every thread: a few dozen flops)
    FE e;
    if (valid) {
      e = first ? build_elem<T, D>(F, Q, b, ay, r, m0f, P0f, d)
                : build_elem<T, D>(F, Q, b, ay, r, v_zero<T, D>(), m_zero<T, D>(), d);
    } else {
      e.A = m_eye<T, D>(); e.C = m_zero<T, D>(); e.J = m_zero<T, D>(); e.b = v_zero<T, D>(); e.eta = v_zero<T, D>();
    }
    bool have_prev;
    FE prev;
    tile_scan<FE, false>(e, valid, tile, n_tiles, AUXSSM_LB(FE, which), wbufF, bcastF, have_prev, prev);
    mt = e.b;
    Pt = e.C;
    // ell increment of this step from the marginal of t - 1 (filtering.py:60-62)
    Vec<T, D> mp = have_prev ? prev.b : m0f;
    Mat<T, D> Pp = have_prev ? prev.C : P0f;
    T inc = T(0);
    if (valid) {
      mp = v_add<T, D>(mv<T, D>(F, mp), b);
      Pp = m_sym<T, D>(m_add<T, D>(mm<T, D>(F, mmT<T, D>(Pp, F)), Q));
      inc = kf_update<T, D>(mp, Pp, ay, r, d);
    }
    ell_sum = inc + (first ? ell0 : T(0));
  };

  // ------------------------------ forward proposal: filter around x ------------------------------
  Vec<T, D> ay, r, ay0, r0, mt, m0f;
  Mat<T, D> Pt, P0f;
  filter_pass(0, xt, x0, ay, r, ay0, r0, mt, Pt, m0f, P0f, sums[S_ELL_F]);

  // ------------------------------ backward sampler: reverse scan of (G_t, e_t) ------------------------------
  auto samp_elem = [&](const Vec<T, D>& m, const Mat<T, D>& P, int tt, const Vec<T, D>& eps) {
    SE s;
    if (tt == Tn - 1) {  // _sample_last_step, sampling.py:116-124
      bool ok;
      Mat<T, D> L = chol<T, D>(P, d, ok);
      if (!ok && d > 1) L = m_zero<T, D>();
      s.G = m_zero<T, D>();
      s.e = v_add<T, D>(m, mv<T, D>(L, eps));
    } else {             // mean_and_chol, sampling.py:61-105, transition tt -> tt + 1
      const Mat<T, D> Fn = dynF(tt), Qn = dynQ(tt);
      const Vec<T, D> bn = dynb(tt);
      const Mat<T, D> S = m_sym<T, D>(m_add<T, D>(mm<T, D>(Fn, mmT<T, D>(P, Fn)), Qn));
      Mat<T, D> Sp = S;
#pragma unroll
      for (int i = 0; i < D; ++i) if (i >= d) Sp.a[i][i] = T(1);
      const Mat<T, D> gain = mm<T, D>(mmT<T, D>(P, Fn), spd_inv<T, D>(Sp, d));   // P F^T S^{-1}
      const Mat<T, D> Sig = m_sym<T, D>(m_sub<T, D>(P, mm<T, D>(gain, mmT<T, D>(S, gain))));
      bool ok;
      Mat<T, D> L = chol<T, D>(Sig, d, ok);
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int k = 0; k < D; ++k) {
          if (!ok && d > 1) L.a[i][k] = T(0);           // jnp.linalg.cholesky -> NaN -> nan_to_num -> 0
          if (!(L.a[i][k] == L.a[i][k])) L.a[i][k] = T(0);
          if (i >= d || k >= d) L.a[i][k] = T(0);
        }
      s.G = gain;
      s.e = v_add<T, D>(v_sub<T, D>(m, mv<T, D>(gain, v_add<T, D>(mv<T, D>(Fn, m), bn))), mv<T, D>(L, eps));
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (i >= d || k >= d) s.G.a[i][k] = T(0);
    return s;
  };
  SE se;
  if (valid) {
    se = samp_elem(mt, Pt, t, state(a.n_eps, t));
  } else {  // identity map: positions past the end of the chain come FIRST in the reverse scan order
    se.G = m_eye<T, D>();
    se.e = v_zero<T, D>();
  }
  {
    bool hp;
    SE prev;
    tile_scan<SE, true>(se, true, tile, n_tiles, AUXSSM_LB(SE, 1), wbufS, bcastS, hp, prev);
  }
  Vec<T, D> xp = se.e;                       // x'_t
  Vec<T, D> xp0 = v_zero<T, D>();
  if (first) {                               // x'_0 = G_0 x'_1 + e_0
    const SE s0 = samp_elem(m0f, P0f, 0, state(a.n_eps, 0));
    xp0 = (Tn > 1) ? v_add<T, D>(mv<T, D>(s0.G, xp), s0.e) : s0.e;
  }

  // ------------------------------ publish x', fetch x'_{t-1} ------------------------------
  T* xpg = a.xprop + (int64_t)c * Tn * d;
  if (valid) {
#pragma unroll
    for (int i = 0; i < D; ++i) if (i < d) { xpg[(int64_t)t * d + i] = xp.a[i]; xnb[(j + 1) * D + i] = xp.a[i]; }
  }
  if (first) {
#pragma unroll
    for (int i = 0; i < D; ++i) if (i < d) { xpg[i] = xp0.a[i]; xnb[i] = xp0.a[i]; }
  }
  int* xflag = a.status + 3 * scan_stride + 2 * (int64_t)(gridDim.x / n_tiles) + sidx;  // x' of tile k published
  __threadfence();
  __syncthreads();
  if (j == 0) {
    atomicExch(&xflag[tile], 1);
    if (tile > 0) {
      while (atomicAdd(&xflag[tile - 1], 0) == 0) { __nanosleep(20); }
      __threadfence();
#pragma unroll
      for (int i = 0; i < D; ++i) if (i < d) xnb[i] = xpg[(int64_t)(t - 1) * d + i];
    }
  }
  __syncthreads();
  Vec<T, D> xp_prev = v_zero<T, D>(), x_prev = valid ? state(a.x, t - 1) : v_zero<T, D>();
#pragma unroll
  for (int i = 0; i < D; ++i) if (i < d) xp_prev.a[i] = xnb[j * D + i];

  // ------------------------------ reverse proposal: filter around x' ------------------------------
  Vec<T, D> ayr, rr, ayr0, rr0, mtr, m0r;
  Mat<T, D> Ptr_, P0r;
  if (first) {
#pragma unroll
    for (int i = 0; i < D; ++i) xnb[i] = xp0.a[i];
  }
  __syncthreads();
  // every thread needs x'_0 (the t = 0 update is recomputed everywhere): broadcast through global memory
  Vec<T, D> xp0_all = v_zero<T, D>();
  {
    if (tile > 0 && j == 0) { while (atomicAdd(&xflag[0], 0) == 0) { __nanosleep(20); } __threadfence(); }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < D; ++i) if (i < d) xp0_all.a[i] = (tile == 0) ? xnb[i] : xpg[i];
  }
  filter_pass(2, xp, xp0_all, ayr, rr, ayr0, rr0, mtr, Ptr_, m0r, P0r, sums[S_ELL_R]);

  // ------------------------------ trajectory log-densities and the MH correction ------------------------------
  auto diag_lik = [&](const Vec<T, D>& z, const Vec<T, D>& y, const Vec<T, D>& rv) {  // log N(y; z, diag r), base.py:138-166
    T s = T(0);
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i < d) { const T df = y.a[i] - z.a[i]; s += T(-0.5) * df * df / rv.a[i] - T(0.5) * log(rv.a[i]) - T(kHalfLog2Pi); }
    return (s == s) ? s : T(0);
  };
  auto pot_val = [&](const Vec<T, D>& z, const Vec<T, D>& yd) {
    T s = T(0);
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i < d) { T g, h, v; pot_derivs<T>(dsc.potential, T(dsc.sig), yd.a[i], z.a[i], g, h, v); s += v; }
    return s;
  };
  if (valid) {
    bool ok;
    const Mat<T, D> Lq = chol<T, D>(Q, d, ok);
    sums[S_PRIOR_XP] = gauss_logpdf<T, D>(Lq, v_sub<T, D>(xp, v_add<T, D>(mv<T, D>(F, xp_prev), b)), d);
    sums[S_PRIOR_X] = gauss_logpdf<T, D>(Lq, v_sub<T, D>(xt, v_add<T, D>(mv<T, D>(F, x_prev), b)), d);
    sums[S_LIK_F] = diag_lik(xp, ay, r);
    sums[S_LIK_R] = diag_lik(xt, ayr, rr);
    sums[S_POT_XP] = pot_val(xp, ydat);
    sums[S_POT_X] = pot_val(xt, ydat);
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i < d) { const T p = xp.a[i] - ut.a[i], q = xt.a[i] - ut.a[i]; sums[S_CORR] += p * p - q * q; }
  }
  if (first) {
    bool ok;
    const Mat<T, D> Lp = chol<T, D>(P0, d, ok);
    sums[S_PRIOR_XP] += gauss_logpdf<T, D>(Lp, v_sub<T, D>(xp0, m0), d);
    sums[S_PRIOR_X] += gauss_logpdf<T, D>(Lp, v_sub<T, D>(x0, m0), d);
    sums[S_LIK_F] += diag_lik(xp0, ay0, r0);
    sums[S_LIK_R] += diag_lik(x0, ayr0, rr0);
    sums[S_POT_XP] += pot_val(xp0, yd0);
    sums[S_POT_X] += pot_val(x0, yd0);
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i < d) { const T p = xp0.a[i] - u0.a[i], q = x0.a[i] - u0.a[i]; sums[S_CORR] += p * p - q * q; }
  }
  T tot[kNSums];
#pragma unroll
  for (int k = 0; k < kNSums; ++k) tot[k] = block_sum_all<T, D>(sums[k], red);

  // ------------------------------ accept / reject (generic.py:70-74, 98-106) ------------------------------
  const int n_chains = gridDim.x / n_tiles;
  int* ticket = a.status + 3 * scan_stride;
  int* decision = ticket + n_chains;
  __shared__ int acc_s;
  if (j == 0) {
    T* part = a.partial + ((int64_t)c * n_tiles + tile) * kNSums;
    for (int k = 0; k < kNSums; ++k) part[k] = tot[k];
    __threadfence();
    const int arrived = atomicAdd(&ticket[c], 1);
    if (arrived == n_tiles - 1) {  // last tile of the chain: deterministic sum over tiles, then the decision
      __threadfence();
      T s[kNSums];
      for (int k = 0; k < kNSums; ++k) s[k] = T(0);
      for (int q = 0; q < n_tiles; ++q)
        for (int k = 0; k < kNSums; ++k) s[k] += a.partial[((int64_t)c * n_tiles + q) * kNSums + k];
      const T lp_prop = s[S_LIK_F] - s[S_ELL_F] + s[S_PRIOR_XP], lt_prop = s[S_PRIOR_XP] + s[S_POT_XP];
      const T lp_rev = s[S_LIK_R] - s[S_ELL_R] + s[S_PRIOR_X], lt_rev = s[S_PRIOR_X] + s[S_POT_X];
      const T log_alpha = lt_prop - lt_rev + lp_rev - lp_prop - s[S_CORR] / delta;
      const T alpha = exp(log_alpha < T(0) ? log_alpha : T(0));
      const int acc = a.uniforms[c] < alpha ? 1 : 0;
      a.accepted[c] = acc;
      if (a.alpha_out) a.alpha_out[c] = alpha;
      __threadfence();
      atomicExch(&decision[c], 1 + acc);
    }
    int dec;
    while ((dec = atomicAdd(&decision[c], 0)) == 0) { __nanosleep(20); }
    acc_s = dec - 1;
  }
  __syncthreads();
  T* xo = a.x_out + (int64_t)c * Tn * d;
  if (valid) {
#pragma unroll
    for (int i = 0; i < D; ++i) if (i < d) xo[(int64_t)t * d + i] = acc_s ? xp.a[i] : xt.a[i];
  }
  if (first) {
#pragma unroll
    for (int i = 0; i < D; ++i) if (i < d) xo[i] = acc_s ? xp0.a[i] : x0.a[i];
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
struct SweepPlan {
  int n_tiles, tile_len, threads;
  size_t off_status, off_payload, off_partial, off_xprop, status_bytes, total;
};

template <typename T, int D> static SweepPlan make_sweep_plan(const auxssm_sweep_desc* d) {
  SweepPlan p;
  const int n_elem = d->lg.T > 1 ? d->lg.T - 1 : 1;
  // one element per thread; tiles of up to 256 steps; more tiles only while every CTA of the launch can be resident
  const int sms = device_sm_count();
  int tile = 256;
  int n_tiles = (n_elem + tile - 1) / tile;
  p.tile_len = tile;
  p.n_tiles = n_tiles;
  p.threads = n_elem < tile ? (n_elem + 31) / 32 * 32 : tile;
  if (p.n_tiles == 1) p.tile_len = p.threads;
  (void)sms;
  const size_t cn = (size_t)d->lg.n_chains * p.n_tiles;
  size_t off = 0;
  p.off_status = off;
  p.status_bytes = (3 * cn + 2 * (size_t)d->lg.n_chains + cn) * sizeof(int);
  off = (off + p.status_bytes + 255) / 256 * 256;
  p.off_payload = off; off = (off + 3 * 2 * cn * sizeof(FiltElem<T, D>) + 255) / 256 * 256;
  p.off_partial = off; off = (off + cn * kNSums * sizeof(T) + 255) / 256 * 256;
  p.off_xprop = off; off = (off + (size_t)d->lg.n_chains * d->lg.T * d->lg.dx * sizeof(T) + 255) / 256 * 256;
  p.total = off;
  return p;
}

template <typename T, int D> static size_t sweep_smem(int threads) {
  return 33 * sizeof(FiltElem<T, D>) + 32 * sizeof(T) + (size_t)(threads + 1) * D * sizeof(T) + 64;
}

static int check_sweep_desc(const auxssm_sweep_desc* d, const char* who) {
  if (!d) { set_error("%s: null descriptor", who); return AUXSSM_ERR_INVALID; }
  if (d->lg.T < 2 || d->lg.dx < 1 || d->lg.n_chains < 1) { set_error("%s: needs T >= 2, d >= 1, n_chains >= 1", who); return AUXSSM_ERR_INVALID; }
  if (d->lg.dx > 4) {
    set_error("%s: the fused sweep serves d <= 4 (got %d); compose auxssm_kalman_filter / _sample / _lgssm_logpdf instead", who, d->lg.dx);
    return AUXSSM_ERR_UNSUPPORTED;
  }
  if ((d->order != 1 && d->order != 2) || d->potential < 0 || d->potential > 2 || (d->delta_stride != 0 && d->delta_stride != 1)) {
    set_error("%s: order must be 1 or 2, potential a valid AUXSSM_POT_*, delta_stride 0 or 1", who);
    return AUXSSM_ERR_INVALID;
  }
  return AUXSSM_OK;
}

template <typename T, int D>
static int launch_sweep(cudaStream_t st, SweepArgs<T>& a, void* scratch, size_t scratch_bytes) {
  const char* who = "auxssm_kalman_sweep";
  const SweepPlan p = make_sweep_plan<T, D>(&a.d);
  if (!scratch || scratch_bytes < p.total) { set_error("%s: scratch too small (%zu < %zu bytes)", who, scratch_bytes, p.total); return AUXSSM_ERR_SCRATCH; }
  auto kern = kalman_sweep_fused_kernel<T, D>;
  const size_t smem = sweep_smem<T, D>(p.threads);
  int rc = ensure_smem(kern, smem, who);
  if (rc) return rc;
  const long grid = (long)a.d.lg.n_chains * p.n_tiles;
  if (p.n_tiles > 1) {  // tiles of one chain wait for each other: all CTAs must be co-resident
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, p.threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (grid > (long)per_sm * device_sm_count()) {
      set_error("%s: %ld CTAs (n_chains x time tiles) exceed the %ld that can be co-resident; use the composed path", who, grid,
                (long)per_sm * device_sm_count());
      return AUXSSM_ERR_UNSUPPORTED;
    }
  }
  unsigned char* sb = static_cast<unsigned char*>(scratch);
  a.status = reinterpret_cast<int*>(sb + p.off_status);
  a.payload = sb + p.off_payload;
  a.partial = reinterpret_cast<T*>(sb + p.off_partial);
  a.xprop = reinterpret_cast<T*>(sb + p.off_xprop);
  a.n_tiles = p.n_tiles;
  a.tile_len = p.tile_len;
  if (cudaMemsetAsync(a.status, 0, p.status_bytes, st) != cudaSuccess) { set_error("%s: cudaMemsetAsync failed", who); return AUXSSM_ERR_LAUNCH; }
  if (p.n_tiles > 1) {
    // tiles of one chain wait for each other (look-back flags, accept decision): a COOPERATIVE launch makes the
    // runtime guarantee that every CTA is resident (or refuse the launch) whatever else shares the GPU
    void* kargs[] = {&a};
    if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3((unsigned)grid), dim3(p.threads), kargs, smem, st) !=
        cudaSuccess) {
      set_error("%s: cooperative launch of %ld CTAs failed: %s", who, grid, cudaGetErrorString(cudaGetLastError()));
      return AUXSSM_ERR_LAUNCH;
    }
  } else {
    kern<<<(unsigned)grid, p.threads, smem, st>>>(a);
  }
  AUXSSM_CHECK_LAUNCH("auxssm_kalman_sweep (kalman_sweep_fused_kernel)");
  return AUXSSM_OK;
}

template <typename T>
static int sweep_entry(void* stream, const auxssm_sweep_desc* d, const T* x, const T* m0, const T* P0, const T* Fs,
                       const T* Qs, const T* bs, const T* ys, const T* n_aux, const T* n_eps, const T* uniforms,
                       const T* delta, T* x_out, int32_t* accepted, T* alpha, void* scratch, size_t scratch_bytes) {
  const char* who = "auxssm_kalman_sweep";
  int rc = check_sweep_desc(d, who);
  if (rc) return rc;
  if (!x || !m0 || !P0 || !Fs || !Qs || !bs || !n_aux || !n_eps || !uniforms || !delta || !x_out || !accepted ||
      (d->potential != AUXSSM_POT_FLAT && !ys)) {
    set_error("%s: null data pointer", who);
    return AUXSSM_ERR_INVALID;
  }
  SweepArgs<T> a;
  a.d = *d;
  a.x = x; a.m0 = m0; a.P0 = P0; a.Fs = Fs; a.Qs = Qs; a.bs = bs; a.ys = ys; a.n_aux = n_aux; a.n_eps = n_eps;
  a.uniforms = uniforms; a.delta = delta; a.x_out = x_out; a.accepted = accepted; a.alpha_out = alpha;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->lg.dx == 1) return launch_sweep<T, 1>(st, a, scratch, scratch_bytes);
  if (d->lg.dx == 2) return launch_sweep<T, 2>(st, a, scratch, scratch_bytes);
  return launch_sweep<T, 4>(st, a, scratch, scratch_bytes);
}

template <typename T> static size_t sweep_scratch(const auxssm_sweep_desc* d) {
  if (check_sweep_desc(d, "auxssm_kalman_sweep_scratch_bytes")) return 0;
  if (d->lg.dx == 1) return make_sweep_plan<T, 1>(d).total;
  if (d->lg.dx == 2) return make_sweep_plan<T, 2>(d).total;
  return make_sweep_plan<T, 4>(d).total;
}

}  // namespace auxssm

extern "C" __attribute__((visibility("default"))) size_t auxssm_kalman_sweep_scratch_bytes(const auxssm_sweep_desc* desc, int32_t elem_size) {
  return elem_size == 4 ? auxssm::sweep_scratch<float>(desc) : auxssm::sweep_scratch<double>(desc);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_sweep_f32(
    void* stream, const auxssm_sweep_desc* desc, const float* x, const float* m0, const float* P0, const float* Fs,
    const float* Qs, const float* bs, const float* ys, const float* normals_aux, const float* normals_eps,
    const float* uniforms, const float* delta, float* x_out, int32_t* accepted, float* alpha, void* scratch,
    size_t scratch_bytes) {
  return auxssm::sweep_entry<float>(stream, desc, x, m0, P0, Fs, Qs, bs, ys, normals_aux, normals_eps, uniforms, delta,
                                    x_out, accepted, alpha, scratch, scratch_bytes);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_sweep_f64(
    void* stream, const auxssm_sweep_desc* desc, const double* x, const double* m0, const double* P0, const double* Fs,
    const double* Qs, const double* bs, const double* ys, const double* normals_aux, const double* normals_eps,
    const double* uniforms, const double* delta, double* x_out, int32_t* accepted, double* alpha, void* scratch,
    size_t scratch_bytes) {
  return auxssm::sweep_entry<double>(stream, desc, x, m0, P0, Fs, Qs, bs, ys, normals_aux, normals_eps, uniforms, delta,
                                     x_out, accepted, alpha, scratch, scratch_bytes);
}

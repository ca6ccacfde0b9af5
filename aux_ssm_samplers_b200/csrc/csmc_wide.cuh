// Conditional SMC sweep for wide states (4 < d <= 32, N <= 256 particles): the per-particle d x d algebra
// of csmc_kernel.cuh restated as particle-matrix products in shared memory.
//
// Same algorithm, inputs, outputs and history layout as csmc_sweep_kernel (csmc.py:69-149,
// resamplings.py:14-37, auxiliary_guided_csmc.py:29-167); only the work decomposition differs.  With one
// thread per particle a d = 30 particle keeps ~9 d-vectors live (parent, normals, prediction, mean, noise,
// residuals ...) and spills; here the particles of a step are the rows of [NPAD][LD] shared-memory matrices
// and every d x d map is one product  Out = X M^T  in which a warp owns four output columns and a lane owns
// TM rows (register tile TM x 4, operands read as 16-byte vectors: the matrix rows are warp broadcasts, the
// particle rows are lane-strided with LD = D + 16/sizeof(T) so that they hit distinct banks).
//
//   xpred = XP F^T + b                       (skipped when F is diagonal: folded into the gather)
//   GUIDED      xn = xpred + (u - xpred) K^T + E cL^T        one pass, two operand streams
//   BOOTSTRAP   xn = xpred + E L^T
//   prior term  -0.5 |(xn - xpred) W^T|^2,   W = chol(Q)^{-1} (chol(P0)^{-1} at t = 0) inverted once per launch
//   backward    -0.5 |W (x_sel - b) - X_t (W F)^T|^2,        W F formed once per launch
//
// The proposal density of the free particles is -0.5 |eps|^2 (their residual is cL eps by construction); the
// pinned particle 0 gets the explicit triangular solve.  Normals, gains and the backward pass's history
// blocks arrive through cp.async one step ahead of their use.
#pragma once
#include "csmc_kernel.cuh"

namespace auxssm {

template <typename T, int D> struct WideCfg {
  static constexpr int V = 16 / sizeof(T);  // elements per 16-byte vector
  static constexpr int LD = D + V;          // row stride of every shared matrix
  static constexpr int NT = 256, NW = NT / 32;
  static constexpr int TN = (D >= 32) ? 8 : 4;  // output columns per warp (register tile TM x TN): every
                                                // lane-strided particle vector read feeds TN * V FMAs
  static constexpr int CT = D / TN;         // column tiles, one per warp
  static constexpr int PG = NW / CT;        // warps sharing a column tile split the particle rows
  static_assert(CT >= 1 && CT <= NW && NW % CT == 0, "unsupported D");
};

// compact global [rows][d] -> padded shared [rows][LD], element-wise asynchronous copies.
// magic = 2^32 / d + 1 gives e / d exactly for e < 2^13, d <= 32.
// Even d: two elements per copy (rows then start on 2-element boundaries on both sides).
template <typename T, int LD>
__device__ __forceinline__ void stage_rows_async(T* dst, const T* src, int total, int d, unsigned magic) {
  if ((d & 1) == 0) {
#pragma unroll 2
    for (int e = 2 * threadIdx.x; e < total; e += 2 * blockDim.x) {
      const int n = (int)__umulhi((unsigned)e, magic), k = e - n * d;
      cp_async<2 * sizeof(T)>(dst + n * LD + k, src + e);
    }
  } else {
#pragma unroll 2
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int n = (int)__umulhi((unsigned)e, magic), k = e - n * d;
      cp_async<sizeof(T)>(dst + n * LD + k, src + e);
    }
  }
}

template <typename T, int D, int LD> __device__ __forceinline__ void load_padded_ld(const T* g, T* s, int d) {
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
    const int i = e / D, k = e % D;
    s[i * LD + k] = (g != nullptr && i < d && k < d) ? g[i * d + k] : T(0);
  }
}

// 1 / diag (0 for padded or non-finite pivots) and the Gaussian normaliser of a lower factor
template <typename T, int D, int LD> __device__ __forceinline__ T factor_consts_ld(const T* L, T* invd, int d) {
  T cst = T(0);
  for (int i = 0; i < D; ++i) {
    const T lii = (i < d) ? L[i * LD + i] : T(0);
    const bool ok = (i < d) && isfinite(lii);
    if (threadIdx.x == 0) invd[i] = ok ? T(1) / lii : T(0);
    if (ok) cst -= log(fabs(lii)) + log_2pi_half<T>();
  }
  return cst;
}

// W = L^{-1}: thread c < D forward-substitutes column c (reads L, reads/writes only its own column of W)
template <typename T, int D, int LD> __device__ __forceinline__ void tri_inverse(const T* L, T* W, const T* invd, int d) {
  const int c = threadIdx.x;
  if (c < D) {
    for (int i = 0; i < D; ++i) {
      T acc = T(0);
      if (i >= c && i < d && c < d) {
        acc = (i == c) ? T(1) : T(0);
        for (int k = c; k < i; ++k) acc = fma_(-L[i * LD + k], W[k * LD + c], acc);
        acc *= invd[i];
      }
      W[i * LD + c] = acc;
    }
  }
}

// -0.5 |L^{-1} r|^2 by one full warp: lane l owns row l (its entry of r and row l of L in registers) and the
// columns are swept left to right, the finished entry z_k travelling by shuffle.  Every lane gets the result.
template <typename T, int D, int LD>
__device__ __forceinline__ T warp_quad_form(const T* L, const T* invd, T r_lane, int lane) {
  constexpr int V = WideCfg<T, D>::V;
  const int row = lane < D ? lane : 0;
  T lrow[D];
#pragma unroll
  for (int k = 0; k < D; k += V) {
    T l[V];
    ldv<T, V>(l, L + row * LD + k);
#pragma unroll
    for (int q = 0; q < V; ++q) lrow[k + q] = l[q];
  }
  const T inv = lane < D ? invd[row] : T(0);
  T r = lane < D ? r_lane : T(0), q2 = T(0);
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const T zk = __shfl_sync(0xffffffffu, r * inv, k);
    if (lane > k) r = fma_(-lrow[k], zk, r);
    q2 = fma_(zk, zk, q2);
  }
  return T(-0.5) * q2;
}

// acc[j][c] += sum_k x(r0 + 32 j, k) * M[c0 + c][k]; `xf(xv, row, k)` produces V operands of row `row`.
// Register double buffering: the operands of k-block kb + 1 are requested before the FMAs of block kb issue
// (two warps per scheduler cannot hide the shared-memory latency on their own).  Two blocks per trip keep the
// buffer indices static and the sweep inside the instruction cache.
template <typename T, int D, int TM, class XF>
__device__ __forceinline__ void wide_mm(const T* M, int r0, int c0, T (&acc)[TM][WideCfg<T, D>::TN], XF xf) {
  constexpr int V = WideCfg<T, D>::V, LD = WideCfg<T, D>::LD, NKB = D / V, TN = WideCfg<T, D>::TN;
  static_assert(NKB % 2 == 0, "even number of k-blocks expected");
  T mv[2][TN][V], xv[2][TM][V];
  auto load = [&](int b, int k) {
#pragma unroll
    for (int c = 0; c < TN; ++c) ldv<T, V>(mv[b][c], M + (c0 + c) * LD + k);
#pragma unroll
    for (int j = 0; j < TM; ++j) xf(xv[b][j], r0 + 32 * j, k);
  };
  auto mac = [&](int b) {
#pragma unroll
    for (int j = 0; j < TM; ++j) {
#pragma unroll
      for (int c = 0; c < TN; ++c) {
#pragma unroll
        for (int q = 0; q < V; ++q) acc[j][c] = fma_(xv[b][j][q], mv[b][c][q], acc[j][c]);
      }
    }
  };
  load(0, 0);
#pragma unroll 1
  for (int kb = 0; kb < NKB; kb += 2) {
    load(1, (kb + 1) * V);
    mac(0);
    if (kb + 2 < NKB) load(0, (kb + 2) * V);
    mac(1);
  }
}

template <typename T, int D> __host__ __device__ inline size_t csmc_wide_smem_elems(int NPAD) {
  constexpr int LD = WideCfg<T, D>::LD, CT = WideCfg<T, D>::CT;
  return (size_t)4 * NPAD * LD + (size_t)6 * D * LD + 16 * D + (size_t)CT * NPAD + NPAD + 128 + 32;
}

template <typename T, int D, int PROP, int NPAD>
__global__ void __launch_bounds__(256, 1) csmc_wide_kernel(CsmcArgs<T> a) {
  using Cfg = WideCfg<T, D>;
  constexpr int V = Cfg::V, LD = Cfg::LD, NT = Cfg::NT, CT = Cfg::CT, PG = Cfg::PG, TM = NPAD / (32 * PG),
                TN = Cfg::TN;
  static_assert(TM >= 1 && NPAD <= NT && NPAD % (32 * PG) == 0, "unsupported NPAD");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const auxssm_csmc_desc& dsc = a.d;
  const int N = dsc.N, d = dsc.d, Tn = dsc.T, i = threadIdx.x;
  const int lane = i & 31, warp = i >> 5;
  const int c0 = TN * (warp % CT), r0 = (warp / CT) * (32 * TM) + lane;  // this thread's tile of every product
  const bool active = i < N;
  const unsigned magic = 0xffffffffu / (unsigned)d + 1u;
  T* p = reinterpret_cast<T*>(smem_raw);
  {
    const size_t tot = csmc_wide_smem_elems<T, D>(NPAD);
    for (size_t e = i; e < tot; e += NT) p[e] = T(0);
  }
  T* bufA = p; p += NPAD * LD;
  T* bufB = p; p += NPAD * LD;
  T* bufE = p; p += NPAD * LD;
  T* bufR = p; p += NPAD * LD;  // xn - xpred, the operand of the prior-density product
  T* mF = p; p += D * LD;
  T* mW = p; p += D * LD;    // chol(Q)^{-1}
  T* mW0 = p; p += D * LD;   // chol(P0)^{-1}
  T* mWF = p; p += D * LD;   // W F
  T* mK = p; p += D * LD;    // GUIDED: K_t          BOOTSTRAP: chol(Q)
  T* mcL = p; p += D * LD;   // GUIDED: chol(Lam_t)  BOOTSTRAP: chol(P0)
  T* vb = p; p += D; T* vm0 = p; p += D; T* invq = p; p += D; T* invp0 = p; p += D; T* invl = p; p += D;
  T* vy = p; p += D; T* vxs = p; p += D; T* vu = p; p += D; T* vmean = p; p += D; T* vsel = p; p += D;
  T* vmu0 = p; p += D; T* vf = p; p += D; p += 4 * D;
  T* zq = p; p += CT * NPAD;
  T* cdf = p; p += NPAD;
  T* red = p; p += 128;
  T* misc = p; p += 32;  // [0] max weight (killing)  [1] cst_l  [8] selected index (int)
  int* isel = reinterpret_cast<int*>(misc + 8);
  __syncthreads();

  // ---- one-off setup: model image, inverse factors, W F ------------------------------------------------
  load_padded_ld<T, D, LD>(a.F, mF, d);
  load_padded_ld<T, D, LD>(a.Lq, mK, d);
  load_padded_ld<T, D, LD>(a.Lp0, mcL, d);
  if (i < D) {
    vb[i] = (i < d) ? a.b[i] : T(0);
    vm0[i] = (i < d) ? a.m0[i] : T(0);
  }
  __syncthreads();
  const T cst_q = factor_consts_ld<T, D, LD>(mK, invq, d);
  const T cst_p0 = factor_consts_ld<T, D, LD>(mcL, invp0, d);
  int offdiag = 0;
  for (int e = i; e < D * D; e += NT) {
    const int r = e / D, k = e % D;
    if (r != k && mF[r * LD + k] != T(0)) offdiag = 1;
  }
  const bool fdiag = __syncthreads_or(offdiag) == 0;  // also publishes invq / invp0
  tri_inverse<T, D, LD>(mK, mW, invq, d);
  tri_inverse<T, D, LD>(mcL, mW0, invp0, d);
  if (i < D) vf[i] = mF[i * LD + i];
  __syncthreads();
  for (int e = i; e < D * D; e += NT) {
    const int r = e / D, k = e % D;
    T acc = T(0);
    for (int j = 0; j < D; ++j) acc = fma_(mW[r * LD + j], mF[j * LD + k], acc);
    mWF[r * LD + k] = acc;
  }
  T cst_l = T(0);
  __syncthreads();

  auto plain = [&](const T* X) {
    return [X](T (&xv)[V], int row, int k) { ldv<T, V>(xv, X + row * LD + k); };
  };

  for (int c = blockIdx.x; c < dsc.n_chains; c += gridDim.x) {
    const int64_t hs = a.hist_per_chain ? c : blockIdx.x;  // history slot
    T* xs_h = a.xs_hist + hs * (int64_t)Tn * N * d;
    T* lw_h = a.lw_hist + hs * (int64_t)Tn * N;
    int32_t* a_h = a.a_hist + hs * (int64_t)(Tn > 1 ? Tn - 1 : 0) * N;
    const T* xstar = a.x_star + (int64_t)c * Tn * d;
    const T* ysc = a.ys ? a.ys + (int64_t)c * dsc.ys_chain_stride : nullptr;
    // normals of step tt -> bufE ([N][LD] rows): asynchronous copy of the external array, or generated in place
    auto stage_normals = [&](int tt) {
      const int64_t e0 = ((int64_t)c * Tn + tt) * N * d;
      if (a.rng.on) {
        constexpr int PER = PhiloxPer<T>::N;
        const uint64_t b0 = (uint64_t)e0 / PER, b1 = (uint64_t)(e0 + (int64_t)N * d - 1) / PER;
        const uint64_t call = a.call0();
        for (uint64_t blk = b0 + threadIdx.x; blk <= b1; blk += blockDim.x) {
          T v[PER];
          philox_block(PHILOX_NORMAL, a.rng.seed, call, blk, v);
#pragma unroll
          for (int q = 0; q < PER; ++q) {
            const int64_t e = (int64_t)(blk * PER + q) - e0;
            if (e >= 0 && e < (int64_t)N * d) {
              const int n = (int)__umulhi((unsigned)e, magic), k = (int)e - n * d;
              bufE[n * LD + k] = v[q];
            }
          }
        }
      } else {
        stage_rows_async<T, LD>(bufE, a.normals + e0, N * d, d, magic);
      }
    };
    auto stage_gains = [&](int tt) {
      if (PROP == AUXSSM_PROP_GUIDED && tt < Tn && (tt <= 1 || dsc.gains_time_stride != 0)) {
        stage_rows_async<T, LD>(mK, a.K + (int64_t)tt * dsc.gains_time_stride, d * d, d, magic);
        stage_rows_async<T, LD>(mcL, a.cL + (int64_t)tt * dsc.gains_time_stride, d * d, d, magic);
      }
    };
    T w = T(0);
    // register prefetch of the per-step scalars (one step ahead, as in csmc_sweep_kernel)
    T ur_n = T(0), uk_n = T(0), uw_n = T(0), xs_n = T(0), y_n = T(0), u_n = T(0), g_n = T(0), s_n = T(1);
    auto prefetch = [&](int tt) {
      const bool go = tt < Tn;
      if (go && active && tt > 0 && dsc.resampling == AUXSSM_RESAMPLE_SYSTEMATIC) {
        const int64_t us = ((int64_t)c * (Tn - 1) + (tt - 1)) * 3;  // (U, V, W), broadcast loads
        ur_n = a.draw_u_res(us); uk_n = a.draw_u_res(us + 1); uw_n = a.draw_u_res(us + 2);
      } else if (go && active && tt > 0) {
        ur_n = a.draw_u_res(((int64_t)c * (Tn - 1) + (tt - 1)) * N + i);
        if (dsc.resampling == AUXSSM_RESAMPLE_KILLING) uk_n = a.draw_u_kill(((int64_t)c * (Tn - 1) + (tt - 1)) * N + i);
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
    stage_normals(0);
    stage_gains(0);
    T* part = bufA;   // particles of the previous step (gather source)
    T* other = bufB;

    // ------------------------------ forward conditional particle filter ------------------------------
    for (int t = 0; t < Tn; ++t) {
      const bool first = (t == 0);
      const T ur = ur_n, uk = uk_n, uw = uw_n, xs_c = xs_n, y_c = y_n, u_c = u_n, g_c = g_n, s_c = s_n;
      prefetch(t + 1);
      // (1) conditional resampling, (2) parent gather, xpred
      int anc = 0;
      if (!first && active) {
        const T r = (dsc.resampling == AUXSSM_RESAMPLE_SYSTEMATIC) ? systematic_plan<T>(cdf, N, ur, uk, uw).r(i, N)
                                                                   : cdf[N - 1] * (T(1) - ur);
        anc = search_left<T>(cdf, r, N);
        if (dsc.resampling == AUXSSM_RESAMPLE_KILLING && uk * misc[0] < w) anc = i;
        if (i == 0) anc = 0;
      }
      if (i < D) {
        vxs[i] = xs_c;
        if (ysc) vy[i] = y_c;
        if (PROP != AUXSSM_PROP_BOOTSTRAP) {
          vu[i] = u_c;
          vmean[i] = u_c + s_c * s_c * g_c;
        }
      }
      T *xpb, *xnb;
      if (first || fdiag) {
        if (active) {
#pragma unroll
          for (int k = 0; k < D; k += V) {
            T v[V];
            if (first) {
              ldv<T, V>(v, vm0 + k);
            } else {
              T xp[V], fv[V], bv[V];
              ldv<T, V>(xp, part + anc * LD + k);
              ldv<T, V>(fv, vf + k);
              ldv<T, V>(bv, vb + k);
#pragma unroll
              for (int q = 0; q < V; ++q) v[q] = fma_(fv[q], xp[q], bv[q]);
            }
            stv<T, V>(other + i * LD + k, v);
          }
        }
        xpb = other; xnb = part;
      } else {
        if (active) {
#pragma unroll
          for (int k = 0; k < D; k += V) {
            T xp[V];
            ldv<T, V>(xp, part + anc * LD + k);
            stv<T, V>(other + i * LD + k, xp);
          }
        }
        __syncthreads();
        T acc[TM][TN];
#pragma unroll
        for (int j = 0; j < TM; ++j)
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) acc[j][cc] = T(0);
        wide_mm<T, D, TM>(mF, r0, c0, acc, plain(other));
        T bv[TN];
        ldv<T, TN>(bv, vb + c0);
#pragma unroll
        for (int j = 0; j < TM; ++j) {
          T o[TN];
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) o[cc] = acc[j][cc] + bv[cc];
          stv<T, TN>(part + (r0 + 32 * j) * LD + c0, o);
        }
        xpb = part; xnb = other;
      }
      cp_async_wait_all();  // normals (and gains) of this step, requested one step ago
      __syncthreads();

      // (3) proposal
      const bool new_gains = PROP == AUXSSM_PROP_GUIDED && (t <= 1 || dsc.gains_time_stride != 0);
      if (new_gains && warp == 0) {
        // 1 / diag(cL) and the normaliser of N(.; mu, Lam_t) (non-finite pivots dropped, mvn/base.py:16-58)
        T v = T(0);
        if (lane < D) {
          const T lii = (lane < d) ? mcL[lane * LD + lane] : T(0);
          const bool ok = (lane < d) && isfinite(lii);
          invl[lane] = ok ? T(1) / lii : T(0);
          v = ok ? -(log(fabs(lii)) + log_2pi_half<T>()) : T(0);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) misc[1] = v;
      }
      T qe = T(0);
      if (PROP == AUXSSM_PROP_GUIDED && active) {
#pragma unroll
        for (int k = 0; k < D; k += V) {
          T e[V];
          ldv<T, V>(e, bufE + i * LD + k);
#pragma unroll
          for (int q = 0; q < V; ++q) qe = fma_(e[q], e[q], qe);
        }
      }
      if (PROP == AUXSSM_PROP_INDEPENDENT) {
        if (active) {
#pragma unroll
          for (int k = 0; k < D; k += V) {
            T e[V], mu[V], o[V], xa[V], r[V];
            ldv<T, V>(e, bufE + i * LD + k);
            ldv<T, V>(mu, (i == 0) ? vxs + k : vmean + k);
            ldv<T, V>(xa, xpb + i * LD + k);
#pragma unroll
            for (int q = 0; q < V; ++q) {
              o[q] = (i == 0) ? mu[q] : fma_(s_c, e[q], mu[q]);
              r[q] = o[q] - xa[q];
            }
            stv<T, V>(xnb + i * LD + k, o);
            stv<T, V>(bufR + i * LD + k, r);
          }
        }
      } else {
        T acc[TM][TN];
#pragma unroll
        for (int j = 0; j < TM; ++j)
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) acc[j][cc] = T(0);
        if (PROP == AUXSSM_PROP_GUIDED) {
          const T* X = xpb;
          const T* uu = vu;
          wide_mm<T, D, TM>(mK, r0, c0, acc, [X, uu](T (&xv)[V], int row, int k) {
            T xa[V], ub[V];
            ldv<T, V>(xa, X + row * LD + k);
            ldv<T, V>(ub, uu + k);
#pragma unroll
            for (int q = 0; q < V; ++q) xv[q] = ub[q] - xa[q];
          });
          if (r0 == 0) {  // mean of the pinned particle's proposal, for its explicit density
            T xa[TN], o[TN];
            ldv<T, TN>(xa, xpb + c0);
#pragma unroll
            for (int cc = 0; cc < TN; ++cc) o[cc] = xa[cc] + acc[0][cc];
            stv<T, TN>(vmu0 + c0, o);
          }
          wide_mm<T, D, TM>(mcL, r0, c0, acc, plain(bufE));
        } else {
          wide_mm<T, D, TM>(first ? mcL : mK, r0, c0, acc, plain(bufE));
        }
#pragma unroll
        for (int j = 0; j < TM; ++j) {
          const int row = r0 + 32 * j;
          T xa[TN], o[TN];
          ldv<T, TN>(xa, xpb + row * LD + c0);
          if (row == 0) {
            ldv<T, TN>(o, vxs + c0);
          } else {
#pragma unroll
            for (int cc = 0; cc < TN; ++cc) o[cc] = xa[cc] + acc[j][cc];
          }
          stv<T, TN>(xnb + row * LD + c0, o);
          if (PROP == AUXSSM_PROP_GUIDED) {
#pragma unroll
            for (int cc = 0; cc < TN; ++cc) xa[cc] = o[cc] - xa[cc];
            stv<T, TN>(bufR + row * LD + c0, xa);
          }
        }
      }
      __syncthreads();
      if (t + 1 < Tn) stage_normals(t + 1);
      if (new_gains) cst_l = misc[1];

      // (4) prior residuals  z = (xn - xpred) W^T, history
      T pin = T(0);
      if (PROP != AUXSSM_PROP_BOOTSTRAP) {
        if (PROP == AUXSSM_PROP_GUIDED && warp == 0)
          pin = warp_quad_form<T, D, LD>(mcL, invl, lane < D ? vxs[lane] - vmu0[lane] : T(0), lane);
        T acc[TM][TN];
#pragma unroll
        for (int j = 0; j < TM; ++j)
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) acc[j][cc] = T(0);
        wide_mm<T, D, TM>(first ? mW0 : mW, r0, c0, acc, plain(bufR));
#pragma unroll
        for (int j = 0; j < TM; ++j) {
          T q2 = T(0);
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) q2 = fma_(acc[j][cc], acc[j][cc], q2);
          zq[(warp % CT) * NPAD + r0 + 32 * j] = q2;
        }
      }
      {
        T* xg = xs_h + (int64_t)t * N * d;
        if ((d & 1) == 0) {
#pragma unroll 2
          for (int e = 2 * i; e < N * d; e += 2 * NT) {
            const int n = (int)__umulhi((unsigned)e, magic), k = e - n * d;
            T v[2];
            ldv<T, 2>(v, xnb + n * LD + k);
            stv<T, 2>(xg + e, v);
          }
        } else {
#pragma unroll 2
          for (int e = i; e < N * d; e += NT) {
            const int n = (int)__umulhi((unsigned)e, magic), k = e - n * d;
            xg[e] = xnb[n * LD + k];
          }
        }
      }
      __syncthreads();
      stage_gains(t + 1);  // K / cL of this step are dead from here on

      // (5) weights
      T lw = T(0);
      if (active) {
        // potential and isotropic terms in one rolled pass over the particle's row (padded entries are zero
        // in every operand and contribute exactly zero)
        T pot = T(0), su = T(0), sm = T(0);
        const T inv_s = T(1) / s_c, inv_sig = T(1) / T(dsc.sig);
#pragma unroll 1
        for (int k = 0; k < D; k += V) {
          T x[V], y[V], uu[V], mm[V];
          ldv<T, V>(x, xnb + i * LD + k);
          ldv<T, V>(y, vy + k);
          if (PROP != AUXSSM_PROP_BOOTSTRAP) ldv<T, V>(uu, vu + k);
          if (PROP == AUXSSM_PROP_INDEPENDENT) ldv<T, V>(mm, vmean + k);
#pragma unroll
          for (int q = 0; q < V; ++q) {
            if (dsc.potential == AUXSSM_POT_SV) {
              pot = fma_(T(-0.5) * y[q] * y[q], exp(-x[q]), pot) - T(0.5) * x[q];
            } else if (dsc.potential == AUXSSM_POT_GAUSSIAN) {
              const T z = (y[q] - x[q]) * inv_sig;
              pot = fma_(T(-0.5) * z, z, pot);
            }
            if (PROP != AUXSSM_PROP_BOOTSTRAP) {
              const T zu = (x[q] - uu[q]) * inv_s;
              su = fma_(zu, zu, su);
            }
            if (PROP == AUXSSM_PROP_INDEPENDENT) {
              const T zm = (x[q] - mm[q]) * inv_s;
              sm = fma_(zm, zm, sm);
            }
          }
        }
        if (dsc.potential == AUXSSM_POT_SV) pot -= T(d) * log_2pi_half<T>();
        if (dsc.potential == AUXSSM_POT_GAUSSIAN) pot -= T(d) * (log(T(dsc.sig)) + log_2pi_half<T>());
        lw = pot;
        if (PROP != AUXSSM_PROP_BOOTSTRAP) {
          T q2 = T(0);
#pragma unroll
          for (int ct = 0; ct < CT; ++ct) q2 += zq[ct * NPAD + i];
          lw += T(-0.5) * q2 + (first ? cst_p0 : cst_q);
        }
        if (PROP == AUXSSM_PROP_INDEPENDENT && (dsc.gradient == 2 || (dsc.gradient == 1 && first))) lw += T(-0.5) * (su - sm);  // see csmc_kernel.cuh
        if (PROP == AUXSSM_PROP_GUIDED) {
          lw += T(-0.5) * su - T(d) * (log(s_c) + log_2pi_half<T>());
          lw -= ((i == 0) ? pin : T(-0.5) * qe) + cst_l;
        }
        lw_h[(int64_t)t * N + i] = lw;
        if (!first) a_h[(int64_t)(t - 1) * N + i] = anc;
      }
      w = block_normalize<T>(lw, active, red);
      const T cd = block_cdf<T>(w, red + 64);
      if (active) cdf[i] = cd;
      if (dsc.resampling == AUXSSM_RESAMPLE_KILLING) {
        const T wmax = block_max<T>(w, red + 96);
        if (i == 0) misc[0] = wmax;
      }
      __syncthreads();
      part = xnb;
      other = xpb;
    }

    // ------------------------------------ backward pass ----------------------------------------------
    T* xo = a.x_out + (int64_t)c * Tn * d;
    int32_t* ao = a.anc + (int64_t)c * Tn;
    const int64_t ub = (int64_t)c * Tn;
    BwdCache<T> bwdc;
    T* cur = other;  // staging buffers of the backward pass: the two particle buffers, `part` read first
    T* nxt = part;
    {
      const T r = cdf[N - 1] * (T(1) - a.draw_u_bwd(ub, bwdc));
      int B = __syncthreads_count(active && cdf[i] < r);
      if (B > N - 1) B = N - 1;
      if (i < D) {
        const T v = part[B * LD + i];
        vsel[i] = v;
        if (i < d) xo[(int64_t)(Tn - 1) * d + i] = v;
      }
      if (i == 0) { ao[Tn - 1] = B; isel[0] = B; }
      __syncthreads();
    }
    if (dsc.backward) {
      if (Tn > 1) stage_rows_async<T, LD>(cur, xs_h + (int64_t)(Tn - 2) * N * d, N * d, d, magic);
      T ub_n = (Tn > 1) ? a.draw_u_bwd(ub + 1, bwdc) : T(0);
      T lwh_n = (Tn > 1 && active) ? lw_h[(int64_t)(Tn - 2) * N + i] : T(0);
      for (int t = Tn - 2, k2 = 1; t >= 0; --t, ++k2) {
        const T ub_c = ub_n, lwh = lwh_n;
        if (t > 0) {
          ub_n = a.draw_u_bwd(ub + k2 + 1, bwdc);
          if (active) lwh_n = lw_h[(int64_t)(t - 1) * N + i];
        }
        cp_async_wait_all();
        __syncthreads();  // X_t staged; everyone is done with `nxt`
        if (t > 0) stage_rows_async<T, LD>(nxt, xs_h + (int64_t)(t - 1) * N * d, N * d, d, magic);
        // v0 = W (x_sel - b) for this warp's four columns: lane = (k-slice, column), butterfly over slices
        T v0[TN];
        {
          constexpr int NS = 32 / TN, KS = D / NS;  // k-slices per warp, k's per lane
          const int cc = lane % TN, kq = lane / TN;
          T pv = T(0);
#pragma unroll
          for (int q = 0; q < KS; ++q) {
            const int k = kq * KS + q;
            pv = fma_(mW[(c0 + cc) * LD + k], vsel[k] - vb[k], pv);
          }
#pragma unroll
          for (int off = TN; off < 32; off <<= 1) pv += __shfl_xor_sync(0xffffffffu, pv, off);
#pragma unroll
          for (int q = 0; q < TN; ++q) v0[q] = __shfl_sync(0xffffffffu, pv, q);
        }
        T acc[TM][TN];
#pragma unroll
        for (int j = 0; j < TM; ++j)
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) acc[j][cc] = T(0);
        wide_mm<T, D, TM>(mWF, r0, c0, acc, plain(cur));
#pragma unroll
        for (int j = 0; j < TM; ++j) {
          T q2 = T(0);
#pragma unroll
          for (int cc = 0; cc < TN; ++cc) {
            const T z = v0[cc] - acc[j][cc];
            q2 = fma_(z, z, q2);
          }
          zq[(warp % CT) * NPAD + r0 + 32 * j] = q2;
        }
        __syncthreads();
        T lw = T(0);
        if (active) {
          T q2 = T(0);
#pragma unroll
          for (int ct = 0; ct < CT; ++ct) q2 += zq[ct * NPAD + i];
          lw = T(-0.5) * q2 + cst_q + lwh;
        }
        const T wb = block_normalize<T>(lw, active, red);
        const T cd = block_cdf<T>(wb, red + 64);
        if (active) cdf[i] = cd;
        __syncthreads();
        const T r = cdf[N - 1] * (T(1) - ub_c);
        int B = __syncthreads_count(active && cdf[i] < r);
        if (B > N - 1) B = N - 1;
        if (i < D) {  // everyone passed the count barrier after reading vsel in v0
          const T v = cur[B * LD + i];
          vsel[i] = v;
          if (i < d) xo[(int64_t)t * d + i] = v;
        }
        if (i == 0) ao[t] = B;
        T* sw = cur; cur = nxt; nxt = sw;
      }
      __syncthreads();
    } else {
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
    cp_async_wait_all();
    __syncthreads();
  }
}

}  // namespace auxssm

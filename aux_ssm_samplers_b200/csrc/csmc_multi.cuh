// Conditional SMC sweep for narrow states and large particle counts (d <= 4, N > 256, N % 4 == 0): four
// consecutive particles per thread.
//
// Same algorithm, inputs, outputs and history layout as csmc_sweep_kernel (csmc.py:69-149,
// resamplings.py:14-37); only the work decomposition differs.  With one thread per particle a 1024-particle
// step is 32 warps meeting at ~10 block barriers, each warp paying the full two-level shuffle reductions and a
// latency-bound binary search for ONE particle per lane.  Here a thread owns particles 4i .. 4i+3:
//   * the draws and the history move as 16-byte vectors,
//   * max / sum / scan run thread-locally over the four values first, so the shuffle and barrier cost per
//     particle drops 4x (8 warps per chain instead of 32),
//   * the four inverse-CDF searches of a thread are branch-free power-of-two descents issued interleaved
//     (4 independent shared-memory loads in flight instead of a serial chain),
//   * the backward pass reads the particles this thread itself wrote (no staging through shared memory) and
//     prefetches the next step's block into registers.
// The summation order of the weights differs from csmc_sweep_kernel's (thread-local partial sums first); the
// oracle's canonical order is that of the one-particle-per-thread kernel, so agreement of the ancestor indices
// with the oracle is exact away from ties of (1 - u) * total against a CDF value at rounding level.
#pragma once
#include "csmc_kernel.cuh"

namespace auxssm {

template <int D> __host__ __device__ inline size_t csmc_multi_smem_elems(int N, int d, int prop) {
  const size_t mats = (prop == AUXSSM_PROP_GUIDED) ? 5 : 3;
  return mats * D * D + 16 * D + (size_t)N * d + N + 256;
}

// Inclusive block scan of one value per thread (same two-level Hillis-Steele pattern as block_cdf) that also
// returns the exclusive prefix, formed from the same operands so that excl(i + 1) == incl(i) inside a warp.
template <typename T> __device__ __forceinline__ T block_scan_excl(T v, T* red, T& excl) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncwarp();
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const T o = __shfl_up_sync(0xffffffffu, v, off);
    if (lane >= off) v = v + o;
  }
  T prev = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane == 0) prev = T(0);
  if (lane == 31) red[w] = v;
  __syncthreads();
  T tot = (lane < nw) ? red[lane] : T(0);
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const T o = __shfl_up_sync(0xffffffffu, tot, off);
    if (lane >= off) tot = tot + o;
  }
  const T offs = __shfl_sync(0xffffffffu, tot, (w > 0) ? w - 1 : 0);
  excl = (w > 0) ? prev + offs : prev;
  return (w > 0) ? v + offs : v;
}

template <typename T, int D, int PROP>
// d = 1 (configs[2]): four CTAs per SM (64 registers) -- the in-kernel Philox path must not halve the occupancy
__global__ void __launch_bounds__(256, (D == 1 && sizeof(T) == 4) ? 4 : 1) csmc_multi_kernel(CsmcArgs<T> a) {
  constexpr int PPT = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const auxssm_csmc_desc& dsc = a.d;
  const int N = dsc.N, d = dsc.d, Tn = dsc.T, i = threadIdx.x;
  const int p0 = PPT * i;
  const bool act = p0 < N;  // N % 4 == 0: a thread's four particles are all in range or all out
  T* p = reinterpret_cast<T*>(smem_raw);
  SModel<T, D> m;
  m.F = p; p += D * D;
  m.Lq = p; p += D * D;
  m.Lp0 = p; p += D * D;
  m.K = m.cL = nullptr;
  if (PROP == AUXSSM_PROP_GUIDED) { m.K = p; p += D * D; m.cL = p; p += D * D; }
  m.b = p; p += D; m.m0 = p; p += D; m.invq = p; p += D; m.invp0 = p; p += D; m.invl = p; p += D;
  m.y = p; p += D; m.xs_t = p; p += D; m.u_t = p; p += D; m.mean_t = p; p += D; m.xsel = p; p += D;
  p += 6 * D;
  T* red = p; p += 128;
  T* misc = p; p += 32;
  int* isel = reinterpret_cast<int*>(misc + 8);
  p = reinterpret_cast<T*>(smem_raw) + ((p - reinterpret_cast<T*>(smem_raw)) + 3) / 4 * 4;  // 16-byte vectors below
  T* cdf = p; p += N;
  T* part = p;  // [N][d]

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
  int hb = 1;  // largest power of two <= N
  while (2 * hb <= N) hb *= 2;
  __syncthreads();

  for (int c = blockIdx.x; c < dsc.n_chains; c += gridDim.x) {
    const int64_t hs = a.hist_per_chain ? c : blockIdx.x;
    T* xs_h = a.xs_hist + hs * (int64_t)Tn * N * d;
    T* lw_h = a.lw_hist + hs * (int64_t)Tn * N;
    int32_t* a_h = a.a_hist + hs * (int64_t)(Tn > 1 ? Tn - 1 : 0) * N;
    const T* xstar = a.x_star + (int64_t)c * Tn * d;
    const T* ysc = a.ys ? a.ys + (int64_t)c * dsc.ys_chain_stride : nullptr;
    T w[PPT], xn[PPT][D];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      w[q] = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) xn[q][k] = T(0);
    }
    // one-step-ahead register prefetch of the draws and per-step parameters
    T ur_n[PPT], uk_n[PPT], eps_n[PPT][D], xs_n = T(0), y_n = T(0), u_n = T(0), g_n = T(0), s_n = T(1);
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      ur_n[q] = uk_n[q] = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) eps_n[q][k] = T(0);
    }
    auto prefetch = [&](int tt) {
      const bool go = tt < Tn;
      if (go && act) {
        if (tt > 0 && dsc.resampling == AUXSSM_RESAMPLE_SYSTEMATIC) {
          const int64_t us = ((int64_t)c * (Tn - 1) + (tt - 1)) * 3;  // (U, V, W), broadcast loads
          ur_n[0] = a.draw_u_res(us); ur_n[1] = a.draw_u_res(us + 1); ur_n[2] = a.draw_u_res(us + 2);
        } else if (tt > 0 && a.rng.on) {
          philox_run<T, PPT>(PHILOX_UNIFORM, a.rng.seed, a.call0() + 1, (uint64_t)(((int64_t)c * (Tn - 1) + (tt - 1)) * N + p0), PPT, ur_n);
          if (dsc.resampling == AUXSSM_RESAMPLE_KILLING)
            philox_run<T, PPT>(PHILOX_UNIFORM, a.rng.seed, a.call0() + 2, (uint64_t)(((int64_t)c * (Tn - 1) + (tt - 1)) * N + p0), PPT, uk_n);
        } else if (tt > 0) {
          ldv<T, PPT>(ur_n, a.u_res + ((int64_t)c * (Tn - 1) + (tt - 1)) * N + p0);
          if (dsc.resampling == AUXSSM_RESAMPLE_KILLING)
            ldv<T, PPT>(uk_n, a.u_kill + ((int64_t)c * (Tn - 1) + (tt - 1)) * N + p0);
        }
        if (a.rng.on) {
          T run[PPT * D];
          philox_run<T, PPT * D>(PHILOX_NORMAL, a.rng.seed, a.call0(), (uint64_t)((((int64_t)c * Tn + tt) * N + p0) * d), PPT * d, run);
#pragma unroll
          for (int q = 0; q < PPT; ++q)
#pragma unroll
            for (int k = 0; k < D; ++k) if (k < d) eps_n[q][k] = run[(D == 1) ? q : q * d + k];
        } else {
        const T* eg = a.normals + (((int64_t)c * Tn + tt) * N + p0) * d;
        if (D == 1) {
          T e4[PPT];
          ldv<T, PPT>(e4, eg);
#pragma unroll
          for (int q = 0; q < PPT; ++q) eps_n[q][0] = e4[q];
        } else {
#pragma unroll
          for (int q = 0; q < PPT; ++q)
#pragma unroll
            for (int k = 0; k < D; ++k) if (k < d) eps_n[q][k] = eg[q * d + k];
        }
        }
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
      T ur[PPT], uk[PPT], eps[PPT][D];
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        ur[q] = ur_n[q]; uk[q] = uk_n[q];
#pragma unroll
        for (int k = 0; k < D; ++k) eps[q][k] = eps_n[q][k];
      }
      const T xs_c = xs_n, y_c = y_n, u_c = u_n, g_c = g_n, s_c = s_n;
      prefetch(t + 1);
      int anc[PPT];
      T xp[PPT][D];
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        anc[q] = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) xp[q][k] = T(0);
      }
      if (!first && act) {
        // #{j : cdf[j] < r} by a power-of-two descent, the four searches of this thread interleaved
        const T tot = cdf[N - 1];
        T r[PPT];
        int lo[PPT];
        if (dsc.resampling == AUXSSM_RESAMPLE_SYSTEMATIC) {
          const SysPlan<T> pl = systematic_plan<T>(cdf, N, ur[0], ur[1], ur[2]);
#pragma unroll
          for (int q = 0; q < PPT; ++q) { r[q] = pl.r(p0 + q, N); lo[q] = 0; }
        } else {
#pragma unroll
          for (int q = 0; q < PPT; ++q) { r[q] = tot * (T(1) - ur[q]); lo[q] = 0; }
        }
        for (int step = hb; step > 0; step >>= 1) {
#pragma unroll
          for (int q = 0; q < PPT; ++q) {
            const int nx = lo[q] + step;
            if (nx <= N && cdf[nx - 1] < r[q]) lo[q] = nx;
          }
        }
        const T wmax = misc[0];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          int an = lo[q] < N - 1 ? lo[q] : N - 1;
          if (dsc.resampling == AUXSSM_RESAMPLE_KILLING && uk[q] * wmax < w[q]) an = p0 + q;
          if (p0 + q == 0) an = 0;
          anc[q] = an;
#pragma unroll
          for (int k = 0; k < D; ++k) if (k < d) xp[q][k] = part[an * d + k];
        }
      }
      if (i < D) {
        m.xs_t[i] = xs_c;
        if (ysc) m.y[i] = y_c;
        if (PROP != AUXSSM_PROP_BOOTSTRAP) {
          m.u_t[i] = u_c;
          m.mean_t[i] = u_c + s_c * s_c * g_c;
        }
      }
      if (PROP != AUXSSM_PROP_BOOTSTRAP) m.s_t = s_c;
      if (PROP == AUXSSM_PROP_GUIDED && (t <= 1 || dsc.gains_time_stride != 0)) {
        load_padded<T, D>(a.K + (int64_t)t * dsc.gains_time_stride, m.K, d);
        load_padded<T, D>(a.cL + (int64_t)t * dsc.gains_time_stride, m.cL, d);
      }
      __syncthreads();  // parents gathered, parameters visible
      if (PROP == AUXSSM_PROP_GUIDED && (t <= 1 || dsc.gains_time_stride != 0)) {
        m.cst_l = factor_consts<T, D>(m.cL, m.invl, d);
        __syncthreads();
      }
      T lw[PPT];
#pragma unroll
      for (int q = 0; q < PPT; ++q)
        lw[q] = propose_and_weight<T, D, PROP>(dsc, m, first, p0 + q == 0, xp[q], eps[q], xn[q]);
      if (act) {
        T* xg = xs_h + ((int64_t)t * N + p0) * d;
        if (D == 1) {
          T x4[PPT];
#pragma unroll
          for (int q = 0; q < PPT; ++q) x4[q] = xn[q][0];
          stv<T, PPT>(part + p0, x4);
          stv<T, PPT>(xg, x4);
        } else {
#pragma unroll
          for (int q = 0; q < PPT; ++q)
#pragma unroll
            for (int k = 0; k < D; ++k)
              if (k < d) { part[(p0 + q) * d + k] = xn[q][k]; xg[q * d + k] = xn[q][k]; }
        }
        stv<T, PPT>(lw_h + (int64_t)t * N + p0, lw);
        if (!first) stv<int, PPT>(a_h + (int64_t)(t - 1) * N + p0, anc);
      }
      // normalise (math/utils.py:24-39): thread-local partials first
      T amax = block_max<T>(act ? fmax(fmax(lw[0], lw[1]), fmax(lw[2], lw[3])) : -INFINITY, red);
      if (!isfinite(amax)) amax = T(0);
      T e[PPT];
#pragma unroll
      for (int q = 0; q < PPT; ++q) e[q] = act ? exp(lw[q] - amax) : T(0);
      const T s = block_sum<T>((e[0] + e[1]) + (e[2] + e[3]), red + 32);
      const T lse = log(s) + amax;
#pragma unroll
      for (int q = 0; q < PPT; ++q) w[q] = act ? exp(lw[q] - lse) : T(0);
      T c4[PPT];
      c4[0] = w[0]; c4[1] = c4[0] + w[1]; c4[2] = c4[1] + w[2]; c4[3] = c4[2] + w[3];
      T excl;
      block_scan_excl<T>(c4[3], red + 64, excl);
      if (act) {
#pragma unroll
        for (int q = 0; q < PPT; ++q) c4[q] = excl + c4[q];
        stv<T, PPT>(cdf + p0, c4);
      }
      if (dsc.resampling == AUXSSM_RESAMPLE_KILLING) {
        const T wmax = block_max<T>(fmax(fmax(w[0], w[1]), fmax(w[2], w[3])), red + 96);
        if (i == 0) misc[0] = wmax;
      }
      __syncthreads();
    }

    // ------------------------------------ backward pass ----------------------------------------------
    T* xo = a.x_out + (int64_t)c * Tn * d;
    int32_t* ao = a.anc + (int64_t)c * Tn;
    const int64_t ub = (int64_t)c * Tn;
    BwdCache<T> bwdc;
    // B = min(#{j : cdf[j] < r}, N - 1) (csmc.py:111 / :131): per-thread counts, warp add, one shared atomic per
    // warp; the owner of particle B then publishes it as the selected state.  All threads call (three barriers).
    auto select = [&](T r, const T (&xc)[PPT][D], int t) {
      int cnt = 0;
      if (act) {
        T c4[PPT];
        ldv<T, PPT>(c4, cdf + p0);
#pragma unroll
        for (int q = 0; q < PPT; ++q) cnt += (c4[q] < r) ? 1 : 0;
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if ((i & 31) == 0 && cnt > 0) atomicAdd(isel, cnt);
      __syncthreads();
      int B = isel[0];
      if (B > N - 1) B = N - 1;
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        if (p0 + q == B) {
#pragma unroll
          for (int k = 0; k < D; ++k) {
            m.xsel[k] = (k < d) ? xc[q][k] : T(0);
            if (k < d) xo[(int64_t)t * d + k] = xc[q][k];
          }
          ao[t] = B;
        }
      }
      __syncthreads();
      if (i == 0) isel[0] = 0;  // re-armed for the next call (ordered by the barriers of the reductions in between)
      return B;
    };
    int B_last;
    {
      if (i == 0) isel[0] = 0;
      __syncthreads();
      const T r = cdf[N - 1] * (T(1) - a.draw_u_bwd(ub, bwdc));
      B_last = select(r, xn, Tn - 1);
    }
    if (dsc.backward) {
      T ub_n = (Tn > 1) ? a.draw_u_bwd(ub + 1, bwdc) : T(0);
      T xt_n[PPT][D], lwh_n[PPT];
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        lwh_n[q] = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) xt_n[q][k] = T(0);
      }
      auto fetch = [&](int tt) {
        if (tt >= 0 && act) {
          const T* xg = xs_h + ((int64_t)tt * N + p0) * d;
          if (D == 1) {
            T x4[PPT];
            ldv<T, PPT>(x4, xg);
#pragma unroll
            for (int q = 0; q < PPT; ++q) xt_n[q][0] = x4[q];
          } else {
#pragma unroll
            for (int q = 0; q < PPT; ++q)
#pragma unroll
              for (int k = 0; k < D; ++k) if (k < d) xt_n[q][k] = xg[q * d + k];
          }
          ldv<T, PPT>(lwh_n, lw_h + (int64_t)tt * N + p0);
        }
      };
      fetch(Tn - 2);
      for (int t = Tn - 2, k2 = 1; t >= 0; --t, ++k2) {
        T xt[PPT][D], lwh[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          lwh[q] = lwh_n[q];
#pragma unroll
          for (int k = 0; k < D; ++k) xt[q][k] = xt_n[q][k];
        }
        const T ub_c = ub_n;
        if (t > 0) ub_n = a.draw_u_bwd(ub + k2 + 1, bwdc);
        fetch(t - 1);
        T xsel[D];
#pragma unroll
        for (int k = 0; k < D; ++k) xsel[k] = m.xsel[k];
        T lw[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          T xpred[D];
          matvec<T, D>(m.F, xt[q], xpred);
#pragma unroll
          for (int k = 0; k < D; ++k) xpred[k] += m.b[k];
          lw[q] = prior_logpdf<T, D>(m, false, xsel, xpred) + lwh[q];
        }
        T amax = block_max<T>(act ? fmax(fmax(lw[0], lw[1]), fmax(lw[2], lw[3])) : -INFINITY, red);
        if (!isfinite(amax)) amax = T(0);
        T e[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) e[q] = act ? exp(lw[q] - amax) : T(0);
        const T s = block_sum<T>((e[0] + e[1]) + (e[2] + e[3]), red + 32);
        const T lse = log(s) + amax;
        T c4[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) c4[q] = act ? exp(lw[q] - lse) : T(0);
        c4[1] = c4[0] + c4[1]; c4[2] = c4[1] + c4[2]; c4[3] = c4[2] + c4[3];
        T excl;
        block_scan_excl<T>(c4[3], red + 64, excl);
        if (act) {
#pragma unroll
          for (int q = 0; q < PPT; ++q) c4[q] = excl + c4[q];
          stv<T, PPT>(cdf + p0, c4);
        }
        __syncthreads();  // cdf complete; every thread holds the old x_sel in registers
        const T r = cdf[N - 1] * (T(1) - ub_c);
        select(r, xt, t);
      }
    } else {
      if (i == 0) {
        int B = B_last;
        for (int t = Tn - 2; t >= 0; --t) {
          B = a_h[(int64_t)t * N + B];
          ao[t] = B;
          for (int k = 0; k < d; ++k) xo[(int64_t)t * d + k] = xs_h[((int64_t)t * N + B) * d + k];
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace auxssm

// Structured Kalman filter (H = I, diagonal R_t, time-invariant diagonal F, Q, b) with SEVERAL warps per chain.
//
// Same recursion and the same blocked-by-four measurement update as filter_warp_kernel (kalman_filter.cu;
// replaces sequential_predict_update, filtering.py:84-147), for state dimensions whose covariance tile does not
// fit the registers of one warp: NW warps (32 NW threads) hold P, Q in a (4 NW) x 8 thread grid, thread (ty, tx)
// owning a (D / 4NW) x (D / 8) tile.  D = 64 runs with NW = 4 (one chain per 128-thread CTA, tiles of 4 x 8).
// The four rows a block absorbs belong to ONE thread row-group (TM = 4), which publishes them through a
// double-buffered shared row buffer; everybody then factorises the 4 x 4 innovation block redundantly in registers.
// Synchronisation is a named barrier per chain (bar.sync id, 32 NW) instead of __syncwarp.
#pragma once

namespace auxssm {

template <int NTHREADS> __device__ __forceinline__ void chain_sync(int id) {
  if (NTHREADS == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTHREADS) : "memory");
}

template <typename T, int D, int NW>
__global__ void __launch_bounds__(128, 2) filter_tile_kernel(FilterArgs<T> a, int chunk_first) {
  constexpr int NTC = 32 * NW;            // threads per chain
  constexpr int CPC = 128 / NTC;          // chains per CTA
  constexpr int GY = 4 * NW, GX = 8, TM = D / GY, TN = D / GX, LDS_ = D + 4, NB = 4, RS = D + 4;
  constexpr int NWORDS = (D + 31) / 32;   // 32-bit words of the observation mask
  static_assert(TM == NB && TN % 2 == 0 && CPC >= 1, "tile shape: one block of four rows per thread row-group");
  if (a.gate != nullptr && *a.gate != 0) return;  // F is not diagonal after all: the general kernel runs instead
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int cq = threadIdx.x / NTC, tid = threadIdx.x % NTC;  // chain slot in the CTA, thread in the chain
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)cq * (D * LDS_ + 4 * D + 32);
  T* stage = base;            // [D][D+4] output stage; the row-exchange buffers alias its head (disjoint in time)
  T* rowbuf = base;
  T* ev = base + D * LDS_;
  T* rdv = ev + D;
  T* svbuf = rdv + D;
  T* mvec = svbuf + D;
  T* red = mvec + D;          // NW partial sums
  unsigned* maskw = reinterpret_cast<unsigned*>(red + 8);  // NWORDS mask words
  const int ty = tid / GX, tx = tid % GX, lane = tid & 31, wq = tid >> 5, bar_id = 1 + cq;
  const int j = chunk_first + blockIdx.y, unit = blockIdx.x * CPC + cq;
  if (unit >= a.d.n_chains * a.d.n_batch) return;  // whole chain group; only the chain barrier is used below
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int d = a.d.dx, Tn = a.d.T;
  int t0, t1;
  chunk_range(j, a.chunk_len, Tn, t0, t1);

  const T* Fg = a.Fs ? a.Fs + unit_off(a.d.Fs, c, b) : nullptr;
  const T* Qg = a.Qs ? a.Qs + unit_off(a.d.Qs, c, b) : nullptr;
  const T* bg = a.bs ? a.bs + unit_off(a.d.bs, c, b) : nullptr;
  T p[TM][TN], q[TM][TN], mreg[TM], fa[TM], fc[TN], ba[TM];
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
      q[r][cc] = (Qg && ra < d && cb < d) ? Qg[ra * d + cb] : T(0);
    }
  }
  const int psz = d + d * d;
  const T* m_src = (j == 0) ? a.m0 + unit_off(a.d.m0, c, b) : a.prefix + ((int64_t)unit * a.n_chunks + (j - 1)) * psz;
  const T* P_src = (j == 0) ? a.P0 + unit_off(a.d.P0, c, b) : m_src + d;
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int ra = ty * TM + r;
    mreg[r] = (tx == 0 && ra < d) ? m_src[ra] : T(0);
#pragma unroll
    for (int cc = 0; cc < TN; ++cc) {
      const int cb = tx * TN + cc;
      p[r][cc] = (ra < d && cb < d) ? P_src[ra * d + cb] : T(0);
    }
  }
  const T* yg = a.ys + unit_off(a.d.ys, c, b);
  const T* rg = a.Rs + unit_off(a.d.Rs, c, b);
  const T* cg = a.cs ? a.cs + unit_off(a.d.cs, c, b) : nullptr;
  const bool want_out = a.Ps != nullptr;  // ell-only call otherwise (see filter_warp_kernel)
  T* msg = want_out ? a.ms + unit_off(a.d.ms, c, b) : nullptr;
  T* Psg = want_out ? a.Ps + unit_off(a.d.Ps, c, b) : nullptr;
  const bool pairs_ok = (d & 1) == 0 && (a.d.Ps.time & 1) == 0 &&
                        (reinterpret_cast<uintptr_t>(Psg) % (2 * sizeof(T))) == 0;
  // thread tid < d carries observation entry tid, prefetched one step ahead
  T y_n = T(0), r_n = T(1), c_n = T(0);
  if (tid < d) {
    y_n = yg[(int64_t)t0 * a.d.ys.time + tid];
    r_n = rg[(int64_t)t0 * a.d.Rs.time + tid];
    if (cg) c_n = cg[(int64_t)t0 * a.d.cs.time + tid];
  }
  T ell = T(0);
  int parity = 0;
  for (int t = t0; t < t1; ++t) {
    const T y_c = y_n, r_c = r_n, c_c = c_n;
    if (t + 1 < t1 && tid < d) {
      y_n = yg[(int64_t)(t + 1) * a.d.ys.time + tid];
      r_n = rg[(int64_t)(t + 1) * a.d.Rs.time + tid];
      if (cg) c_n = cg[(int64_t)(t + 1) * a.d.cs.time + tid];
    }
    const bool obs = (tid < d) && finite_(y_c);
    const unsigned wmask = __ballot_sync(0xffffffffu, obs);
    if (lane == 0 && wq < NWORDS) maskw[wq] = wmask;
    if (tid < D) { ev[tid] = y_c - c_c; rdv[tid] = r_c; }
    if (t > 0) {
#pragma unroll
      for (int r = 0; r < TM; ++r) {
#pragma unroll
        for (int c2 = 0; c2 < TN / 2; ++c2) {
          const Pair<T> ff = pair_mul_s(fa[r], pair_make(fc[2 * c2], fc[2 * c2 + 1]));
          const Pair<T> pp = pair_fma(ff, pair_make(p[r][2 * c2], p[r][2 * c2 + 1]),
                                      pair_make(q[r][2 * c2], q[r][2 * c2 + 1]));
          pair_get(pp, p[r][2 * c2], p[r][2 * c2 + 1]);
        }
        mreg[r] = fma_(fa[r], mreg[r], ba[r]);
      }
    }
    chain_sync<NTC>(bar_id);
    unsigned mw[NWORDS];
    bool any_obs = false;
#pragma unroll
    for (int k = 0; k < NWORDS; ++k) { mw[k] = maskw[k]; any_obs = any_obs || (mw[k] != 0u); }
    T quad = T(0);
    // ---- measurement update: one block of four scalar observations per thread row-group ----
#pragma unroll 1
    for (int grp = 0; grp < GY; ++grp) {
      const int i0 = grp * NB;
      unsigned word = mw[0];
#pragma unroll
      for (int k = 1; k < NWORDS; ++k) if ((i0 >> 5) == k) word = mw[k];
      const unsigned bits = (word >> (i0 & 31)) & 0xFu;
      if (bits != 0u) {  // uniform over the chain
        T* rb = rowbuf + parity * (NB * RS);
        parity ^= 1;
        if (ty == grp) {
#pragma unroll
          for (int aa = 0; aa < NB; ++aa) {
            stv<T, TN>(rb + aa * RS + tx * TN, p[aa]);
            if (tx == 0) rb[aa * RS + D] = mreg[aa];
          }
        }
        chain_sync<NTC>(bar_id);
        T sr[NB][NB], rdl[NB], evl[NB], RA[NB][TM], RC[NB][TN], v[NB];
        ldv<T, NB>(rdl, rdv + i0);
        ldv<T, NB>(evl, ev + i0);
#pragma unroll
        for (int aa = 0; aa < NB; ++aa) {
          ldv<T, NB>(sr[aa], rb + aa * RS + i0);
          ldv<T, TM>(RA[aa], rb + aa * RS + ty * TM);
          ldv<T, TN>(RC[aa], rb + aa * RS + tx * TN);
          v[aa] = evl[aa] - rb[aa * RS + D];
        }
        T L[NB][NB], inv[NB];
        if (bits == 0xFu) {
#pragma unroll
          for (int aa = 0; aa < NB; ++aa)
#pragma unroll
            for (int bb = 0; bb <= aa; ++bb) L[aa][bb] = (aa == bb) ? sr[aa][aa] + rdl[aa] : sr[aa][bb];
        } else {  // masked block: a missing entry is a unit pivot with a zero row
#pragma unroll
          for (int aa = 0; aa < NB; ++aa) {
            const T oka = ((bits >> aa) & 1u) ? T(1) : T(0);
#pragma unroll
            for (int bb = 0; bb <= aa; ++bb) {
              const T okb = ((bits >> bb) & 1u) ? T(1) : T(0);
              L[aa][bb] = (aa == bb) ? (oka != T(0) ? sr[aa][aa] + rdl[aa] : T(1)) : oka * okb * sr[aa][bb];
            }
#pragma unroll
            for (int r = 0; r < TM; ++r) RA[aa][r] *= oka;
#pragma unroll
            for (int cc = 0; cc < TN; ++cc) RC[aa][cc] *= oka;
            v[aa] *= oka;
          }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          T piv = L[k][k];
#pragma unroll
          for (int jj = 0; jj < k; ++jj) piv = fma_(-L[k][jj], L[k][jj], piv);
          if (tid == 0) svbuf[i0 + k] = piv;
          T lkk;
          pivot_fast_<T>(piv, lkk, inv[k]);
          L[k][k] = lkk;
#pragma unroll
          for (int r = k + 1; r < NB; ++r) {
            T acc = L[r][k];
#pragma unroll
            for (int jj = 0; jj < k; ++jj) acc = fma_(-L[r][jj], L[k][jj], acc);
            L[r][k] = acc * inv[k];
          }
        }
        const T n10 = -L[1][0], n20 = -L[2][0], n21 = -L[2][1], n30 = -L[3][0], n31 = -L[3][1], n32 = -L[3][2];
        auto fsub = [&](T& x0, T& x1, T& x2, T& x3) {
          x0 = x0 * inv[0];
          x1 = fma_(n10, x0, x1) * inv[1];
          x2 = fma_(n21, x1, fma_(n20, x0, x2)) * inv[2];
          x3 = fma_(n32, x2, fma_(n31, x1, fma_(n30, x0, x3))) * inv[3];
        };
        auto fsub2 = [&](T& a0, T& b0, T& a1, T& b1, T& a2, T& b2, T& a3, T& b3) {
          Pair<T> x0 = pair_mul_s(inv[0], pair_make(a0, b0));
          Pair<T> x1 = pair_mul_s(inv[1], pair_fma_s(n10, x0, pair_make(a1, b1)));
          Pair<T> x2 = pair_mul_s(inv[2], pair_fma_s(n21, x1, pair_fma_s(n20, x0, pair_make(a2, b2))));
          Pair<T> x3 = pair_mul_s(inv[3], pair_fma_s(n32, x2, pair_fma_s(n31, x1, pair_fma_s(n30, x0, pair_make(a3, b3)))));
          pair_get(x0, a0, b0); pair_get(x1, a1, b1); pair_get(x2, a2, b2); pair_get(x3, a3, b3);
        };
        // (Solving Z once and exchanging its column entries through shared memory, as filter_warp_kernel does, was
        // measured here too: 41.3 instead of 39.2 ms per call -- the second 128-thread barrier per block costs more than
        // the saved substitutions.)
#pragma unroll
        for (int r = 0; r < TM; r += 2)
          fsub2(RA[0][r], RA[0][r + 1], RA[1][r], RA[1][r + 1], RA[2][r], RA[2][r + 1], RA[3][r], RA[3][r + 1]);
#pragma unroll
        for (int cc = 0; cc < TN; cc += 2)
          fsub2(RC[0][cc], RC[0][cc + 1], RC[1][cc], RC[1][cc + 1], RC[2][cc], RC[2][cc + 1], RC[3][cc], RC[3][cc + 1]);
        fsub(v[0], v[1], v[2], v[3]);
#pragma unroll
        for (int aa = 0; aa < NB; ++aa) {
          Pair<T> nrc[TN / 2];
#pragma unroll
          for (int c2 = 0; c2 < TN / 2; ++c2) nrc[c2] = pair_make(-RC[aa][2 * c2], -RC[aa][2 * c2 + 1]);
#pragma unroll
          for (int r = 0; r < TM; ++r) {
#pragma unroll
            for (int c2 = 0; c2 < TN / 2; ++c2) {
              const Pair<T> pp = pair_fma_s(RA[aa][r], nrc[c2], pair_make(p[r][2 * c2], p[r][2 * c2 + 1]));
              pair_get(pp, p[r][2 * c2], p[r][2 * c2 + 1]);
            }
            mreg[r] = fma_(RA[aa][r], v[aa], mreg[r]);  // only the tx == 0 copies are ever read
          }
          quad = fma_(v[aa], v[aa], quad);
        }
      }
    }
    // ---- log-likelihood increment of this step (nan_to_num -> 0, filtering.py:125) ----
    chain_sync<NTC>(bar_id);  // svbuf complete
    unsigned myw = mw[0];
#pragma unroll
    for (int k = 1; k < NWORDS; ++k) if ((tid >> 5) == k) myw = mw[k];
    T part = ((tid < d) && ((myw >> (tid & 31)) & 1u)) ? T(-0.5) * log(svbuf[tid]) - half_log_2pi<T>() : T(0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (NW > 1) {
      if (lane == 0) red[wq] = part;
    }
    // ---- outputs: tile -> shared -> coalesced rows ----
    const bool stage_now = want_out || (t + 1 == t1 && j == 0 && j + 1 < a.n_chunks);
    if (stage_now) {
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        stv<T, TN>(stage + (ty * TM + r) * LDS_ + tx * TN, p[r]);
        if (tx == 0) mvec[ty * TM + r] = mreg[r];
      }
    }
    chain_sync<NTC>(bar_id);
    if (NW > 1) {
      part = T(0);
#pragma unroll
      for (int k = 0; k < NW; ++k) part += red[k];
    }
    T inc = part - T(0.5) * quad;
    if (!(inc == inc)) inc = T(0);
    if (any_obs) ell += inc;
    if (!want_out) { chain_sync<NTC>(bar_id); continue; }
    T* Pt = Psg + (int64_t)t * a.d.Ps.time;
    if (pairs_ok) {
      constexpr int LPR = D / 4, RPP = NTC / LPR;
      const int lr = tid / LPR, c4 = (tid % LPR) * 4;
#pragma unroll 4
      for (int r0 = 0; r0 < d; r0 += RPP) {
        const int r = r0 + lr;
        if (r < d) {
          T v4[4];
          ldv<T, 4>(v4, stage + r * LDS_ + c4);
          const T va[2] = {v4[0], v4[1]}, vb[2] = {v4[2], v4[3]};
          if (c4 < d) stv<T, 2>(Pt + r * d + c4, va);
          if (c4 + 2 < d) stv<T, 2>(Pt + r * d + c4 + 2, vb);
        }
      }
    } else {
      for (int e = tid; e < d * d; e += NTC) {
        const int r = e / d, cc = e - r * d;
        Pt[e] = stage[r * LDS_ + cc];
      }
    }
    if (tid < d) msg[(int64_t)t * a.d.ms.time + tid] = mvec[tid];
    chain_sync<NTC>(bar_id);
  }
  if (j == 0 && j + 1 < a.n_chunks) {
    T* pf = a.prefix + ((int64_t)unit * a.n_chunks + j) * psz;
    if (tid < d) pf[tid] = mvec[tid];
    for (int e = tid; e < d * d; e += NTC) {
      const int r = e / d, cc = e - r * d;
      pf[d + e] = stage[r * LDS_ + cc];
    }
  }
  if (tid == 0) a.ellpart[(int64_t)unit * a.n_chunks + j] = ell;
}

template <int D, int NW> constexpr size_t filter_tile_smem_elems() {
  return (size_t)(128 / (32 * NW)) * (D * (D + 4) + 4 * D + 32);
}

}  // namespace auxssm

// Backward (pathwise smoothing) sampling for linear-Gaussian state-space models.
//
// Replaces aux_samplers/_primitives/kalman/sampling.py of the reference:
//   sampling:11-40, _sampling_op_impl:52-55, mean_and_chol:61-105, _sampling_init_one:109-112,
//   _sample_last_step:116-124, _sampling_init:127-136 (the N(0, I) draws of :128 are an input).
//
// Per time step t < T-1, with (m, P) the filtering marginal at t and (F, Q, b) the transition to t+1:
//   S = F P F^T + Q = Ls Ls^T,  Zs = Ls^{-1} F P,  gain G = P F^T S^{-1} = Zs^T Ls^{-1},
//   Sigma = P - G S G^T = P - Zs^T Zs = L L^T (NaN -> 0),  e = m - G (F m + b) + L eps,
//   x_t = G x_{t+1} + e,      x_{T-1} = m + chol(P) eps.
//
// The expensive part -- the two factorisations and the gain -- depends on (m_t, P_t) only, NOT on
// x_{t+1}: exactly the observation behind the reference's parallel formulation (sampling.py:28-32,
// elements (G_t, e_t) built by a vmap, then an associative scan).  So:
//   pass 1  `sample_elements_kernel`: every (chain, time step) is an independent work item -- one CTA per
//           (unit, tile of 16 steps), throughput-bound, all SMs busy whatever the number of chains;
//           it stores (G_t^T, e_t) to scratch;
//   pass 2  `sample_replay_kernel`: the affine recursion x_t = G_t x_{t+1} + e_t per chain.  Composing maps
//           (the scan operator G2 G1) costs d^3 per element while applying one to a vector costs d^2, so
//           the recursion is replayed sequentially (d^2 FMAs and one 4 d^2-byte read per step).
// Both `parallel` settings run this schedule (they are the same function of the inputs); chains are
// processed in waves so that the scratch stays below kScratchCap.
#include "common.cuh"
#include "smem_linalg.cuh"

namespace auxssm {

constexpr size_t kScratchCap = (size_t)6 << 30;  // bytes of (G_t, e_t) scratch per wave of chains
constexpr int kTileLen = 16;                     // time steps per CTA in pass 1

template <typename T> struct SampleArgs {
  auxssm_lgssm_desc d;
  const T *ms, *Ps, *Fs, *Qs, *bs, *eps;
  T* xs;
  T* steps;  // [wave units][T][dx*dx + dx]   (G_t^T, e_t); entry T-1 holds e = x_{T-1} only
  int* gate; // diagonal-F gate (common.cuh): the LEAN kernel runs when *gate == 0, the general one otherwise
  int unit0, n_units;
};

template <typename T, int D> struct SampBuf {
  T *P, *W, *S, *Z, *Ft, *Q;
  T *m, *bv, *eps, *r, *e, *dg;
};

// lean = time-invariant diagonal F promised by the caller: only P, S, Z live in shared memory (Q is read
// straight from global / L2 once per element, F is a vector), which doubles the resident work items per SM.
template <int D> __host__ __device__ constexpr size_t samp_smem_elems(bool lean) {
  return (size_t)(lean ? 3 : 6) * D * Tile<D>::LD + 6 * D + CholWs<D>::WORDS + 16;
}

// In-place lower Cholesky with the reference's nan_to_num semantics: jnp.linalg.cholesky returns an
// all-NaN factor when it fails, which nan_to_num turns into the zero matrix (sampling.py:102-104).
template <typename T, int D, bool PADDED = false>
__device__ __forceinline__ void chol_nan_to_zero(T* A, T* dg, int n) {
  constexpr int LD = Tile<D>::LD;
  chol_inplace<T, D, PADDED>(A, dg, n);
  bool bad = false;
  if constexpr (Tile<D>::NT == 32) {
    const int l = threadIdx.x;
    const bool mine = (l < n) && (!(dg[l < n ? l : 0] > T(0)) || !finite_(dg[l < n ? l : 0]));
    bad = __any_sync(0xffffffffu, mine) != 0;
  } else {
    for (int i = 0; i < n; ++i) bad = bad || !(dg[i] > T(0)) || !finite_(dg[i]);
  }
  if (n == 1) bad = !finite_(dg[0]);  // scalar branch of the reference is a plain sqrt
  if (bad) {
    AUXSSM_FOR_ROWS(i, j, n, n) A[i * LD + j] = T(0);
  }
  cta_sync<D>();
}

// PAD path (one warp, D <= 32): y_lane = (chol(A) x)_lane with the reference's NaN -> 0 semantics, the factor itself is
// never stored (chol_reg FUSE_MV).  A failed factorisation shows as a non-positive / non-finite pivot in dg.
template <typename T, int D> __device__ __forceinline__ T chol_times_vec_pad(T* A, T* dg, const T* x, int n) {
  T y = T(0);
  chol_reg<T, D, true, false, true>(A, dg, nullptr, dg + 2 * D, n, x, &y);
  const int l = threadIdx.x;
  const bool mine = (l < n) && (!(dg[l < n ? l : 0] > T(0)) || !finite_(dg[l < n ? l : 0]));
  const bool bad = (n == 1) ? !finite_(dg[0]) : (__any_sync(0xffffffffu, mine) != 0);
  return (bad || l >= n) ? T(0) : y;
}

// One warp, fp32, D = 32:  P <- P - Z^T Z  on the tensor cores with the 3xTF32 split (hi*hi + hi*lo + lo*hi, fp32
// accumulate: error 5e-7 of max|P|, the same as the FFMA form; single-pass TF32 would be 2e-4 -- profiles/
// r2_tf32x3_probe.jsonl).  Why here: the elements kernel is bound by SHARED-MEMORY bandwidth (ncu: 0.85 wavefronts per
// clock per SM); the register-tiled FFMA2 product reads 12 words per lane per k (3 LDS.128 = 12 wavefronts), the MMA
// fragments 8 words per lane per EIGHT k -- Z^T Z is symmetric in its operand, so the A fragment of m-tile i is made of
// the B-fragment values of n-tiles 2i and 2i + 1 -- and only the 6 of 8 (16 x 8) output tiles that touch the lower
// triangle are formed (the result feeds a Cholesky factorisation).
__device__ __forceinline__ unsigned to_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// One warp's share: the two 16-row m-tiles M0, M1 against all D / 8 n-tiles (those touching the lower triangle only).
template <int D, int LD, int M0, int M1> __device__ __forceinline__ void ztz_sub_tiles(const float* Z, float* P) {
  constexpr int NTI = D / 8;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float c[2][NTI][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NTI; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[i][j][q] = 0.f;
  constexpr int MT[2] = {M0, M1};
  constexpr int JMAX = (16 * (M0 > M1 ? M0 : M1) + 15) / 8;  // last n-tile any of the two m-tiles needs
#pragma unroll
  for (int k0 = 0; k0 < D; k0 += 8) {
    unsigned hi[NTI][2], lo[NTI][2];
#pragma unroll
    for (int j = 0; j < NTI; ++j) {
      if (j <= JMAX) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float v = Z[(k0 + t + 4 * q) * LD + 8 * j + g];
          hi[j][q] = to_tf32(v);
          lo[j][q] = to_tf32(v - __uint_as_float(hi[j][q]));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      // a0:(m = 16 M + g, k = t)  a1:(m + 8, k = t)  a2:(m, k = t + 4)  a3:(m + 8, k = t + 4): the B values of n-tiles 2M, 2M+1
      const unsigned ah[4] = {hi[2 * MT[i]][0], hi[2 * MT[i] + 1][0], hi[2 * MT[i]][1], hi[2 * MT[i] + 1][1]};
      const unsigned al[4] = {lo[2 * MT[i]][0], lo[2 * MT[i] + 1][0], lo[2 * MT[i]][1], lo[2 * MT[i] + 1][1]};
#pragma unroll
      for (int j = 0; j < NTI; ++j) {
        if (8 * j <= 16 * MT[i] + 15) {  // the tile touches the lower triangle
          mma_tf32_16x8x8(c[i][j], al, hi[j]);
          mma_tf32_16x8x8(c[i][j], ah, lo[j]);
          mma_tf32_16x8x8(c[i][j], ah, hi[j]);
        }
      }
    }
  }
  // c0:(row g, col 2t) c1:(row g, col 2t + 1) c2:(row g + 8, col 2t) c3:(row g + 8, col 2t + 1)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NTI; ++j)
      if (8 * j <= 16 * MT[i] + 15) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float2* pp = reinterpret_cast<float2*>(P + (16 * MT[i] + g + 8 * h) * LD + 8 * j + 2 * t);
          float2 v = *pp;
          v.x -= c[i][j][2 * h];
          v.y -= c[i][j][2 * h + 1];
          *pp = v;
        }
      }
}
// D = 32: one warp, m-tiles {0, 1}.  D = 64: two warps; the lower triangle gives m-tile M (16 rows) 2M + 2 n-tiles of
// work, so the pairs {0, 3} and {1, 2} balance the two warps (10 tiles each).
template <int D, int LD> __device__ __forceinline__ void ztz_sub_3xtf32(const float* Z, float* P) {
  cta_sync<D>();
  if constexpr (D == 32) {
    ztz_sub_tiles<32, LD, 0, 1>(Z, P);
  } else {
    static_assert(D == 64, "tensor-core product: D = 32 or 64");
    if ((threadIdx.x >> 5) == 0) ztz_sub_tiles<64, LD, 0, 3>(Z, P); else ztz_sub_tiles<64, LD, 1, 2>(Z, P);
  }
  cta_sync<D>();
}

// D = 64 lean path (two warps, ONE item per CTA): y_r = (chol(A) x)_r with NaN -> 0, the factor never stored.
template <typename T> __device__ __forceinline__ T chol_times_vec64(T* A, T* dg, const T* x, int n) {
  T y = T(0);
  chol_rows64<T, sizeof(T) == 4, true>(A, dg, nullptr, dg + 2 * 64, n, x, &y);
  const int l = threadIdx.x;
  const bool mine = (l < n) && (!(dg[l < n ? l : 0] > T(0)) || !finite_(dg[l < n ? l : 0]));
  const bool bad = __syncthreads_or(mine ? 1 : 0) != 0;  // W = 1: the item's 64 threads are the whole CTA
  return (bad || l >= n) ? T(0) : y;
}

// global [dx x dx] <- shared [dx x LD] for a multi-warp item: 16-byte chunks (dx % 4 == 0, 16-byte aligned g)
template <typename T, int D> __device__ __forceinline__ void store_chunks(T* __restrict__ g, const T* s, int dx) {
  constexpr int LD = Tile<D>::LD, NT = Tile<D>::NT;
  const int cpr = dx / 4;
#pragma unroll 4
  for (int ch = threadIdx.x; ch < dx * cpr; ch += NT) {
    const int r = ch / cpr, c4 = (ch - r * cpr) * 4;
    T v[4];
    ldv<T, 4>(v, s + r * LD + c4);
    stv<T, 4>(g + r * dx + c4, v);
  }
}

// B <- L^{-T} B for the n x m block (thread j owns column j), right-looking: for i = n-1..0
//   z_i = b_i / L_ii ;  b_k -= L_ik z_i  (k < i)      -- uses ROW i of L, i.e. broadcast vector loads.
// dg = the workspace filled by the factorisation.  Register-resident for D <= 32.
template <typename T, int D, bool PADDED = false>
__device__ __forceinline__ void bwd_solve(const T* L, T* B, const T* dg, int n, int m) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  const int j = threadIdx.x;
  if (j < m) {
    if constexpr (D <= 64) {
      constexpr int V = (D >= 4) ? 4 : D;
      const T* dginv = dg + D;
      T z[D];
#pragma unroll
      for (int i = 0; i < D; ++i) z[i] = (PADDED || i < n) ? B[i * LD + j] : T(0);
#pragma unroll
      for (int i = D - 1; i >= 0; --i) {
        if (PADDED || i < n) {
          const T zi = z[i] * dginv[i];
          z[i] = zi;
          B[i * LD + j] = zi;
          const T nzi = -zi;
#pragma unroll
          for (int k = 0; k < ((i + V - 1) / V) * V; k += V) {
            T l[V];
            ldv<T, V>(l, L + i * LD + k);
#pragma unroll
            for (int q = 0; q < V; q += 2) {
              if (q + 1 < V && k + q + 1 < i) {  // packed pair (two FMAs per issue slot, see smem_linalg.cuh)
                const Pair<T> r2 = pair_fma_s(nzi, pair_make(l[q], l[q + 1]), pair_make(z[k + q], z[k + q + 1]));
                pair_get(r2, z[k + q], z[k + q + 1]);
              } else if (k + q < i) {
                z[k + q] = fma_(nzi, l[q], z[k + q]);
              }
            }
          }
        }
      }
    } else {
      for (int i = n - 1; i >= 0; --i) {
        T acc = B[i * LD + j];
        for (int k = i + 1; k < n; ++k) acc = fma_(-L[k * LD + i], B[k * LD + j], acc);
        B[i * LD + j] = acc / L[i * LD + i];
      }
    }
  }
  cta_sync<D>();
}

// One-warp vectorised passes of the lean path (even dx): a lane owns four consecutive columns of a row, 32 / (D / 4)
// rows per pass; shared accesses are 16-byte vectors, the compact global rows (dx elements, 2-element aligned) move
// as 2-element vectors.
// S_ab = f_a f_b P_ab + Q_ab,  Z_ab = f_a P_ab  (time-invariant diagonal F; fd is zero beyond dx)
template <typename T, int D>
__device__ __forceinline__ void lean_predict_rows(const T* P, const T* fd, const T* Qs, T* S, T* Z, int dx) {
  constexpr int LD = Tile<D>::LD, LPR = D / 4, RPP = 32 / LPR;
  const int lane = threadIdx.x & 31, lr = lane / LPR, c4 = (lane % LPR) * 4;
  T fc[4];
  ldv<T, 4>(fc, fd + c4);
#pragma unroll 4
  for (int r0 = 0; r0 < dx; r0 += RPP) {
    const int r = r0 + lr;
    if (r < dx) {
      T pv[4], q4[4], s4[4], z4[4];
      ldv<T, 4>(pv, P + r * LD + c4);
      ldv<T, 4>(q4, Qs + r * LD + c4);  // the CTA's shared, zero-padded image of the time-invariant Q
      const T fa = fd[r];
      s4[0] = fma_(fa * fc[0], pv[0], q4[0]); s4[1] = fma_(fa * fc[1], pv[1], q4[1]);
      s4[2] = fma_(fa * fc[2], pv[2], q4[2]); s4[3] = fma_(fa * fc[3], pv[3], q4[3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) z4[q] = fa * pv[q];
      stv<T, 4>(S + r * LD + c4, s4);
      stv<T, 4>(Z + r * LD + c4, z4);
    }
  }
}
// The same pass for a multi-warp item (D = 64: 64 threads): a thread owns 4-column chunks, chunk id = tid + NT * pass;
// 16-byte shared accesses and 16-byte global loads of the time-invariant Q (dx % 4 == 0, 16-byte aligned base).  The
// generic row loop it replaces was 18 % of the d = 64 kernel's time (address arithmetic + scalar L2 loads).
template <typename T, int D>
__device__ __forceinline__ void lean_predict_chunks(const T* P, const T* fd, const T* __restrict__ Qg, T* S, T* Z, int dx) {
  constexpr int LD = Tile<D>::LD, NT = Tile<D>::NT, CPR = D / 4;  // chunks per row
  const int cpr = dx / 4;
#pragma unroll 4
  for (int ch = threadIdx.x; ch < dx * CPR; ch += NT) {
    const int r = ch / CPR, cq = ch % CPR, c4 = cq * 4;
    if (cq < cpr) {
      T pv[4], q4[4], fc[4], s4[4], z4[4];
      ldv<T, 4>(pv, P + r * LD + c4);
      ldv<T, 4>(q4, Qg + r * dx + c4);
      ldv<T, 4>(fc, fd + c4);
      const T fa = fd[r];
#pragma unroll
      for (int q = 0; q < 4; ++q) { s4[q] = fma_(fa * fc[q], pv[q], q4[q]); z4[q] = fa * pv[q]; }
      stv<T, 4>(S + r * LD + c4, s4);
      stv<T, 4>(Z + r * LD + c4, z4);
    }
  }
}
// global [dx x dx] <- shared [dx x LD]
template <typename T, int D> __device__ __forceinline__ void store_rows_pairs(T* __restrict__ g, const T* s, int dx) {
  constexpr int LD = Tile<D>::LD, LPR = D / 4, RPP = 32 / LPR;
  const int lane = threadIdx.x & 31, lr = lane / LPR, c4 = (lane % LPR) * 4;
#pragma unroll 4
  for (int r0 = 0; r0 < dx; r0 += RPP) {
    const int r = r0 + lr;
    if (r < dx) {
      T v[4];
      ldv<T, 4>(v, s + r * LD + c4);
      const T va[2] = {v[0], v[1]}, vb[2] = {v[2], v[3]};
      if (c4 < dx) stv<T, 2>(g + r * dx + c4, va);
      if (c4 + 2 < dx) stv<T, 2>(g + r * dx + c4 + 2, vb);
    }
  }
}

// Work items (one warp each) sharing a CTA in the lean one-warp-per-item configuration.  The straight-line body
// of an item (two unrolled Choleskys, forward / backward substitution) is ~100 KB of SASS, three times the 32 KB
// instruction cache: warps that drift apart each stream it from L2 (ncu: `no_instruction` was the largest stall).
// Seven warps per CTA, re-aligned by one block barrier per item, fetch every line once for all seven.
template <typename T, int D, bool LEAN> struct ElemCfg {
  static constexpr int W = (LEAN && D <= 32) ? 7 : 1;  // (four D = 64 items per CTA were measured: no gain)
  static constexpr int MINB = (Occ<T, D>::MINB / W) > 0 ? Occ<T, D>::MINB / W : 1;
  // (a 144-register cap through __maxnreg__ was measured: 15.3 ms per call instead of 11.4 -- ptxas spills the same
  // loop-invariant addresses either way and schedules worse; the 128 registers of __launch_bounds__(224, 2) stay)
};

// pass 1: (G_t^T, e_t) for one unit and a tile of time steps
template <typename T, int D, bool LEAN>
__global__ void __launch_bounds__(Tile<D>::NT * ElemCfg<T, D, LEAN>::W, ElemCfg<T, D, LEAN>::MINB)
sample_elements_kernel(SampleArgs<T> a, int n_tiles) {
  if (a.gate != nullptr && ((*a.gate != 0) == LEAN)) return;  // the other instantiation does this launch's work
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = Tile<D>::NT, LD = Tile<D>::LD, W = ElemCfg<T, D, LEAN>::W;
  T* p = reinterpret_cast<T*>(smem_raw) + (size_t)threadIdx.y * samp_smem_elems<D>(LEAN);
  for (size_t e = threadIdx.x; e < samp_smem_elems<D>(LEAN); e += NT) p[e] = T(0);
  __syncthreads();
  SampBuf<T, D> s;
  s.P = p; p += D * LD; s.S = p; p += D * LD; s.Z = p; p += D * LD;
  s.W = s.Ft = s.Q = nullptr;
  if (!LEAN) { s.W = p; p += D * LD; s.Ft = p; p += D * LD; s.Q = p; p += D * LD; }
  s.m = p; p += D; s.bv = p; p += D; s.eps = p; p += D; s.r = p; p += D; s.e = p; p += D;
  T* fd = p; p += D;  // diag(F) (LEAN)
  s.dg = p;
  constexpr bool PAD = LEAN && (D <= 32);
  // PAD path: one zero-padded image of the time-invariant Q per CTA, shared by its W work items (behind their private
  // regions).  Reading Q_ab from global memory for every item was the kernel's long-scoreboard stall (1.35 per issue).
  T* sQ = reinterpret_cast<T*>(smem_raw) + (size_t)W * samp_smem_elems<D>(LEAN);
  const bool q_shared = PAD && a.d.Qs.chain == 0 && a.d.Qs.batch == 0 && a.Qs != nullptr;
  if (q_shared) {
    const T* Qg0 = a.Qs + unit_off(a.d.Qs, 0, 0);  // chain- and batch-shared (checked by the launcher)
    const int tid = threadIdx.y * NT + threadIdx.x, nthr = NT * W;
    for (int e = tid; e < D * LD; e += nthr) {
      const int r = e / LD, cc = e - r * LD;
      sQ[e] = (r < a.d.dx && cc < a.d.dx) ? Qg0[r * a.d.dx + cc] : T(0);
    }
    __syncthreads();
  }
  // D = 64: P_{t+1} is staged by the TMA unit (one cp.async.bulk per 256-byte row, completed on an mbarrier) into the
  // S buffer as soon as chol(S) is dead, and the two buffers swap roles -- no DRAM round trip on the item's critical
  // path (ncu before: long-scoreboard 2.8 stalls per issue).  Needs 16-byte rows: dx % 4 == 0 and aligned bases.
  constexpr bool BULK = LEAN && (D == 64);
  void* mbar = reinterpret_cast<void*>(s.dg + CholWs<D>::WORDS);  // the 16 spare words behind the workspace
  const bool bulk_ok = BULK && (a.d.dx % 4 == 0) && (a.d.Ps.time % 4 == 0) &&
                       (reinterpret_cast<uintptr_t>(a.Ps + unit_off(a.d.Ps, 0, 0)) % 16 == 0) &&
                       (a.d.Ps.chain % 4 == 0) && (a.d.Ps.batch % 4 == 0);
  unsigned mphase = 0;
  if (BULK) {
    if (threadIdx.x == 0 && bulk_ok) mbar_init(mbar, 1);
    __syncthreads();
  }
  if (PAD) {  // identity on the padded diagonal of S and P: factorisations / solves run guard-free over D
    if (threadIdx.x >= a.d.dx && threadIdx.x < D) {
      s.S[threadIdx.x * LD + threadIdx.x] = T(1);
      s.P[threadIdx.x * LD + threadIdx.x] = T(1);
    }
    __syncthreads();
  }
  // item = (unit, time tile), flattened so that only the last CTA of the grid can be partly filled
  const int64_t item = (int64_t)blockIdx.x * W + threadIdx.y;
  const bool live = item < (int64_t)a.n_units * n_tiles;
  const int ul = live ? (int)(item / n_tiles) : 0, unit = a.unit0 + ul;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx, Tn = a.d.T, ssz = dx * dx + dx;
  const int t_lo = live ? (int)(item % n_tiles) * kTileLen : Tn;
  const int t_hi = (t_lo + kTileLen < Tn) ? t_lo + kTileLen : Tn;
  const int i = threadIdx.x;
  bool first_dyn = true;
  int fdiag = 0;  // 0 unknown, 1 diagonal time-invariant F, -1 dense
  // PAD path: P_t of the NEXT item is requested (cp.async) into the S buffer as soon as chol(S) has been used for
  // the last time, a quarter of an item ahead of its use; the two buffers then swap roles.  No DRAM round trip
  // on the per-item critical path and no extra shared memory.  m_t / eps_t travel through two registers.
  const T* Pg = a.Ps + unit_off(a.d.Ps, c, b);
  const T* mg = a.ms + unit_off(a.d.ms, c, b);
  const T* eg = a.eps + unit_off(a.d.eps, c, b);
  T m_n = T(0), e_n = T(0);
  // 2-element vector accesses to the compact global rows need even dx and 2-element aligned bases
  const bool pairs_ok = PAD && (dx & 1) == 0 && a.Qs != nullptr &&
                        (reinterpret_cast<uintptr_t>(a.Qs + unit_off(a.d.Qs, c, b)) % (2 * sizeof(T))) == 0 &&
                        (reinterpret_cast<uintptr_t>(a.steps) % (2 * sizeof(T))) == 0;
  // PAD path: the four advancing global pointers of an item (next P_t / m_t / eps_t to prefetch, current output slot)
  // live in the item's spare shared words, not in registers: they are touched twice per item, and keeping them (plus
  // the indices they derive from) live across the factorisations made ptxas spill to local memory, whose reloads
  // missed the small L1 the streaming P tiles leave over (ncu: 8 % of the kernel in long-scoreboard stalls on LDL).
  volatile unsigned long long* ctx = reinterpret_cast<volatile unsigned long long*>(s.dg + CholWs<D>::WORDS);
  auto pad_prefetch = [&]() {
    const T* Pn = reinterpret_cast<const T*>(ctx[0]);
    const T* mn = reinterpret_cast<const T*>(ctx[1]);
    const T* en = reinterpret_cast<const T*>(ctx[2]);
    warp_stage_mat_async<T, LD>(s.S, Pn, dx);
    if (i < dx) { m_n = mn[i]; e_n = en[i]; }
    __syncwarp();
    if (i == 0) {
      ctx[0] = reinterpret_cast<unsigned long long>(Pn + a.d.Ps.time);
      ctx[1] = reinterpret_cast<unsigned long long>(mn + a.d.ms.time);
      ctx[2] = reinterpret_cast<unsigned long long>(en + a.d.eps.time);
    }
  };
  if (PAD) {
    if (i == 0) {
      ctx[0] = reinterpret_cast<unsigned long long>(Pg + (int64_t)t_lo * a.d.Ps.time);
      ctx[1] = reinterpret_cast<unsigned long long>(mg + (int64_t)t_lo * a.d.ms.time);
      ctx[2] = reinterpret_cast<unsigned long long>(eg + (int64_t)t_lo * a.d.eps.time);
      ctx[3] = reinterpret_cast<unsigned long long>(a.steps + ((int64_t)ul * Tn + t_lo) * ssz);
    }
    __syncwarp();
    if (t_lo < t_hi) pad_prefetch();
  }
  auto bulk_stage = [&](int tt) {  // P_tt -> the S buffer, rows by the TMA unit
    fence_async_smem();
    cta_sync<D>();
    if (i == 0) mbar_expect(mbar, (unsigned)(dx * dx * sizeof(T)));
    cta_sync<D>();
    if (i < dx) {
      bulk_g2s(s.S + i * LD, Pg + (int64_t)tt * a.d.Ps.time + (int64_t)i * dx, (unsigned)(dx * sizeof(T)), mbar);
      m_n = mg[(int64_t)tt * a.d.ms.time + i];
      e_n = eg[(int64_t)tt * a.d.eps.time + i];
    }
  };
  if (BULK && bulk_ok && t_lo < t_hi) bulk_stage(t_lo);
  for (int it = 0; it < kTileLen; ++it) {
    if (W > 1) __syncthreads();  // keeps the CTA's warps on the same instruction-cache lines
    const int t = t_lo + it;
    if (t >= t_hi) { if (W > 1) continue; else break; }
    T* g = PAD ? nullptr : a.steps + ((int64_t)ul * Tn + t) * ssz;
    cta_sync<D>();
    if (PAD) {
      cp_async_wait_all();
      T* sw = s.P; s.P = s.S; s.S = sw;
      if (i < dx) { s.m[i] = m_n; s.eps[i] = e_n; }
      cta_sync<D>();
    } else if (BULK && bulk_ok) {
      mbar_wait(mbar, mphase);
      mphase ^= 1u;
      T* sw = s.P; s.P = s.S; s.S = sw;
      if (i < dx) { s.m[i] = m_n; s.eps[i] = e_n; }
      cta_sync<D>();
    } else {
      if (i < dx) {
        s.m[i] = mg[(int64_t)t * a.d.ms.time + i];
        s.eps[i] = eg[(int64_t)t * a.d.eps.time + i];
      }
      load_mat<T, D, false>(Pg + (int64_t)t * a.d.Ps.time, s.P, dx, dx);
    }
    if (t == Tn - 1) {
      // last step: x_{T-1} = m + chol(P) eps  (sampling.py:116-124)
      if constexpr (PAD) {
        const T le = chol_times_vec_pad<T, D>(s.P, s.dg, s.eps, dx);
        g = reinterpret_cast<T*>(ctx[3]);
        if (i < dx) g[dx * dx + i] = s.m[i] + le;
      } else if constexpr (BULK) {
        const T le = chol_times_vec64<T>(s.P, s.dg, s.eps, dx);
        if (i < dx) g[dx * dx + i] = s.m[i] + le;
      } else {
        chol_nan_to_zero<T, D, PAD>(s.P, s.dg, dx);
        mv_nn<T, D, 1>(s.P, s.eps, s.e, s.m, dx, dx);
        if (i < dx) g[dx * dx + i] = s.e[i];
      }
      continue;
    }
    if (LEAN) {
      // time-invariant diagonal F: S_ab = f_a f_b P_ab + Q_ab, (F P)_ab = f_a P_ab, r = f m + b   (O(d^2))
      if (first_dyn) {
        const T* Fg = a.Fs + unit_off(a.d.Fs, c, b);
        if (i < dx) { fd[i] = Fg[i * dx + i]; s.bv[i] = a.bs[unit_off(a.d.bs, c, b) + i]; }
        first_dyn = false;
      }
      cta_sync<D>();
      if (PAD && q_shared) {
        lean_predict_rows<T, D>(s.P, fd, sQ, s.S, s.Z, dx);
      } else if (D == 64 && (dx % 4) == 0 &&
                 (reinterpret_cast<uintptr_t>(a.Qs + unit_off(a.d.Qs, c, b)) % (4 * sizeof(T))) == 0) {
        lean_predict_chunks<T, D>(s.P, fd, a.Qs + unit_off(a.d.Qs, c, b), s.S, s.Z, dx);
      } else {
        const T* Qg = a.Qs + unit_off(a.d.Qs, c, b);
        AUXSSM_FOR_ROWS(ri, rj, dx, dx) {
          const T fa = fd[ri], pv = s.P[ri * LD + rj];
          s.S[ri * LD + rj] = fma_(fa * fd[rj], pv, Qg[ri * dx + rj]);
          s.Z[ri * LD + rj] = fa * pv;
        }
      }
      if (i < dx) s.r[i] = fma_(fd[i], s.m[i], s.bv[i]);
      cta_sync<D>();
    } else {
      if (first_dyn || a.d.Fs.time != 0)
        load_mat<T, D, true>(a.Fs + unit_off(a.d.Fs, c, b) + (int64_t)t * a.d.Fs.time, s.Ft, dx, dx);
      if (first_dyn || a.d.Qs.time != 0)
        load_mat<T, D, false>(a.Qs + unit_off(a.d.Qs, c, b) + (int64_t)t * a.d.Qs.time, s.Q, dx, dx);
      if (first_dyn || a.d.bs.time != 0) {
        if (i < dx) s.bv[i] = a.bs[unit_off(a.d.bs, c, b) + (int64_t)t * a.d.bs.time + i];
      }
      first_dyn = false;
      cta_sync<D>();
      if (fdiag == 0) {
        int off = (a.d.Fs.time != 0) ? 1 : 0;
        AUXSSM_FOR_ROWS(ri, rj, dx, dx) if (ri != rj && s.Ft[ri * LD + rj] != T(0)) off = 1;
        fdiag = (__syncthreads_or(off) == 0) ? 1 : -1;
      }
      if (fdiag > 0) {
        AUXSSM_FOR_ROWS(ri, rj, dx, dx) {
          const T fa = s.Ft[ri * LD + ri], pv = s.P[ri * LD + rj];
          s.S[ri * LD + rj] = fma_(fa * s.Ft[rj * LD + rj], pv, s.Q[ri * LD + rj]);
          s.Z[ri * LD + rj] = fa * pv;
        }
        if (i < dx) s.r[i] = fma_(s.Ft[i * LD + i], s.m[i], s.bv[i]);
        cta_sync<D>();
      } else {
        mm_tn<T, D, 0>(s.P, s.Ft, s.W, nullptr, dx);   // W = P F^T
        mm_tn<T, D, 1>(s.Ft, s.W, s.S, s.Q, dx);       // S = F P F^T + Q
        symmetrize<T, D>(s.S, dx);
        mm_tn<T, D, 0>(s.Ft, s.P, s.Z, nullptr, dx);   // Z = F P
        mv_tn<T, D, 1>(s.Ft, s.m, s.r, s.bv, dx, dx);  // r = F m + b
      }
    }
    if constexpr (PAD) chol_reg<T, D, true, true>(s.S, nullptr, s.dg + D, s.dg + 2 * D, dx);  // lower triangle + 1/diag only
    else if constexpr (BULK) chol_rows64<T, sizeof(T) == 4>(s.S, s.dg, s.dg + D, s.dg + 2 * D, dx);
    else chol_inplace<T, D, PAD>(s.S, s.dg, dx);
    // (a fused forward + backward substitution that keeps the column in registers and stores G^T straight to global
    // memory was measured: 11.06 ms per call instead of 10.34 -- 192 bytes of spills at the 128-register cap)
    fwd_solve<T, D, PAD>(s.S, s.Z, nullptr, dx, dx, s.dg);  // Zs = Ls^{-1} F P
    // Sigma = P - Zs^T Zs  (the lower triangle feeds the Cholesky)
    if constexpr (LEAN && (D == 32 || D == 64) && sizeof(T) == 4)
      ztz_sub_3xtf32<D, LD>(reinterpret_cast<const float*>(s.Z), reinterpret_cast<float*>(s.P));
    else mm_tn<T, D, 2>(s.Z, s.Z, s.P, s.P, PAD ? D : dx);
    bwd_solve<T, D, PAD>(s.S, s.Z, s.dg, dx, dx);           // Z <- G^T = Ls^{-T} Zs
    if (PAD && t + 1 < t_hi) pad_prefetch();                // chol(S) is dead: its buffer receives P_{t+1}
    if (BULK && bulk_ok && t + 1 < t_hi) bulk_stage(t + 1);  // chol(S) is dead: the TMA unit refills its buffer
    T l_eps = T(0);
    if constexpr (PAD) l_eps = chol_times_vec_pad<T, D>(s.P, s.dg, s.eps, dx);  // L eps, L = chol(Sigma) (NaN -> 0)
    else if constexpr (BULK) l_eps = chol_times_vec64<T>(s.P, s.dg, s.eps, dx);
    else chol_nan_to_zero<T, D, PAD>(s.P, s.dg, dx);        // L = chol(Sigma), NaN -> 0
    if (PAD) {
      mv_tn_pad<T, D, 2>(s.Z, s.r, s.e, s.m);               // e = m - G (F m + b)     (G v = (G^T)^T v)
      if (i < D) s.e[i] += l_eps;                           // e += L eps
      cta_sync<D>();
    } else if (BULK) {
      mv_tn<T, D, 2>(s.Z, s.r, s.e, s.m, dx, dx);
      if (i < dx) s.e[i] += l_eps;
      cta_sync<D>();
    } else {
      mv_tn<T, D, 2>(s.Z, s.r, s.e, s.m, dx, dx);
      mv_nn<T, D, 1>(s.P, s.eps, s.e, s.e, dx, dx);
    }
    if (PAD) g = reinterpret_cast<T*>(ctx[3]);
    if (PAD && pairs_ok) store_rows_pairs<T, D>(g, s.Z, dx);
    else if (BULK && (dx % 4) == 0 && (reinterpret_cast<uintptr_t>(g) % (4 * sizeof(T))) == 0) store_chunks<T, D>(g, s.Z, dx);
    else store_mat<T, D>(g, s.Z, dx, dx);
    if (i < dx) g[dx * dx + i] = s.e[i];
    if (PAD) {
      __syncwarp();
      if (i == 0) ctx[3] = reinterpret_cast<unsigned long long>(g + ssz);
    }
  }
}

// pass 2: x_{T-1} = e_{T-1};  x_t = G_t x_{t+1} + e_t.  One CTA per unit (one warp for D <= 32, two for D = 64), thread i
// owns x_t[i] and column i of G_t^T (coalesced over i).  The loop is strictly sequential and streams (d^2 + d) s bytes per
// step, so its speed is the number of bytes in flight (Little's law): the columns of the next steps are requested
// into registers before the product of step t is formed (one step ahead gave 3.3 TB/s = 51 % of the HBM peak, two
// 4.3 TB/s; D <= 32 now keeps three in flight).
template <typename T, int D> struct ReplayCfg { static constexpr int NT = (D <= 32) ? 32 : 64; };

template <typename T, int D>
__global__ void __launch_bounds__(ReplayCfg<T, D>::NT) sample_replay_kernel(SampleArgs<T> a) {
  constexpr int NT = ReplayCfg<T, D>::NT;
  const int ul = blockIdx.x, unit = a.unit0 + ul, dx = a.d.dx, ssz = dx * dx + dx, Tn = a.d.T;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  __shared__ __align__(16) T xb[2][D];
  const int i = threadIdx.x;
  const bool act = i < dx;
  const T* base = a.steps + (int64_t)ul * Tn * ssz;
  T* xo = a.xs + unit_off(a.d.xs, c, b);
  auto sync = [&]() { if (NT == 32) __syncwarp(); else __syncthreads(); };
  if (i < D) { xb[0][i] = T(0); xb[1][i] = T(0); }
  sync();
  T x = T(0);
  if (act) {
    x = base[(int64_t)(Tn - 1) * ssz + dx * dx + i];
    xb[0][i] = x;
    xo[(int64_t)(Tn - 1) * a.d.xs.time + i] = x;
  }
  auto fetch = [&](int tt, T (&gn)[D], T& en) {
    if (tt >= 0 && act) {
      const T* g = base + (int64_t)tt * ssz;  // G^T (row k, column i) then e
#pragma unroll
      for (int k = 0; k < D; ++k) if (k < dx) gn[k] = g[k * dx + i];
      en = g[dx * dx + i];
    }
  };
  constexpr int V = (D >= 4) ? 4 : D;
  if constexpr (D <= 32) {
    // One warp: a ring of NB register buffers, the time loop unrolled NB-fold so that buffer s always serves steps
    // t0 - s, t0 - s - NB, ... (no register rotation); a buffer is refilled for NB steps later as soon as its
    // multiply-adds have been issued, which keeps NB - 1 steps (3 x 3.7 KB per warp at d = 30) in flight during a
    // product instead of two.  Same accumulation order as before: bit-identical draws.
    constexpr int NB = 4;
    T g[NB][D], e[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      e[s] = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) g[s][k] = T(0);
    }
#pragma unroll
    for (int s = 0; s < NB; ++s) fetch(Tn - 2 - s, g[s], e[s]);
    sync();
    int cur = 0;
    for (int t0 = Tn - 2; t0 >= 0; t0 -= NB) {
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        const int t = t0 - s;
        if (t >= 0) {  // warp-uniform
          T acc0 = e[s], acc1 = T(0);
#pragma unroll
          for (int k = 0; k < D; k += V) {
            T xv[V];
            ldv<T, V>(xv, &xb[cur][k]);
#pragma unroll
            for (int q = 0; q < V; ++q) {
              if (q & 1) acc1 = fma_(g[s][k + q], xv[q], acc1); else acc0 = fma_(g[s][k + q], xv[q], acc0);
            }
          }
          fetch(t - NB, g[s], e[s]);
          x = acc0 + acc1;
          if (act) {
            xb[cur ^ 1][i] = x;
            xo[(int64_t)t * a.d.xs.time + i] = x;
          }
          sync();
          cur ^= 1;
        }
      }
    }
    return;
  }
  T g1[D], g2[D], e1 = T(0), e2 = T(0);  // D = 64 (two warps): steps t - 1 and t - 2 in flight
#pragma unroll
  for (int k = 0; k < D; ++k) { g1[k] = T(0); g2[k] = T(0); }
  fetch(Tn - 2, g1, e1);
  fetch(Tn - 3, g2, e2);
  sync();
  int cur = 0;
  for (int t = Tn - 2; t >= 0; --t) {
    T gc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { gc[k] = g1[k]; g1[k] = g2[k]; }
    const T ec = e1;
    e1 = e2;
    fetch(t - 2, g2, e2);
    T acc0 = ec, acc1 = T(0);
#pragma unroll
    for (int k = 0; k < D; k += V) {
      T xv[V];
      ldv<T, V>(xv, &xb[cur][k]);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (q & 1) acc1 = fma_(gc[k + q], xv[q], acc1); else acc0 = fma_(gc[k + q], xv[q], acc0);
      }
    }
    x = acc0 + acc1;
    if (act) {
      xb[cur ^ 1][i] = x;
      xo[(int64_t)t * a.d.xs.time + i] = x;
    }
    sync();
    cur ^= 1;
  }
}

template <typename T> __global__ void poison_if_gate_kernel(const int* gate, T* p, int64_t n) {
  if (*gate == 0) return;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) p[e] = T(NAN);
}

struct SamplePlan {
  size_t per_unit;  // bytes of scratch per unit
  int wave;         // units per wave
  size_t total;
};

static SamplePlan make_sample_plan(const auxssm_lgssm_desc* d, size_t es) {
  SamplePlan p;
  const size_t units = (size_t)d->n_chains * d->n_batch;
  p.per_unit = (size_t)d->T * ((size_t)d->dx * d->dx + d->dx) * es;
  size_t wave = kScratchCap / p.per_unit;
  if (wave < 1) wave = 1;
  if (wave > units) wave = units;
  p.wave = (int)wave;
  p.total = kGateBytes + (p.per_unit * wave + 255) / 256 * 256;  // [0, kGateBytes): the diagonal-F gate word
  return p;
}

template <typename T, int D> static int launch_sample(cudaStream_t st, SampleArgs<T>& a, int wave, int units) {
  constexpr int NT = Tile<D>::NT;
  const bool lean = (a.d.flags & AUXSSM_FLAG_F_IS_DIAGONAL) != 0 && a.d.T > 1 && a.d.Fs.time == 0 &&
                    a.d.Qs.time == 0 && a.d.bs.time == 0;
  int rc = AUXSSM_OK;
  a.gate = lean ? a.gate : nullptr;
  if (lean) {
    rc = launch_count_offdiag<T>(st, a.Fs, a.d, a.gate);
    if (rc) return rc;
  }
  constexpr int WL = ElemCfg<T, D, true>::W, WG = ElemCfg<T, D, false>::W;
  const size_t smem_l = samp_smem_elems<D>(true) * sizeof(T) * WL + (D <= 32 ? (size_t)D * Tile<D>::LD * sizeof(T) : 0);
  const size_t smem_g = samp_smem_elems<D>(false) * sizeof(T) * WG;
  if (lean) rc = ensure_smem(sample_elements_kernel<T, D, true>, smem_l, "auxssm_kalman_sample");
  if (rc) return rc;
  // The general kernel does not fit the shared memory of one CTA at d = 64 in fp64.  Without the hint that is an
  // AUXSSM_ERR_UNSUPPORTED; with the hint the structured kernel runs under the gate and, should the hint turn out
  // to be false, a poison kernel fills the scratch with NaN instead: the call then returns NaN trajectories
  // (loudly wrong), never silently wrong ones.
  const bool general_fits = smem_g <= device_max_smem();
  if (general_fits || !lean) {
    rc = ensure_smem(sample_elements_kernel<T, D, false>, smem_g, "auxssm_kalman_sample");
    if (rc) return rc;
  }
  const int n_tiles = (a.d.T + kTileLen - 1) / kTileLen;
  for (int u0 = 0; u0 < units; u0 += wave) {
    a.unit0 = u0;
    a.n_units = (units - u0 < wave) ? units - u0 : wave;
    const int64_t items = (int64_t)a.n_units * n_tiles;
    if (lean) {
      sample_elements_kernel<T, D, true><<<(unsigned)((items + WL - 1) / WL), dim3(NT, WL), smem_l, st>>>(a, n_tiles);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_sample (elements, diagonal F)");
    }
    if (general_fits) {
      sample_elements_kernel<T, D, false><<<(unsigned)((items + WG - 1) / WG), dim3(NT, WG), smem_g, st>>>(a, n_tiles);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_sample (elements)");
    } else {
      const int64_t n = (int64_t)a.n_units * a.d.T * ((int64_t)a.d.dx * a.d.dx + a.d.dx);
      poison_if_gate_kernel<T><<<592, 256, 0, st>>>(a.gate, a.steps, n);
      AUXSSM_CHECK_LAUNCH("auxssm_kalman_sample (poison on a false diagonal-F hint)");
    }
    sample_replay_kernel<T, D><<<a.n_units, ReplayCfg<T, D>::NT, 0, st>>>(a);
    AUXSSM_CHECK_LAUNCH("auxssm_kalman_sample (replay)");
  }
  return AUXSSM_OK;
}

template <typename T>
static int sample_entry(void* stream, const auxssm_lgssm_desc* d, const T* ms, const T* Ps, const T* Fs, const T* Qs,
                        const T* bs, const T* eps, T* xs, void* scratch, size_t scratch_bytes) {
  const char* who = "auxssm_kalman_sample";
  int rc = check_desc(d, who);
  if (rc) return rc;
  if (!ms || !Ps || !eps || !xs || (d->T > 1 && (!Fs || !Qs || !bs))) {
    set_error("%s: null data pointer", who);
    return AUXSSM_ERR_INVALID;
  }
  SamplePlan p = make_sample_plan(d, sizeof(T));
  if (scratch == nullptr || scratch_bytes < kGateBytes + p.per_unit) {
    set_error("%s: scratch too small (%zu bytes; one unit needs %zu)", who, scratch_bytes, kGateBytes + p.per_unit);
    return AUXSSM_ERR_SCRATCH;
  }
  int wave = p.wave;
  if (scratch_bytes < p.total) wave = (int)((scratch_bytes - kGateBytes) / p.per_unit);  // smaller scratch => more waves
  SampleArgs<T> a;
  a.d = *d;
  a.ms = ms; a.Ps = Ps; a.Fs = Fs; a.Qs = Qs; a.bs = bs; a.eps = eps; a.xs = xs;
  a.gate = static_cast<int*>(scratch);
  a.steps = reinterpret_cast<T*>(static_cast<unsigned char*>(scratch) + kGateBytes);
  a.unit0 = 0; a.n_units = 0;
  const int units = d->n_chains * d->n_batch;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (padded_dim(d->dx)) {
    case 4: return launch_sample<T, 4>(st, a, wave, units);
    case 8: return launch_sample<T, 8>(st, a, wave, units);
    case 16: return launch_sample<T, 16>(st, a, wave, units);
    case 32: return launch_sample<T, 32>(st, a, wave, units);
    case 64: return launch_sample<T, 64>(st, a, wave, units);
    default: set_error("%s: unsupported dimension %d", who, d->dx); return AUXSSM_ERR_UNSUPPORTED;
  }
}

}  // namespace auxssm

// This translation unit carries the fp32 entry point (and the dtype-independent scratch query); the fp64
// instantiations are compiled by kalman_sample_f64.cu, which includes this file with AUXSSM_ONLY_F64 defined, so
// that the two halves build in parallel.
#ifndef AUXSSM_ONLY_F64
extern "C" __attribute__((visibility("default"))) size_t auxssm_kalman_sample_scratch_bytes(const auxssm_lgssm_desc* desc, int32_t elem_size) {
  if (auxssm::check_desc(desc, "auxssm_kalman_sample_scratch_bytes")) return 0;
  return auxssm::make_sample_plan(desc, (size_t)elem_size).total;
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_sample_f32(void* stream, const auxssm_lgssm_desc* desc, const float* ms,
                                            const float* Ps, const float* Fs, const float* Qs, const float* bs,
                                            const float* eps, float* xs, void* scratch, size_t scratch_bytes) {
  return auxssm::sample_entry<float>(stream, desc, ms, Ps, Fs, Qs, bs, eps, xs, scratch, scratch_bytes);
}
#else
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_sample_f64(void* stream, const auxssm_lgssm_desc* desc, const double* ms,
                                            const double* Ps, const double* Fs, const double* Qs, const double* bs,
                                            const double* eps, double* xs, void* scratch, size_t scratch_bytes) {
  return auxssm::sample_entry<double>(stream, desc, ms, Ps, Fs, Qs, bs, eps, xs, scratch, scratch_bytes);
}
#endif

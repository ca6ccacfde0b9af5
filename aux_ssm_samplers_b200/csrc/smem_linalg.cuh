// CTA-cooperative dense linear algebra on small matrices held in shared memory.
//
// One CTA owns one (chain, inner-batch, time-chunk) unit; all d x d operands of a
// Kalman step / scan-element combine live in shared memory in a padded row-major
// layout with a compile-time leading dimension Tile<D>::LD, where D is max(dx, dy)
// rounded up to {4, 8, 16, 32, 64}.  INVARIANT: every matrix buffer is zero outside
// its logical [rows x cols] block, so products may run over whole register tiles.
//
// The product primitive is  C = X^T Y  with BOTH operands "k-major" (row k of X and of
// Y is contiguous): each thread owns a TM x TN register tile and reads its X/Y slices
// with one vector LDS each per k.  Every product on the hot path is phrased in that
// form (transposed copies of F and H are staged once; P, Q, R, S, C, J are symmetric).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace auxssm {

template <int D> struct Tile;
template <> struct Tile<4>  { static constexpr int LD = 4,  NT = 32,  TM = 1, TN = 1; };
template <> struct Tile<8>  { static constexpr int LD = 12, NT = 32,  TM = 1, TN = 2; };
template <> struct Tile<16> { static constexpr int LD = 20, NT = 32,  TM = 2, TN = 4; };
// D = 32 runs as ONE warp per work item (8 x 4 register tiles, __syncwarp only): with two warps one of them
// idled at a block barrier during every Cholesky / triangular-solve phase.
template <> struct Tile<32> { static constexpr int LD = 36, NT = 32,  TM = 8, TN = 4; };
// D = 64: two warps per work item.  The factorisations and substitutions keep one warp / one thread per column
// busy whatever the CTA size, so a 256-thread CTA (255 registers: one CTA per SM) left 7 of its 8 warps at a
// barrier 90 % of the time; 64 threads leave room for four items per SM (shared-memory bound).
template <> struct Tile<64> { static constexpr int LD = 68, NT = 64, TM = 8, TN = 8; };  // one 8 x 8 register tile per thread

// Resident CTAs per SM the kernels are compiled for: caps registers at 128/thread for D <= 32 so that
// 7-8 chains per SM stay co-resident (1024 chains = one wave on 148 SMs).
template <typename T, int D> struct Occ {
  static constexpr int MINB = (D <= 32) ? (sizeof(T) == 4 ? 512 : 256) / Tile<D>::NT : 1;
};

// Barrier over the Tile<D>::NT threads of ONE work item.  Items are laid out along threadIdx.y when a CTA hosts
// several of them (sampler elements kernel), so multi-warp items use the named barrier 1 + threadIdx.y; with one
// item per CTA that is simply a block barrier.
template <int D> __device__ __forceinline__ void cta_sync() {
  if (Tile<D>::NT == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)threadIdx.y), "n"(Tile<D>::NT) : "memory");
}

template <typename T, int N> struct alignas((sizeof(T) * N > 16) ? 16 : sizeof(T) * N) Vec { T v[N]; };

template <typename T, int N> __device__ __forceinline__ void ldv(T (&dst)[N], const T* p) {
  Vec<T, N> t = *reinterpret_cast<const Vec<T, N>*>(p);
#pragma unroll
  for (int i = 0; i < N; ++i) dst[i] = t.v[i];
}
template <typename T, int N> __device__ __forceinline__ void stv(T* p, const T (&src)[N]) {
  Vec<T, N> t;
#pragma unroll
  for (int i = 0; i < N; ++i) t.v[i] = src[i];
  *reinterpret_cast<Vec<T, N>*>(p) = t;
}

template <typename T> __device__ __forceinline__ T fma_(T a, T b, T c);
template <> __device__ __forceinline__ float fma_<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_<double>(double a, double b, double c) { return fma(a, b, c); }

template <typename T> __device__ __forceinline__ bool finite_(T x) { return isfinite(x); }

// Packed pairs.  fp32: one aligned 64-bit register pair driven by FFMA2 (fma.rn.f32x2, sm_100a), two FMAs per
// issue slot.  The arithmetic rate is the same as scalar FFMA (measured: profiles/micro/ffma2_probe.cu), but
// the kernels here are issue-bound, so halving the FMA instruction count frees slots for the loads and
// shuffles around them.  Each half is an ordinary round-to-nearest fp32 FMA: results are bit-identical to the
// scalar form.  fp64 keeps two scalar DFMAs behind the same interface.
template <typename T> struct Pair;
template <> struct Pair<float> { unsigned long long v; };
template <> struct Pair<double> { double lo, hi; };

__device__ __forceinline__ Pair<float> pair_make(float a, float b) {
  Pair<float> p;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p.v) : "f"(a), "f"(b));
  return p;
}
__device__ __forceinline__ Pair<double> pair_make(double a, double b) { return Pair<double>{a, b}; }
__device__ __forceinline__ void pair_get(Pair<float> p, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
}
__device__ __forceinline__ void pair_get(Pair<double> p, double& a, double& b) { a = p.lo; b = p.hi; }
// a * b + c, element-wise
__device__ __forceinline__ Pair<float> pair_fma(Pair<float> a, Pair<float> b, Pair<float> c) {
  Pair<float> r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ Pair<double> pair_fma(Pair<double> a, Pair<double> b, Pair<double> c) {
  return Pair<double>{fma(a.lo, b.lo, c.lo), fma(a.hi, b.hi, c.hi)};
}
// a * b, element-wise, and s * b
__device__ __forceinline__ Pair<float> pair_mul(Pair<float> a, Pair<float> b) {
  Pair<float> r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ Pair<double> pair_mul(Pair<double> a, Pair<double> b) {
  return Pair<double>{a.lo * b.lo, a.hi * b.hi};
}
// s * b + c with the scalar s applied to both halves (FFMA2 takes it as a broadcast operand, no extra move)
template <typename T> __device__ __forceinline__ Pair<T> pair_fma_s(T s, Pair<T> b, Pair<T> c) {
  return pair_fma(pair_make(s, s), b, c);
}
template <typename T> __device__ __forceinline__ Pair<T> pair_mul_s(T s, Pair<T> b) {
  return pair_mul(pair_make(s, s), b);
}

// Ampere-style asynchronous global -> shared copies (LDGSTS): the data lands in shared memory without passing
// through registers and without stalling the issuing warp; `cp_async_wait_all` + a barrier publish it.
template <int BYTES> __device__ __forceinline__ void cp_async(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(g), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completed on an mbarrier --------------------------------------
// One elected thread arms the barrier with the byte count (`mbar_expect`), any thread may then issue row copies of a
// multiple of 16 bytes between 16-byte aligned addresses; consumers spin on the phase parity.  The copy engine writes
// shared memory through the async proxy: `fence_async_smem` orders the CTA's earlier generic accesses of the
// destination before it.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One warp: compact global [d][d] -> padded shared [d][LD] (LD even), asynchronously.  Even d: 2 elements per copy.
template <typename T, int LD> __device__ __forceinline__ void warp_stage_mat_async(T* dst, const T* src, int d) {
  const unsigned magic = 0xffffffffu / (unsigned)d + 1u;  // e / d exactly for e < 2^13, d <= 32
  const int lane = threadIdx.x & 31, total = d * d;
  if ((d & 1) == 0 && (reinterpret_cast<uintptr_t>(src) % (2 * sizeof(T))) == 0) {
#pragma unroll 4
    for (int e = 2 * lane; e < total; e += 64) {
      const int n = (int)__umulhi((unsigned)e, magic), k = e - n * d;
      cp_async<2 * sizeof(T)>(dst + n * LD + k, src + e);
    }
  } else {
#pragma unroll 4
    for (int e = lane; e < total; e += 32) {
      const int n = (int)__umulhi((unsigned)e, magic), k = e - n * d;
      cp_async<sizeof(T)>(dst + n * LD + k, src + e);
    }
  }
}

// MODE 0: C = X^T Y ; MODE 1: C = E + X^T Y ; MODE 2: C = E - X^T Y.   K = inner extent.
// X, Y, C, E are D x LD shared buffers; C may alias E but must not alias X or Y.
template <typename T, int D, int MODE>
__device__ __forceinline__ void mm_tn(const T* __restrict__ X, const T* __restrict__ Y, T* C, const T* E, int K) {
  using Tl = Tile<D>;
  constexpr int LD = Tl::LD, TM = Tl::TM, TN = Tl::TN, NT = Tl::NT;
  constexpr int TX = D / TN, TILES = (D / TM) * TX;
  cta_sync<D>();
  for (int tile = threadIdx.x; tile < TILES; tile += NT) {
    const int ty = tile / TX, tx = tile % TX;
    T acc[TM][TN];
    const T* xp = X + ty * TM;
    const T* yp = Y + tx * TN;
    if constexpr (TN % 2 == 0) {
      // accumulators as column pairs: one packed FMA per (row, column pair)
      Pair<T> acc2[TM][TN / 2];
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN / 2; ++c) acc2[r][c] = pair_make(T(0), T(0));
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        T xa[TM], yb[TN];
        ldv<T, TM>(xa, xp + k * LD);
        ldv<T, TN>(yb, yp + k * LD);
        Pair<T> yb2[TN / 2];
#pragma unroll
        for (int c = 0; c < TN / 2; ++c) yb2[c] = pair_make(yb[2 * c], yb[2 * c + 1]);
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
          for (int c = 0; c < TN / 2; ++c) acc2[r][c] = pair_fma_s(xa[r], yb2[c], acc2[r][c]);
      }
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN / 2; ++c) pair_get(acc2[r][c], acc[r][2 * c], acc[r][2 * c + 1]);
    } else {
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = T(0);
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        T xa[TM], yb[TN];
        ldv<T, TM>(xa, xp + k * LD);
        ldv<T, TN>(yb, yp + k * LD);
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
          for (int c = 0; c < TN; ++c) acc[r][c] = fma_(xa[r], yb[c], acc[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      T* cp = C + (ty * TM + r) * LD + tx * TN;
      T out[TN];
      if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < TN; ++c) out[c] = acc[r][c];
      } else {
        T e[TN];
        ldv<T, TN>(e, E + (ty * TM + r) * LD + tx * TN);
#pragma unroll
        for (int c = 0; c < TN; ++c) out[c] = (MODE == 1) ? e[c] + acc[r][c] : e[c] - acc[r][c];
      }
      stv<T, TN>(cp, out);
    }
  }
  cta_sync<D>();
}

// y[0..M) (op)= A^T x  where A is k-major: y[i] = sum_k A[k][i] x[k], k < K.  MODE as mm_tn with e.
template <typename T, int D, int MODE>
__device__ __forceinline__ void mv_tn(const T* A, const T* x, T* y, const T* e, int K, int M) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  T acc = T(0);
  const int i = threadIdx.x;
  if (i < M) {
    for (int k = 0; k < K; ++k) acc = fma_(A[k * LD + i], x[k], acc);
  }
  T out = acc;
  if (MODE == 1 && i < M) out = e[i] + acc;
  if (MODE == 2 && i < M) out = e[i] - acc;
  cta_sync<D>();  // y may alias x or e
  if (i < M) y[i] = out;
  cta_sync<D>();
}

// y = A x with A row-major (non k-major): y[i] = sum_k A[i][k] x[k].
template <typename T, int D, int MODE>
__device__ __forceinline__ void mv_nn(const T* A, const T* x, T* y, const T* e, int M, int K) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  T acc = T(0);
  const int i = threadIdx.x;
  if (i < M) {
    for (int k = 0; k < K; ++k) acc = fma_(A[i * LD + k], x[k], acc);
  }
  T out = acc;
  if (MODE == 1 && i < M) out = e[i] + acc;
  if (MODE == 2 && i < M) out = e[i] - acc;
  cta_sync<D>();
  if (i < M) y[i] = out;
  cta_sync<D>();
}

template <typename T, int D> __device__ __forceinline__ void zero_mat(T* A) {
  constexpr int LD = Tile<D>::LD, NT = Tile<D>::NT;
  for (int i = threadIdx.x; i < D * LD; i += NT) A[i] = T(0);
}

// ---- element-wise helpers: warp w owns rows w, w + NW, ...; lanes run over columns.  No integer
// division, global accesses are row-contiguous (coalesced), shared accesses are conflict-free.
// For D <= 32 a row fits one warp pass: the column loop runs at most once, which the compiler can turn
// into a predicate (no divergence bookkeeping); the row loop is unrolled by four.
#define AUXSSM_FOR_ROWS(r, c, rows, cols)                                                   \
  _Pragma("unroll 4")                                                                       \
  for (int r = (int)(threadIdx.x >> 5); r < (rows); r += Tile<D>::NT / 32)                 \
    for (int c = (int)(threadIdx.x & 31), once_ = 1; once_ && c < (cols); once_ = (D > 32), c += 32)

// dst = src^T: dst[j][i] = src[i][j] for i < rows, j < cols, zero elsewhere (whole D x D tile).
template <typename T, int D> __device__ __forceinline__ void transpose_to(const T* src, T* dst, int rows, int cols) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  AUXSSM_FOR_ROWS(i, j, D, D) dst[i * LD + j] = (j < rows && i < cols) ? src[j * LD + i] : T(0);
  cta_sync<D>();
}

// A = 0.5 (A + A^T) on the n x n block (reads into registers, barrier, writes).
template <typename T, int D> __device__ __forceinline__ void symmetrize(T* A, int n) {
  constexpr int LD = Tile<D>::LD, NW = Tile<D>::NT / 32, RPW = (D + NW - 1) / NW, CPL = (D + 31) / 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cta_sync<D>();
  T v[RPW * CPL];
#pragma unroll
  for (int qi = 0; qi < RPW; ++qi) {
    const int i = w + qi * NW;
#pragma unroll
    for (int qj = 0; qj < CPL; ++qj) {
      const int j = lane + 32 * qj;
      v[qi * CPL + qj] = (i < n && j < n) ? T(0.5) * (A[i * LD + j] + A[j * LD + i]) : T(0);
    }
  }
  cta_sync<D>();
#pragma unroll
  for (int qi = 0; qi < RPW; ++qi) {
    const int i = w + qi * NW;
#pragma unroll
    for (int qj = 0; qj < CPL; ++qj) {
      const int j = lane + 32 * qj;
      if (i < n && j < n) A[i * LD + j] = v[qi * CPL + qj];
    }
  }
  cta_sync<D>();
}

// Copy the logical block [rows x cols] between shared buffers.
template <typename T, int D> __device__ __forceinline__ void copy_mat(const T* src, T* dst, int rows, int cols) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  AUXSSM_FOR_ROWS(i, j, rows, cols) dst[i * LD + j] = src[i * LD + j];
  cta_sync<D>();
}

// Load a [rows x cols] row-major contiguous global block into shared (optionally transposed).
// All global loads are issued before the first shared store (memory-level parallelism).
template <typename T, int D, bool TRANS>
__device__ __forceinline__ void load_mat(const T* __restrict__ g, T* s, int rows, int cols) {
  constexpr int LD = Tile<D>::LD, NW = Tile<D>::NT / 32, RPW = (D + NW - 1) / NW, CPL = (D + 31) / 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T v[RPW * CPL];
#pragma unroll
  for (int qi = 0; qi < RPW; ++qi) {
    const int i = w + qi * NW;
#pragma unroll
    for (int qj = 0; qj < CPL; ++qj) {
      const int j = lane + 32 * qj;
      v[qi * CPL + qj] = (i < rows && j < cols) ? g[i * cols + j] : T(0);
    }
  }
#pragma unroll
  for (int qi = 0; qi < RPW; ++qi) {
    const int i = w + qi * NW;
#pragma unroll
    for (int qj = 0; qj < CPL; ++qj) {
      const int j = lane + 32 * qj;
      if (i < rows && j < cols) {
        if (TRANS) s[j * LD + i] = v[qi * CPL + qj]; else s[i * LD + j] = v[qi * CPL + qj];
      }
    }
  }
}
template <typename T, int D>
__device__ __forceinline__ void store_mat(T* __restrict__ g, const T* s, int rows, int cols) {
  constexpr int LD = Tile<D>::LD;
  AUXSSM_FOR_ROWS(i, j, rows, cols) g[i * cols + j] = s[i * LD + j];
}
template <typename T, int D>
__device__ __forceinline__ void store_mat_t(T* __restrict__ g, const T* s, int rows, int cols) {
  // global [rows x cols] <- transpose of shared [cols x rows]
  constexpr int LD = Tile<D>::LD;
  AUXSSM_FOR_ROWS(i, j, rows, cols) g[i * cols + j] = s[j * LD + i];
}

// 1 / sqrt(x) and sqrt(x) for a Cholesky pivot.  fp32: MUFU.RSQ + one Newton step (full fp32 accuracy,
// ~35 cycles instead of sqrt + divide on the serial critical path); NaN for x < 0, like sqrt.
template <typename T> __device__ __forceinline__ void pivot_(T x, T& l, T& inv);
template <> __device__ __forceinline__ void pivot_<float>(float x, float& l, float& inv) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU.RSQ, no denormal fix-up code
  r = r * fmaf(-0.5f * x, r * r, 1.5f);
  inv = r;
  l = (x == 0.f) ? 0.f : x * r;
}
template <> __device__ __forceinline__ void pivot_<double>(double x, double& l, double& inv) {
  l = sqrt(x);
  inv = 1.0 / l;
}
// The structured fp32 hot paths (one-warp sampler items, warp / tile filters) use the bare MUFU.RSQ: its documented
// maximum relative error, 2^-22.9 = 1.3e-7, is one fp32 rounding, and the Newton step + zero guard above were 5 of the
// ~17 non-FMA instructions on the serial path of every pivot (64 pivots per sampler item, 32 per filter step).  A zero
// pivot gives l = 0 * inf = NaN, which the callers already treat as a failed factorisation (NaN -> 0, sampling.py:102).
template <typename T> __device__ __forceinline__ void pivot_fast_(T x, T& l, T& inv) { pivot_<T>(x, l, inv); }
template <> __device__ __forceinline__ void pivot_fast_<float>(float x, float& l, float& inv) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  inv = r;
  l = x * r;
}

// Register-resident Cholesky for D <= 32: warp 0 only, lane i owns row i in registers; the freshly
// scaled column k is exchanged through a double-buffered shared vector (one STS + (D-k)/4 broadcast
// LDS.128 per column, __syncwarp only).  `dg` receives diag(L), `dginv` (optional) its reciprocal.
// `colbuf` needs 2*D words.  All D x D loops are fully unrolled so the row stays in registers.
// PADDED = the caller keeps an identity block on the padded diagonal (rows/cols >= n), so all D columns
// can be processed without any run-time guard (branch-free, ~20% fewer instructions at d = 30).
// KEEP_UPPER: leave whatever the elimination left above the diagonal instead of zeroing it on the way out (32 selects
// + predicate packing per lane); for callers that only ever read the lower triangle (the substitutions).
// FUSE_MV: the factor is not written back at all; lane i returns y_i = sum_{j <= i} L_ij x_j in *yreg instead (the
// sampler needs L only to form L eps: saves the masked write-back, the reload and a shared-memory round trip).
// dg / dginv may be null when the caller does not need them (dead stores and the sqrt product disappear).
template <typename T, int D, bool PADDED = false, bool KEEP_UPPER = false, bool FUSE_MV = false>
__device__ __forceinline__ void chol_reg(T* A, T* dg, T* dginv, T* colbuf, int n, const T* xvec = nullptr,
                                         T* yreg = nullptr) {
  constexpr int LD = Tile<D>::LD;
  constexpr int V = (D >= 4) ? 4 : D;
  cta_sync<D>();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    T row[D];
#pragma unroll
    for (int j = 0; j < D; j += V) {
      T tmp[V];
      if (lane < D) ldv<T, V>(tmp, A + lane * LD + j);
#pragma unroll
      for (int q = 0; q < V; ++q) row[j + q] = (PADDED || lane < n) ? tmp[q] : T(0);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
      if (PADDED || k < n) {
        const T akk = __shfl_sync(0xffffffffu, row[k], k);
        T lkk, inv;
        if constexpr (PADDED) pivot_fast_<T>(akk, lkk, inv); else pivot_<T>(akk, lkk, inv);
        // PADDED (fast pivot): l_kk = a_kk * inv is what row[k] * inv gives on lane k itself, no select needed
        const T lik = PADDED ? row[k] * inv : ((lane == k) ? lkk : row[k] * inv);
        row[k] = lik;
        // Column exchange through the double-buffered shared vector.  (Exchanging it with 31 - k shuffles per column
        // instead was measured -- profiles/micro/shfl_vs_lds_probe.cu: SHFL and LDS share a pipe, 0.42 vs 0.68 clocks
        // per broadcast word -- but the sampler call went 10.2 -> 11.5 ms: four times the instructions on the serial
        // path of every column.)
        T* cb = colbuf + (k & 1) * D;
        if (lane < D) cb[lane] = lik;
        if (PADDED ? (lane == 0) : (lane == k)) {  // every lane holds the same pivot: PADDED lets lane 0 publish it
          if (dg) dg[k] = lkk;
          if (dginv) dginv[k] = inv;
        }
        __syncwarp();
        const T nlik = -lik;
#pragma unroll
        for (int j = ((k + 1) / V) * V; j < D; j += V) {
          T c[V];
          ldv<T, V>(c, cb + j);
#pragma unroll
          for (int q = 0; q < V; q += 2) {
            if (q + 1 < V && j + q > k) {  // both columns of the pair are still open: one packed FMA
              const Pair<T> r2 = pair_fma_s(nlik, pair_make(c[q], c[q + 1]), pair_make(row[j + q], row[j + q + 1]));
              pair_get(r2, row[j + q], row[j + q + 1]);
            } else {
              if (j + q > k) row[j + q] = fma_(nlik, c[q], row[j + q]);
              if (q + 1 < V && j + q + 1 > k) row[j + q + 1] = fma_(nlik, c[q + 1], row[j + q + 1]);
            }
          }
        }
      }
    }
    if constexpr (FUSE_MV) {
      T a0 = T(0), a1 = T(0);
#pragma unroll
      for (int j = 0; j < D; j += V) {
        T xv[V];
        ldv<T, V>(xv, xvec + j);
#pragma unroll
        for (int q = 0; q < V; ++q) {
          if (j + q <= lane) {
            if (q & 1) a1 = fma_(row[j + q], xv[q], a1); else a0 = fma_(row[j + q], xv[q], a0);
          }
        }
      }
      *yreg = a0 + a1;
    } else if ((PADDED && lane < D) || lane < n) {
#pragma unroll
      for (int j = 0; j < D; j += V) {
        T tmp[V];
#pragma unroll
        for (int q = 0; q < V; ++q) tmp[q] = (KEEP_UPPER || j + q <= lane) ? row[j + q] : T(0);
        stv<T, V>(A + lane * LD + j, tmp);
      }
    }
  }
  cta_sync<D>();
}

// Register-resident forward substitution for D <= 32: thread j owns column j of B (n x m) in
// registers, rows of L are read as broadcast vector loads, two accumulators per dot product.
// `dginv` = 1 / diag(L).  Thread m handles the optional extra vector `bvec`.
template <typename T, int D, bool PADDED = false>
__device__ __forceinline__ void fwd_solve_reg(const T* L, T* B, T* bvec, const T* dginv, int n, int m) {
  constexpr int LD = Tile<D>::LD;
  constexpr int V = (D >= 4) ? 4 : D;
  cta_sync<D>();
  const int j = threadIdx.x;
  if (bvec != nullptr && m >= Tile<D>::NT) {
    // no spare thread for the extra right-hand side: thread 0 solves it first (m == NT == D only)
    if (j == 0) {
      for (int i = 0; i < n; ++i) {
        T acc = bvec[i];
        for (int k = 0; k < i; ++k) acc = fma_(-L[i * LD + k], bvec[k], acc);
        bvec[i] = acc * dginv[i];
      }
    }
    cta_sync<D>();
  }
  const bool is_vec = (j == m) && (bvec != nullptr);
  if (j < m || is_vec) {
    T z[D];
    if constexpr (D % 4 == 0 && D >= 8) {
      // Four rows at a time.  The dot products L[i+q, :i] . z[:i] (q = 0..3) share the z operands and are four
      // INDEPENDENT accumulation chains (row-at-a-time they were one chain per row: the phase was bound by FMA
      // latency, ncu `wait` stalls 45 %); i is a multiple of 4, so no chunk straddles the diagonal block, which is
      // then solved in registers.  Same loads and FMA count as the row-at-a-time form.
#pragma unroll
      for (int i = 0; i < D; i += 4) {
        if (PADDED || i < n) {
          T rhs[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) rhs[q] = (PADDED || i + q < n) ? (is_vec ? bvec[i + q] : B[(i + q) * LD + j]) : T(0);
          Pair<T> acc[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = pair_make(T(0), T(0));
#pragma unroll
          for (int k = 0; k < i; k += 4) {
            const Pair<T> z01 = pair_make(z[k], z[k + 1]), z23 = pair_make(z[k + 2], z[k + 3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              T l[4];
              ldv<T, 4>(l, L + (i + q) * LD + k);
              acc[q] = pair_fma(pair_make(l[0], l[1]), z01, acc[q]);
              acc[q] = pair_fma(pair_make(l[2], l[3]), z23, acc[q]);
            }
          }
          T d4[4][4];
#pragma unroll
          for (int q = 1; q < 4; ++q) ldv<T, 4>(d4[q], L + (i + q) * LD + i);
          T sum[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            T a0, a1;
            pair_get(acc[q], a0, a1);
            sum[q] = rhs[q] - (a0 + a1);
          }
          z[i] = sum[0] * dginv[i];
          z[i + 1] = fma_(-d4[1][0], z[i], sum[1]) * dginv[i + 1];
          z[i + 2] = fma_(-d4[2][1], z[i + 1], fma_(-d4[2][0], z[i], sum[2])) * dginv[i + 2];
          z[i + 3] = fma_(-d4[3][2], z[i + 2], fma_(-d4[3][1], z[i + 1], fma_(-d4[3][0], z[i], sum[3]))) * dginv[i + 3];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (PADDED || i + q < n) {
              if (is_vec) bvec[i + q] = z[i + q]; else B[(i + q) * LD + j] = z[i + q];
            } else {
              z[i + q] = T(0);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) z[i + q] = T(0);
        }
      }
    } else {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (PADDED || i < n) {
        // dot product L[i, :i] . z[:i] in two interleaved partial sums (even / odd k), packed pairwise
        const T rhs = is_vec ? bvec[i] : B[i * LD + j];
        Pair<T> s2 = pair_make(T(0), T(0));
        T tail = T(0);
#pragma unroll
        for (int k = 0; k < ((i + V - 1) / V) * V; k += V) {
          T l[V];
          ldv<T, V>(l, L + i * LD + k);
#pragma unroll
          for (int q = 0; q < V; q += 2) {
            if (q + 1 < V && k + q + 1 < i) {
              s2 = pair_fma(pair_make(l[q], l[q + 1]), pair_make(z[k + q], z[k + q + 1]), s2);
            } else if (k + q < i) {
              tail = fma_(l[q], z[k + q], tail);
            }
          }
        }
        T s_even, s_odd;
        pair_get(s2, s_even, s_odd);
        z[i] = (rhs - ((s_even + tail) + s_odd)) * dginv[i];
        if (is_vec) bvec[i] = z[i]; else B[i * LD + j] = z[i];
      }
    }
    }
  }
  cta_sync<D>();
}

// In-place lower Cholesky of the n x n block (thread i owns row i).  `dg` (>= n entries) receives
// the diagonal of L; it is also written back to the diagonal.  Non-positive pivots give NaN.
template <typename T, int D> __device__ __forceinline__ void chol_generic(T* A, T* dg, int n) {
  constexpr int LD = Tile<D>::LD;
  const int i = threadIdx.x;
  for (int k = 0; k < n; ++k) {
    cta_sync<D>();
    const T lkk = sqrt(A[k * LD + k]);
    T lik = T(0);
    if (i == k) dg[k] = lkk;
    if (i > k && i < n) {
      lik = A[i * LD + k] / lkk;
      A[i * LD + k] = lik;
    }
    cta_sync<D>();
    if (i > k && i < n) {
      for (int j = k + 1; j <= i; ++j) A[i * LD + j] = fma_(-lik, A[j * LD + k], A[i * LD + j]);
    }
  }
  cta_sync<D>();
  if (i < n) {
    A[i * LD + i] = dg[i];
    for (int j = i + 1; j < n; ++j) A[i * LD + j] = T(0);
  }
  cta_sync<D>();
}

// Z = L^{-1} B in place for the n x m block B (thread j owns column j).  `bvec` (may be null) is
// one extra right-hand-side vector of length n handled by thread m (requires NT > m).
template <typename T, int D> __device__ __forceinline__ void fwd_solve_generic(const T* L, T* B, T* bvec, int n, int m) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  const int j = threadIdx.x;
  if (j < m) {
    for (int i = 0; i < n; ++i) {
      T acc = B[i * LD + j];
      for (int k = 0; k < i; ++k) acc = fma_(-L[i * LD + k], B[k * LD + j], acc);
      B[i * LD + j] = acc / L[i * LD + i];
    }
  } else if (j == m && bvec != nullptr) {
    for (int i = 0; i < n; ++i) {
      T acc = bvec[i];
      for (int k = 0; k < i; ++k) acc = fma_(-L[i * LD + k], bvec[k], acc);
      bvec[i] = acc / L[i * LD + i];
    }
  }
  cta_sync<D>();
}


// Workspace convention: a "dg" buffer is 4*D words = [diag(L) | 1/diag(L) | 2*D column exchange].
template <int D> struct CholWs { static constexpr int WORDS = 4 * D; };

// Register-resident Cholesky for D = 64: warp 0 only, lane l owns rows l (columns 0..31 suffice: lower triangle) and
// l + 32 (all 64 columns); same column-exchange scheme as chol_reg.  The 256-thread generic version spent its time
// in 64 block barriers and 8-way bank conflicts (one thread per row, row stride 68 words).
template <typename T, bool PADDED = false>
__device__ __forceinline__ void chol_reg64(T* A, T* dg, T* dginv, T* colbuf, int n) {
  constexpr int D = 64, H = 32, LD = Tile<D>::LD, V = 4;
  cta_sync<64>();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    T r0[H], r1[D];
#pragma unroll
    for (int j = 0; j < D; j += V) {
      T t0[V], t1[V];
      if (j < H) ldv<T, V>(t0, A + lane * LD + j);
      ldv<T, V>(t1, A + (lane + H) * LD + j);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (j < H) r0[j + q] = (PADDED || lane < n) ? t0[q] : T(0);
        r1[j + q] = (PADDED || lane + H < n) ? t1[q] : T(0);
      }
    }
    // columns 0..31: pivot in row k of the first row set; both row sets are updated
#pragma unroll
    for (int k = 0; k < H; ++k) {
      if (PADDED || k < n) {
        const T akk = __shfl_sync(0xffffffffu, r0[k], k);
        T lkk, inv;
        pivot_<T>(akk, lkk, inv);
        T* cb = colbuf + (k & 1) * D;
        const T l0 = (lane == k) ? lkk : r0[k] * inv;
        const T l1 = r1[k] * inv;
        r0[k] = l0;
        r1[k] = l1;
        cb[lane] = l0;
        cb[lane + H] = l1;
        if (lane == k) { dg[k] = lkk; dginv[k] = inv; }
        __syncwarp();
        const T nl0 = -l0, nl1 = -l1;
#pragma unroll
        for (int j = ((k + 1) / V) * V; j < D; j += V) {
          T c[V];
          ldv<T, V>(c, cb + j);
#pragma unroll
          for (int q = 0; q < V; q += 2) {
            if (j + q > k) {  // both columns of the pair are open: packed FMAs
              const Pair<T> c2 = pair_make(c[q], c[q + 1]);
              if (j + q + 1 < H) {
                const Pair<T> a2 = pair_fma_s(nl0, c2, pair_make(r0[(j + q) & (H - 1)], r0[(j + q + 1) & (H - 1)]));
                pair_get(a2, r0[(j + q) & (H - 1)], r0[(j + q + 1) & (H - 1)]);
              }
              const Pair<T> b2 = pair_fma_s(nl1, c2, pair_make(r1[j + q], r1[j + q + 1]));
              pair_get(b2, r1[j + q], r1[j + q + 1]);
            } else if (j + q + 1 > k) {
              if (j + q + 1 < H) r0[(j + q + 1) & (H - 1)] = fma_(nl0, c[q + 1], r0[(j + q + 1) & (H - 1)]);
              r1[j + q + 1] = fma_(nl1, c[q + 1], r1[j + q + 1]);
            }
          }
        }
      }
    }
    // columns 32..63: pivot in row k - 32 of the second row set, which alone is still open
#pragma unroll
    for (int k = H; k < D; ++k) {
      if (PADDED || k < n) {
        const T akk = __shfl_sync(0xffffffffu, r1[k], k - H);
        T lkk, inv;
        pivot_<T>(akk, lkk, inv);
        T* cb = colbuf + (k & 1) * D;
        const T l1 = (lane == k - H) ? lkk : r1[k] * inv;
        r1[k] = l1;
        cb[lane + H] = l1;
        if (lane == k - H) { dg[k] = lkk; dginv[k] = inv; }
        __syncwarp();
        const T nl1 = -l1;
#pragma unroll
        for (int j = ((k + 1) / V) * V; j < D; j += V) {
          T c[V];
          ldv<T, V>(c, cb + j);
#pragma unroll
          for (int q = 0; q < V; q += 2) {
            if (j + q > k) {
              const Pair<T> b2 = pair_fma_s(nl1, pair_make(c[q], c[q + 1]), pair_make(r1[j + q], r1[j + q + 1]));
              pair_get(b2, r1[j + q], r1[j + q + 1]);
            } else if (j + q + 1 > k) {
              r1[j + q + 1] = fma_(nl1, c[q + 1], r1[j + q + 1]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < D; j += V) {
      T t0[V], t1[V];
#pragma unroll
      for (int q = 0; q < V; ++q) {
        t0[q] = (j < H && j + q <= lane) ? r0[j < H ? j + q : 0] : T(0);
        t1[q] = (j + q <= lane + H) ? r1[j + q] : T(0);
      }
      if (PADDED || lane < n) stv<T, V>(A + lane * LD + j, t0);
      if (PADDED || lane + H < n) stv<T, V>(A + (lane + H) * LD + j, t1);
    }
  }
  cta_sync<64>();
}

// Two-warp Cholesky for D = 64: thread r (of the item's 64) owns row r in registers and BOTH warps work on every column
// (the one-warp chol_reg64 left the item's second warp at a barrier for ~29 % of the d = 64 sampler's time).  One named
// barrier per column: the RAW column k (a_rk, pivot included) is exchanged through a double-buffered shared vector, every
// thread derives 1/l_kk from a_kk itself, and the update is written row[c] -= (a_rk / a_kk) a_ck, i.e. one scalar
// s_r = a_rk inv^2 per thread and one FMA per element against the raw column -- no second exchange of the scaled column.
// FAST: bare MUFU.RSQ pivots (fp32 structured paths, see pivot_fast_).
// FUSE_MV: the factor is not written back; thread r returns y_r = sum_{j <= r} L_rj x_j in *yreg (see chol_reg).
template <typename T, bool FAST = false, bool FUSE_MV = false>
__device__ __forceinline__ void chol_rows64(T* A, T* dg, T* dginv, T* colbuf, int n, const T* xvec = nullptr,
                                            T* yreg = nullptr) {
  constexpr int D = 64, LD = Tile<D>::LD, V = 4;
  cta_sync<64>();
  const int r = threadIdx.x;
  T row[D];
#pragma unroll
  for (int j = 0; j < D; j += V) {
    T tmp[V];
    ldv<T, V>(tmp, A + r * LD + j);
#pragma unroll
    for (int q = 0; q < V; ++q) row[j + q] = (r < n) ? tmp[q] : T(0);
  }
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (k < n) {  // uniform over the item
      T* cb = colbuf + (k & 1) * D;
      cb[r] = row[k];
      cta_sync<64>();
      const T akk = cb[k];
      T lkk, inv;
      if constexpr (FAST) pivot_fast_<T>(akk, lkk, inv); else pivot_<T>(akk, lkk, inv);
      const T lrk = (r == k) ? lkk : row[k] * inv;
      row[k] = lrk;
      if (r == 0) {
        dg[k] = lkk;
        if (dginv) dginv[k] = inv;
      }
      const T ns = -(lrk * inv);
#pragma unroll
      for (int j = ((k + 1) / V) * V; j < D; j += V) {
        T c[V];
        ldv<T, V>(c, cb + j);
#pragma unroll
        for (int q = 0; q < V; q += 2) {
          if (j + q > k) {
            const Pair<T> r2 = pair_fma_s(ns, pair_make(c[q], c[q + 1]), pair_make(row[j + q], row[j + q + 1]));
            pair_get(r2, row[j + q], row[j + q + 1]);
          } else if (j + q + 1 > k) {
            row[j + q + 1] = fma_(ns, c[q + 1], row[j + q + 1]);
          }
        }
      }
    }
  }
  if constexpr (FUSE_MV) {
    T a0 = T(0), a1 = T(0);
#pragma unroll
    for (int j = 0; j < D; j += V) {
      T xv[V];
      ldv<T, V>(xv, xvec + j);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (j + q <= r) {
          if (q & 1) a1 = fma_(row[j + q], xv[q], a1); else a0 = fma_(row[j + q], xv[q], a0);
        }
      }
    }
    *yreg = a0 + a1;
  } else if (r < n) {
#pragma unroll
    for (int j = 0; j < D; j += V) {
      T tmp[V];
#pragma unroll
      for (int q = 0; q < V; ++q) tmp[q] = (j + q <= r) ? row[j + q] : T(0);
      stv<T, V>(A + r * LD + j, tmp);
    }
  }
  cta_sync<64>();
}

// In-place lower Cholesky of the n x n block; strict upper triangle zeroed; non-PD pivots give NaN.
template <typename T, int D, bool PADDED = false, bool KEEP_UPPER = false>
__device__ __forceinline__ void chol_inplace(T* A, T* dg, int n) {
  if constexpr (D <= 32) chol_reg<T, D, PADDED, KEEP_UPPER>(A, dg, dg + D, dg + 2 * D, n);
  else if constexpr (D == 64 && Tile<D>::NT == 64) chol_rows64<T, false>(A, dg, dg + D, dg + 2 * D, n);
  else if constexpr (D == 64 && Tile<D>::NT >= 32) chol_reg64<T, PADDED>(A, dg, dg + D, dg + 2 * D, n);
  else chol_generic<T, D>(A, dg, n);
}
// Z = L^{-1} B in place (n x m block) and optionally bvec <- L^{-1} bvec; `dg` is the workspace the
// factorisation of L filled.
template <typename T, int D, bool PADDED = false>
__device__ __forceinline__ void fwd_solve(const T* L, T* B, T* bvec, int n, int m, const T* dg) {
  if constexpr (D <= 64) fwd_solve_reg<T, D, PADDED>(L, B, bvec, dg + D, n, m);  // D = 64: one thread per column
  else fwd_solve_generic<T, D>(L, B, bvec, n, m);
}

// Branch-free matrix-vector products over the whole padded tile (operands zero beyond their logical size):
// y = e (+|-) A^T x   (A k-major)  and  y = e (+) A x  (A row-major).  MODE as mm_tn.  D <= 32 (one value per thread).
template <typename T, int D, int MODE>
__device__ __forceinline__ void mv_tn_pad(const T* A, const T* x, T* y, const T* e) {
  constexpr int LD = Tile<D>::LD, V = (D >= 4) ? 4 : D;
  cta_sync<D>();
  const int i = threadIdx.x;
  T out = T(0);
  if (i < D) {
    T a0 = T(0), a1 = T(0);
#pragma unroll
    for (int k = 0; k < D; k += V) {
      T xv[V];
      ldv<T, V>(xv, x + k);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (q & 1) a1 = fma_(A[(k + q) * LD + i], xv[q], a1); else a0 = fma_(A[(k + q) * LD + i], xv[q], a0);
      }
    }
    const T acc = a0 + a1;
    out = (MODE == 0) ? acc : (MODE == 1 ? e[i] + acc : e[i] - acc);
  }
  cta_sync<D>();
  if (i < D) y[i] = out;
  cta_sync<D>();
}
template <typename T, int D, int MODE>
__device__ __forceinline__ void mv_nn_pad(const T* A, const T* x, T* y, const T* e) {
  constexpr int LD = Tile<D>::LD, V = (D >= 4) ? 4 : D;
  cta_sync<D>();
  const int i = threadIdx.x;
  T out = T(0);
  if (i < D) {
    T a0 = T(0), a1 = T(0);
#pragma unroll
    for (int k = 0; k < D; k += V) {
      T xv[V], av[V];
      ldv<T, V>(xv, x + k);
      ldv<T, V>(av, A + i * LD + k);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        if (q & 1) a1 = fma_(av[q], xv[q], a1); else a0 = fma_(av[q], xv[q], a0);
      }
    }
    const T acc = a0 + a1;
    out = (MODE == 0) ? acc : (MODE == 1 ? e[i] + acc : e[i] - acc);
  }
  cta_sync<D>();
  if (i < D) y[i] = out;
  cta_sync<D>();
}

// X = (L L^T)^{-1} B in place (n x m block, thread j owns column j).
template <typename T, int D> __device__ __forceinline__ void cho_solve_cols(const T* L, T* B, int n, int m) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  const int j = threadIdx.x;
  if (j < m) {
    for (int i = 0; i < n; ++i) {
      T acc = B[i * LD + j];
      for (int k = 0; k < i; ++k) acc = fma_(-L[i * LD + k], B[k * LD + j], acc);
      B[i * LD + j] = acc / L[i * LD + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      T acc = B[i * LD + j];
      for (int k = i + 1; k < n; ++k) acc = fma_(-L[k * LD + i], B[k * LD + j], acc);
      B[i * LD + j] = acc / L[i * LD + i];
    }
  }
  cta_sync<D>();
}

// LU with partial pivoting of the n x n block, in place (packed L\U in pivot order), thread i owns
// row i.  perm[p] = original row placed at position p.  `red` needs >= 2*NT/32 + 2 words of T and
// `ired` >= NT/32 + 1 ints (shared).  Rows are physically swapped, LAPACK style.
template <typename T, int D>
__device__ __forceinline__ void lu_inplace(T* A, int* perm, T* red, int* ired, int n) {
  constexpr int LD = Tile<D>::LD, NT = Tile<D>::NT, NW = NT / 32;
  const int i = threadIdx.x, lane = i & 31, w = i >> 5;
  if (i < n) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    cta_sync<D>();
    // argmax |A[r][k]|, r >= k
    T v = (i >= k && i < n) ? fabs(A[i * LD + k]) : T(-1);
    if (i >= k && i < n && !(v == v)) v = T(INFINITY);  // NaN pivots first, so they propagate like LAPACK
    int r = i;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, v, off);
      const int orr = __shfl_xor_sync(0xffffffffu, r, off);
      if (ov > v || (ov == v && orr < r)) { v = ov; r = orr; }
    }
    if (NW > 1) {
      if (lane == 0) { red[w] = v; ired[w] = r; }
      __syncthreads();
      v = red[0]; r = ired[0];
      for (int q = 1; q < NW; ++q) {
        if (red[q] > v || (red[q] == v && ired[q] < r)) { v = red[q]; r = ired[q]; }
      }
      __syncthreads();
    }
    const int p = r;  // pivot row (uniform)
    if (p != k) {
      for (int j = i; j < n; j += NT) {
        const T t = A[k * LD + j];
        A[k * LD + j] = A[p * LD + j];
        A[p * LD + j] = t;
      }
      if (i == 0) { const int t = perm[k]; perm[k] = perm[p]; perm[p] = t; }
    }
    cta_sync<D>();
    const T piv = A[k * LD + k];
    T l = T(0);
    if (i > k && i < n) {
      l = A[i * LD + k] / piv;
    }
    cta_sync<D>();
    if (i > k && i < n) {
      A[i * LD + k] = l;
      for (int j = k + 1; j < n; ++j) A[i * LD + j] = fma_(-l, A[k * LD + j], A[i * LD + j]);
    }
  }
  cta_sync<D>();
}

// Solve (P^T L U) X = B for the n x m block B, result written to X (X must not alias B).
// Thread j owns column j.  `bvec`/`xvec`: optional extra vector RHS handled by thread m.
template <typename T, int D>
__device__ __forceinline__ void lu_solve(const T* LU, const int* perm, const T* B, T* X, const T* bvec, T* xvec,
                                         int n, int m) {
  constexpr int LD = Tile<D>::LD;
  cta_sync<D>();
  const int j = threadIdx.x;
  if (j < m) {
    for (int i = 0; i < n; ++i) {
      T acc = B[perm[i] * LD + j];
      for (int k = 0; k < i; ++k) acc = fma_(-LU[i * LD + k], X[k * LD + j], acc);
      X[i * LD + j] = acc;
    }
    for (int i = n - 1; i >= 0; --i) {
      T acc = X[i * LD + j];
      for (int k = i + 1; k < n; ++k) acc = fma_(-LU[i * LD + k], X[k * LD + j], acc);
      X[i * LD + j] = acc / LU[i * LD + i];
    }
  }
  // the extra vector: thread m, or thread 0 after its column when the CTA has no spare thread (m >= NT)
  if (bvec != nullptr && (j == m || (m >= Tile<D>::NT && j == 0))) {
    for (int i = 0; i < n; ++i) {
      T acc = bvec[perm[i]];
      for (int k = 0; k < i; ++k) acc = fma_(-LU[i * LD + k], xvec[k], acc);
      xvec[i] = acc;
    }
    for (int i = n - 1; i >= 0; --i) {
      T acc = xvec[i];
      for (int k = i + 1; k < n; ++k) acc = fma_(-LU[i * LD + k], xvec[k], acc);
      xvec[i] = acc / LU[i * LD + i];
    }
  }
  cta_sync<D>();
}

// Block-wide sum of one value per thread; result valid in every thread.  `red` >= NT/32 words.
template <typename T, int D> __device__ __forceinline__ T cta_sum(T v, T* red) {
  constexpr int NW = Tile<D>::NT / 32;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (NW == 1) return v;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = T(0);
  for (int q = 0; q < NW; ++q) s += red[q];
  return s;
}

}  // namespace auxssm

// Backward (pathwise smoothing) sampling for linear-Gaussian state-space models.
//
// Replaces aux_samplers/_primitives/kalman/sampling.py of the reference:
//   sampling:11-40, _sampling_op_impl:52-55, mean_and_chol:61-105, _sampling_init_one:109-112,
//   _sample_last_step:116-124, _sampling_init:127-136 (the N(0, I) draws of :128 are an input).
//
// Per time step t < T-1, with (m, P) the filtering marginal at t and (F, Q, b) the transition to t+1:
//   S = F P F^T + Q = Ls Ls^T,  Zs = Ls^{-1} F P,  gain G = P F^T S^{-1} = Zs^T Ls^{-1},
//   Sigma = P - G S G^T = P - Zs^T Zs = L L^T (NaN -> 0),  e = m - G (F m + b) + L eps,
//   x_t = G x_{t+1} + e.
//
//  * parallel == 0 (or one time chunk): one CTA per unit walks t = T-1 .. 0; G is never formed,
//    x_t = m + Zs^T Ls^{-1} (x_{t+1} - F m - b) + L eps.
//  * parallel == 1 with n > 1 time chunks: a chunked REVERSE scan over the affine maps (G, e) with
//    the reference's operator (G2 G1, G2 e1 + e2):
//      pass A  the last chunk runs the recursion directly; every other chunk forms (G_t, e_t)
//              explicitly, stores them, and folds its composite map;
//      pass B  one CTA per unit chains the chunk-boundary states through the composites;
//      pass C  every chunk but the last replays x_t = G_t x_{t+1} + e_t from its boundary.
#include "common.cuh"
#include "smem_linalg.cuh"

namespace auxssm {

template <typename T> struct SampleArgs {
  auxssm_lgssm_desc d;
  const T *ms, *Ps, *Fs, *Qs, *bs, *eps;
  T* xs;
  T* steps;     // [units][T-1][dx*dx + dx]   (G_t^T, e_t)       (chunked only)
  T* comp;      // [units][n_chunks][dx*dx + dx] composite maps   (chunked only)
  T* boundary;  // [units][n_chunks][dx]  x at the first time of chunk j (chunked only)
  int n_chunks, chunk_len;
};

template <typename T, int D> struct SCarve {
  T* p;
  __device__ T* mat() { T* r = p; p += D * Tile<D>::LD; return r; }
  __device__ T* vec() { T* r = p; p += D; return r; }
};

template <typename T, int D> struct SampBuf {
  T *P, *W, *S, *Z, *Ft, *Q, *Gagg, *Gt;
  T *m, *bv, *eps, *x, *r, *dg, *e, *eagg, *red;
};

// The direct recursion needs 6 matrix buffers; the chunk-aggregate pass one more (the composite map).
template <int D> __host__ __device__ constexpr size_t samp_smem_elems(bool aggregate) {
  return (size_t)(aggregate ? 7 : 6) * D * Tile<D>::LD + 12 * D + 16;
}

template <typename T, int D> __device__ __forceinline__ SampBuf<T, D> carve_samp(SCarve<T, D>& cv, bool aggregate) {
  SampBuf<T, D> s;
  s.P = cv.mat(); s.W = cv.mat(); s.S = cv.mat(); s.Z = cv.mat(); s.Ft = cv.mat(); s.Q = cv.mat();
  s.Gagg = aggregate ? cv.mat() : nullptr;
  s.Gt = nullptr;
  s.m = cv.vec(); s.bv = cv.vec(); s.eps = cv.vec(); s.x = cv.vec(); s.r = cv.vec();
  s.e = cv.vec(); s.eagg = cv.vec(); s.red = cv.vec();
  s.dg = cv.p; cv.p += CholWs<D>::WORDS;
  return s;
}

// chunk j covers times [j*len, min((j+1)*len, T-1)) for j < n-1 ... the LAST chunk also owns t = T-1.
__device__ __forceinline__ void samp_chunk_range(int j, int n_chunks, int chunk_len, int T, int& t0, int& t1) {
  t0 = j * chunk_len;
  t1 = (j + 1) * chunk_len;
  if (j == n_chunks - 1 || t1 > T) t1 = T;
}

// In-place lower Cholesky with the reference's nan_to_num semantics: jnp.linalg.cholesky returns an
// all-NaN factor when it fails, which nan_to_num turns into the zero matrix (sampling.py:102-104).
template <typename T, int D> __device__ __forceinline__ void chol_nan_to_zero(T* A, T* dg, int n) {
  constexpr int LD = Tile<D>::LD, NT = Tile<D>::NT;
  chol_inplace<T, D>(A, dg, n);
  bool bad = false;
  for (int i = 0; i < n; ++i) bad = bad || !(dg[i] > T(0)) || !finite_(dg[i]);
  if (n == 1) bad = !finite_(dg[0]);  // scalar branch of the reference is a plain sqrt
  if (bad) {
    for (int e = threadIdx.x; e < n * n; e += NT) A[(e / n) * LD + (e % n)] = T(0);
  }
  cta_sync<D>();
}

template <typename T, int D>
__device__ __forceinline__ void load_step_inputs(const SampleArgs<T>& a, int c, int b, int t, bool first_dyn,
                                                 SampBuf<T, D>& s, bool with_dyn) {
  const int dx = a.d.dx;
  cta_sync<D>();
  if (threadIdx.x < dx) {
    s.m[threadIdx.x] = a.ms[unit_off(a.d.ms, c, b) + (int64_t)t * a.d.ms.time + threadIdx.x];
    s.eps[threadIdx.x] = a.eps[unit_off(a.d.eps, c, b) + (int64_t)t * a.d.eps.time + threadIdx.x];
  }
  load_mat<T, D, false>(a.Ps + unit_off(a.d.Ps, c, b) + (int64_t)t * a.d.Ps.time, s.P, dx, dx);
  if (with_dyn) {
    if (first_dyn || a.d.Fs.time != 0)
      load_mat<T, D, true>(a.Fs + unit_off(a.d.Fs, c, b) + (int64_t)t * a.d.Fs.time, s.Ft, dx, dx);
    if (first_dyn || a.d.Qs.time != 0)
      load_mat<T, D, false>(a.Qs + unit_off(a.d.Qs, c, b) + (int64_t)t * a.d.Qs.time, s.Q, dx, dx);
    if (first_dyn || a.d.bs.time != 0) {
      if (threadIdx.x < dx) s.bv[threadIdx.x] = a.bs[unit_off(a.d.bs, c, b) + (int64_t)t * a.d.bs.time + threadIdx.x];
    }
  }
  cta_sync<D>();
}

// x <- m + chol(P) eps  (last step, sampling.py:116-124)
template <typename T, int D> __device__ __forceinline__ void last_step(SampBuf<T, D>& s, int dx) {
  chol_nan_to_zero<T, D>(s.P, s.dg, dx);
  mv_nn<T, D, 1>(s.P, s.eps, s.x, s.m, dx, dx);
}

template <typename T, int D>
__device__ __forceinline__ void store_x(const SampleArgs<T>& a, int c, int b, int t, const T* x) {
  if (threadIdx.x < a.d.dx) a.xs[unit_off(a.d.xs, c, b) + (int64_t)t * a.d.xs.time + threadIdx.x] = x[threadIdx.x];
}

// pass A, last chunk (and the whole job when n_chunks == 1): direct recursion.
template <typename T, int D>
__global__ void __launch_bounds__(Tile<D>::NT, Occ<T, D>::MINB) sample_recursion_kernel(SampleArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SCarve<T, D> cv{reinterpret_cast<T*>(smem_raw)};
  constexpr int NT = Tile<D>::NT, LD = Tile<D>::LD;
  for (size_t e = threadIdx.x; e < samp_smem_elems<D>(false); e += NT) cv.p[e] = T(0);
  __syncthreads();
  SampBuf<T, D> s = carve_samp<T, D>(cv, false);
  const int unit = blockIdx.x;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx, Tn = a.d.T;
  int t0, t1;
  samp_chunk_range(a.n_chunks - 1, a.n_chunks, a.chunk_len, Tn, t0, t1);
  load_step_inputs<T, D>(a, c, b, Tn - 1, true, s, false);
  last_step<T, D>(s, dx);
  store_x<T, D>(a, c, b, Tn - 1, s.x);
  bool first = true;
  int fdiag_known = 0;
  for (int t = Tn - 2; t >= t0; --t) {
    load_step_inputs<T, D>(a, c, b, t, first, s, true);
    first = false;
    if (fdiag_known == 0) {  // is F time-invariant and diagonal?  (decided once per CTA)
      int off = (a.d.Fs.time != 0) ? 1 : 0;
      AUXSSM_FOR_ROWS(ri, rj, dx, dx) if (ri != rj && s.Ft[ri * LD + rj] != T(0)) off = 1;
      fdiag_known = (__syncthreads_or(off) == 0) ? 1 : -1;
    }
    if (fdiag_known > 0) {
      // diagonal F: S_ab = f_a f_b P_ab + Q_ab, (F P)_ab = f_a P_ab, r = x_{t+1} - f m - b  (O(d^2))
      AUXSSM_FOR_ROWS(ri, rj, dx, dx) {
        const T fa = s.Ft[ri * LD + ri], pv = s.P[ri * LD + rj];
        s.S[ri * LD + rj] = fma_(fa * s.Ft[rj * LD + rj], pv, s.Q[ri * LD + rj]);
        s.Z[ri * LD + rj] = fa * pv;
      }
      if (threadIdx.x < dx) {
        const int ri = threadIdx.x;
        s.r[ri] = s.x[ri] - fma_(s.Ft[ri * LD + ri], s.m[ri], s.bv[ri]);
      }
      cta_sync<D>();
    } else {
      mm_tn<T, D, 0>(s.P, s.Ft, s.W, nullptr, dx);   // W = P F^T
      mm_tn<T, D, 1>(s.Ft, s.W, s.S, s.Q, dx);       // S = F P F^T + Q
      symmetrize<T, D>(s.S, dx);
      mm_tn<T, D, 0>(s.Ft, s.P, s.Z, nullptr, dx);   // Z = F P
      mv_tn<T, D, 1>(s.Ft, s.m, s.r, s.bv, dx, dx);  // r = F m + b
      if (threadIdx.x < dx) s.r[threadIdx.x] = s.x[threadIdx.x] - s.r[threadIdx.x];  // x_{t+1} - F m - b
    }
    chol_inplace<T, D>(s.S, s.dg, dx);
    fwd_solve<T, D>(s.S, s.Z, s.r, dx, dx, s.dg);        // Zs = Ls^{-1} F P ; r <- Ls^{-1} r
    mm_tn<T, D, 2>(s.Z, s.Z, s.P, s.P, dx);        // Sigma = P - Zs^T Zs
    symmetrize<T, D>(s.P, dx);
    chol_nan_to_zero<T, D>(s.P, s.dg, dx);         // L = chol(Sigma), NaN -> 0
    mv_tn<T, D, 1>(s.Z, s.r, s.x, s.m, dx, dx);    // x = m + Zs^T Ls^{-1} (x_{t+1} - F m - b)
    mv_nn<T, D, 1>(s.P, s.eps, s.x, s.x, dx, dx);  // x += L eps
    store_x<T, D>(a, c, b, t, s.x);
  }
  if (a.n_chunks > 1 && threadIdx.x < dx)
    a.boundary[((int64_t)unit * a.n_chunks + (a.n_chunks - 1)) * dx + threadIdx.x] = s.x[threadIdx.x];
  (void)LD;
}

// pass A, chunks 0..n-2: explicit (G_t, e_t), stored; composite folded from the chunk's last step down.
template <typename T, int D>
__global__ void __launch_bounds__(Tile<D>::NT, Occ<T, D>::MINB) sample_aggregate_kernel(SampleArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SCarve<T, D> cv{reinterpret_cast<T*>(smem_raw)};
  constexpr int NT = Tile<D>::NT, LD = Tile<D>::LD;
  for (size_t e = threadIdx.x; e < samp_smem_elems<D>(true); e += NT) cv.p[e] = T(0);
  __syncthreads();
  SampBuf<T, D> s = carve_samp<T, D>(cv, true);
  const int unit = blockIdx.x, j = blockIdx.y;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  const int dx = a.d.dx, Tn = a.d.T;
  const int ssz = dx * dx + dx;
  int t0, t1;
  samp_chunk_range(j, a.n_chunks, a.chunk_len, Tn, t0, t1);
  bool first = true;
  for (int t = t1 - 1; t >= t0; --t) {
    load_step_inputs<T, D>(a, c, b, t, first, s, true);
    mm_tn<T, D, 0>(s.P, s.Ft, s.W, nullptr, dx);
    mm_tn<T, D, 1>(s.Ft, s.W, s.S, s.Q, dx);
    symmetrize<T, D>(s.S, dx);
    mm_tn<T, D, 0>(s.Ft, s.P, s.Z, nullptr, dx);
    mv_tn<T, D, 1>(s.Ft, s.m, s.r, s.bv, dx, dx);  // r = F m + b
    chol_inplace<T, D>(s.S, s.dg, dx);
    fwd_solve<T, D>(s.S, s.Z, nullptr, dx, dx, s.dg);    // Zs
    mm_tn<T, D, 2>(s.Z, s.Z, s.P, s.P, dx);        // Sigma
    symmetrize<T, D>(s.P, dx);
    // G^T = Ls^{-T} Zs : back substitution, thread q owns column q of Z
    cta_sync<D>();
    if (threadIdx.x < dx) {
      const int q = threadIdx.x;
      for (int i = dx - 1; i >= 0; --i) {
        T acc = s.Z[i * LD + q];
        for (int k = i + 1; k < dx; ++k) acc = fma_(-s.S[k * LD + i], s.Z[k * LD + q], acc);
        s.Z[i * LD + q] = acc / s.S[i * LD + i];
      }
    }
    cta_sync<D>();
    chol_nan_to_zero<T, D>(s.P, s.dg, dx);         // L
    mv_tn<T, D, 2>(s.Z, s.r, s.e, s.m, dx, dx);    // e = m - G (F m + b)      (G v = (G^T)^T v)
    mv_nn<T, D, 1>(s.P, s.eps, s.e, s.e, dx, dx);  // e += L eps
    // store (G^T, e)
    T* g = a.steps + ((int64_t)unit * (Tn - 1) + t) * ssz;
    store_mat<T, D>(g, s.Z, dx, dx);
    if (threadIdx.x < dx) g[dx * dx + threadIdx.x] = s.e[threadIdx.x];
    // fold: composite <- (G_t, e_t) o composite
    if (first) {
      transpose_to<T, D>(s.Z, s.Gagg, dx, dx);     // Gagg = G_t (plain)
      if (threadIdx.x < dx) s.eagg[threadIdx.x] = s.e[threadIdx.x];
      cta_sync<D>();
    } else {
      mm_tn<T, D, 0>(s.Z, s.Gagg, s.W, nullptr, dx);    // W = G_t Gagg
      mv_tn<T, D, 1>(s.Z, s.eagg, s.eagg, s.e, dx, dx); // eagg = G_t eagg + e_t
      copy_mat<T, D>(s.W, s.Gagg, dx, dx);
    }
    first = false;
  }
  cta_sync<D>();
  T* g = a.comp + ((int64_t)unit * a.n_chunks + j) * ssz;
  store_mat<T, D>(g, s.Gagg, dx, dx);
  if (threadIdx.x < dx) g[dx * dx + threadIdx.x] = s.eagg[threadIdx.x];
}

// pass B: boundary_j = Gagg_j boundary_{j+1} + eagg_j, j = n-2 .. 0.  One warp per unit.
template <typename T> __global__ void sample_boundary_kernel(SampleArgs<T> a) {
  const int unit = blockIdx.x, dx = a.d.dx, ssz = dx * dx + dx;
  __shared__ T xb[AUXSSM_MAX_DIM];
  T* bnd = a.boundary + (int64_t)unit * a.n_chunks * dx;
  if (threadIdx.x < dx) xb[threadIdx.x] = bnd[(int64_t)(a.n_chunks - 1) * dx + threadIdx.x];
  __syncthreads();
  for (int j = a.n_chunks - 2; j >= 0; --j) {
    const T* g = a.comp + ((int64_t)unit * a.n_chunks + j) * ssz;
    T acc = T(0);
    if (threadIdx.x < dx) {
      acc = g[dx * dx + threadIdx.x];
      for (int k = 0; k < dx; ++k) acc = fma_(g[threadIdx.x * dx + k], xb[k], acc);
    }
    __syncthreads();
    if (threadIdx.x < dx) { xb[threadIdx.x] = acc; bnd[(int64_t)j * dx + threadIdx.x] = acc; }
    __syncthreads();
  }
}

// pass C: x_t = G_t x_{t+1} + e_t inside chunks 0..n-2, starting from boundary_{j+1}.
template <typename T> __global__ void sample_replay_kernel(SampleArgs<T> a) {
  const int unit = blockIdx.x, j = blockIdx.y, dx = a.d.dx, ssz = dx * dx + dx, Tn = a.d.T;
  const int c = unit / a.d.n_batch, b = unit % a.d.n_batch;
  __shared__ T xb[AUXSSM_MAX_DIM];
  int t0, t1;
  samp_chunk_range(j, a.n_chunks, a.chunk_len, Tn, t0, t1);
  if (threadIdx.x < dx) xb[threadIdx.x] = a.boundary[((int64_t)unit * a.n_chunks + j + 1) * dx + threadIdx.x];
  __syncthreads();
  for (int t = t1 - 1; t >= t0; --t) {
    const T* g = a.steps + ((int64_t)unit * (Tn - 1) + t) * ssz;  // G^T then e
    T acc = T(0);
    if (threadIdx.x < dx) {
      acc = g[dx * dx + threadIdx.x];
      for (int k = 0; k < dx; ++k) acc = fma_(g[k * dx + threadIdx.x], xb[k], acc);
    }
    __syncthreads();
    if (threadIdx.x < dx) {
      xb[threadIdx.x] = acc;
      a.xs[unit_off(a.d.xs, c, b) + (int64_t)t * a.d.xs.time + threadIdx.x] = acc;
    }
    __syncthreads();
  }
}

static size_t s_align_up(size_t x, size_t al) { return (x + al - 1) / al * al; }

struct SamplePlan {
  int n_chunks, chunk_len;
  size_t off_steps, off_comp, off_boundary, total;
};

static SamplePlan make_sample_plan(const auxssm_lgssm_desc* d, size_t es) {
  SamplePlan p;
  p.n_chunks = resolve_time_chunks(d);
  // chunks partition times 0..T-2 (the last chunk additionally owns T-1); at least 1 step per chunk
  const int n_step = d->T > 1 ? d->T - 1 : 1;
  if (p.n_chunks > n_step) p.n_chunks = n_step;
  p.chunk_len = (n_step + p.n_chunks - 1) / p.n_chunks;
  p.n_chunks = (n_step + p.chunk_len - 1) / p.chunk_len;
  const size_t units = (size_t)d->n_chains * d->n_batch;
  const size_t ssz = (size_t)d->dx * d->dx + d->dx;
  size_t off = 0;
  p.off_steps = off;
  if (p.n_chunks > 1) off = s_align_up(off + units * (size_t)(d->T - 1) * ssz * es, 256);
  p.off_comp = off;
  if (p.n_chunks > 1) off = s_align_up(off + units * p.n_chunks * ssz * es, 256);
  p.off_boundary = off;
  if (p.n_chunks > 1) off = s_align_up(off + units * p.n_chunks * d->dx * es, 256);
  p.total = off;
  return p;
}

template <typename T, int D> static int launch_sample(cudaStream_t st, const SampleArgs<T>& a) {
  const int units = a.d.n_chains * a.d.n_batch;
  constexpr int NT = Tile<D>::NT;
  const size_t smem = samp_smem_elems<D>(false) * sizeof(T);
  int rc = ensure_smem(sample_recursion_kernel<T, D>, smem, "auxssm_kalman_sample");
  if (rc) return rc;
  if (a.n_chunks > 1) {
    rc = ensure_smem(sample_aggregate_kernel<T, D>, samp_smem_elems<D>(true) * sizeof(T),
                     "auxssm_kalman_sample (parallel)");
    if (rc) return rc;
  }
  sample_recursion_kernel<T, D><<<units, NT, smem, st>>>(a);
  AUXSSM_CHECK_LAUNCH("auxssm_kalman_sample (recursion)");
  if (a.n_chunks > 1) {
    const size_t smem_agg = samp_smem_elems<D>(true) * sizeof(T);
    sample_aggregate_kernel<T, D><<<dim3(units, a.n_chunks - 1), NT, smem_agg, st>>>(a);
    AUXSSM_CHECK_LAUNCH("auxssm_kalman_sample (aggregate)");
    sample_boundary_kernel<T><<<units, 64, 0, st>>>(a);
    AUXSSM_CHECK_LAUNCH("auxssm_kalman_sample (boundary)");
    sample_replay_kernel<T><<<dim3(units, a.n_chunks - 1), 64, 0, st>>>(a);
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
  const SamplePlan p = make_sample_plan(d, sizeof(T));
  if (scratch_bytes < p.total || (p.total > 0 && scratch == nullptr)) {
    set_error("%s: scratch too small (%zu < %zu bytes)", who, scratch_bytes, p.total);
    return AUXSSM_ERR_SCRATCH;
  }
  SampleArgs<T> a;
  a.d = *d;
  a.ms = ms; a.Ps = Ps; a.Fs = Fs; a.Qs = Qs; a.bs = bs; a.eps = eps; a.xs = xs;
  unsigned char* sb = static_cast<unsigned char*>(scratch);
  a.steps = reinterpret_cast<T*>(sb + p.off_steps);
  a.comp = reinterpret_cast<T*>(sb + p.off_comp);
  a.boundary = reinterpret_cast<T*>(sb + p.off_boundary);
  a.n_chunks = p.n_chunks;
  a.chunk_len = p.chunk_len;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (padded_dim(d->dx)) {
    case 4: return launch_sample<T, 4>(st, a);
    case 8: return launch_sample<T, 8>(st, a);
    case 16: return launch_sample<T, 16>(st, a);
    case 32: return launch_sample<T, 32>(st, a);
    case 64: return launch_sample<T, 64>(st, a);
    default: set_error("%s: unsupported dimension %d", who, d->dx); return AUXSSM_ERR_UNSUPPORTED;
  }
}

}  // namespace auxssm

extern "C" __attribute__((visibility("default"))) size_t auxssm_kalman_sample_scratch_bytes(const auxssm_lgssm_desc* desc, int32_t elem_size) {
  if (auxssm::check_desc(desc, "auxssm_kalman_sample_scratch_bytes")) return 0;
  return auxssm::make_sample_plan(desc, (size_t)elem_size).total;
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_sample_f32(void* stream, const auxssm_lgssm_desc* desc, const float* ms,
                                            const float* Ps, const float* Fs, const float* Qs, const float* bs,
                                            const float* eps, float* xs, void* scratch, size_t scratch_bytes) {
  return auxssm::sample_entry<float>(stream, desc, ms, Ps, Fs, Qs, bs, eps, xs, scratch, scratch_bytes);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_sample_f64(void* stream, const auxssm_lgssm_desc* desc, const double* ms,
                                            const double* Ps, const double* Fs, const double* Qs, const double* bs,
                                            const double* eps, double* xs, void* scratch, size_t scratch_bytes) {
  return auxssm::sample_entry<double>(stream, desc, ms, Ps, Fs, Qs, bs, eps, xs, scratch, scratch_bytes);
}

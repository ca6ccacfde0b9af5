// Entry points of the conditional-SMC path: the fused sweep (csmc_kernel.cuh), the guided-proposal
// gains pre-pass, and the unfused per-step primitives (normalize / conditional resampling / categorical)
// for models outside the fused kernel's closed family set.
#include "csmc_launch.cuh"

namespace auxssm {

// the three proposal families are instantiated in csmc_{bootstrap,independent,guided}.cu
extern template int launch_csmc_family<float, AUXSSM_PROP_BOOTSTRAP>(cudaStream_t, CsmcArgs<float>&, int, int*);
extern template int launch_csmc_family<double, AUXSSM_PROP_BOOTSTRAP>(cudaStream_t, CsmcArgs<double>&, int, int*);
extern template int launch_csmc_family<float, AUXSSM_PROP_INDEPENDENT>(cudaStream_t, CsmcArgs<float>&, int, int*);
extern template int launch_csmc_family<double, AUXSSM_PROP_INDEPENDENT>(cudaStream_t, CsmcArgs<double>&, int, int*);
extern template int launch_csmc_family<float, AUXSSM_PROP_GUIDED>(cudaStream_t, CsmcArgs<float>&, int, int*);
extern template int launch_csmc_family<double, AUXSSM_PROP_GUIDED>(cudaStream_t, CsmcArgs<double>&, int, int*);

// ---- planning --------------------------------------------------------------------------------------

static size_t c_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct CsmcPlan {
  int slots;
  size_t off_xs, off_lw, off_a, total;
};

template <typename T> static int plan_slots(const auxssm_csmc_desc* d) {
  CsmcArgs<T> a{};
  a.d = *d;
  a.hist_per_chain = 1;  // so that the planning call is not limited by max_slots
  int slots = 1, rc = AUXSSM_OK;
  switch (d->proposal) {
    case AUXSSM_PROP_BOOTSTRAP: rc = launch_csmc_family<T, AUXSSM_PROP_BOOTSTRAP>((cudaStream_t)-1, a, 0, &slots); break;
    case AUXSSM_PROP_INDEPENDENT: rc = launch_csmc_family<T, AUXSSM_PROP_INDEPENDENT>((cudaStream_t)-1, a, 0, &slots); break;
    case AUXSSM_PROP_GUIDED: rc = launch_csmc_family<T, AUXSSM_PROP_GUIDED>((cudaStream_t)-1, a, 0, &slots); break;
    default: rc = AUXSSM_ERR_INVALID;
  }
  if (rc != AUXSSM_OK || slots < 1) {
    // no device (CPU-side size query) or failure: assume two CTAs per SM
    slots = 2 * device_sm_count();
    if (slots > d->n_chains) slots = d->n_chains;
  }
  return slots;
}

static CsmcPlan make_csmc_plan(const auxssm_csmc_desc* d, size_t es, int slots) {
  CsmcPlan p;
  p.slots = slots;
  size_t off = 0;
  p.off_xs = off; off = c_align_up(off + (size_t)slots * d->T * d->N * d->d * es, 256);
  p.off_lw = off; off = c_align_up(off + (size_t)slots * d->T * d->N * es, 256);
  p.off_a = off;  off = c_align_up(off + (size_t)slots * (d->T > 1 ? d->T - 1 : 0) * d->N * sizeof(int32_t), 256);
  p.total = off;
  return p;
}

static int check_csmc_desc(const auxssm_csmc_desc* d, const char* who) {
  if (!d) { set_error("%s: null descriptor", who); return AUXSSM_ERR_INVALID; }
  if (d->T < 1 || d->N < 1 || d->d < 1 || d->n_chains < 1) {
    set_error("%s: T, N, d, n_chains must be >= 1", who);
    return AUXSSM_ERR_INVALID;
  }
  if (d->N > 1024) { set_error("%s: N = %d exceeds 1024 particles", who, d->N); return AUXSSM_ERR_UNSUPPORTED; }
  if (d->d > 32) { set_error("%s: d = %d exceeds 32", who, d->d); return AUXSSM_ERR_UNSUPPORTED; }
  if (d->proposal < 0 || d->proposal > 2 || d->potential < 0 || d->potential > 2 || d->resampling < 0 ||
      d->resampling > 2) {
    set_error("%s: unknown proposal / potential / resampling id", who);
    return AUXSSM_ERR_INVALID;
  }
  return AUXSSM_OK;
}

template <typename T>
static int csmc_entry(void* stream, const auxssm_csmc_desc* d, const T* x_star, const T* m0, const T* Lp0, const T* F,
                      const T* b, const T* Lq, const T* ys, const T* u, const T* shd, const T* grad, const T* K,
                      const T* cL, const T* normals, const T* u_res, const T* u_kill, const T* u_bwd, T* x_out,
                      int32_t* anc, T* xs_hist, T* lw_hist, int32_t* a_hist, void* scratch, size_t scratch_bytes) {
  const char* who = "auxssm_csmc_sweep";
  int rc = check_csmc_desc(d, who);
  if (rc) return rc;
  const bool rng_on = !normals && !u_res && !u_kill && !u_bwd;  // in-kernel Philox draws
  if (!x_star || !m0 || !Lp0 || !F || !b || !Lq || !x_out || !anc ||
      (!rng_on && (!normals || !u_bwd || (d->T > 1 && !u_res)))) {
    set_error("%s: null data pointer (pass all draw arrays, or none of them for in-kernel draws)", who);
    return AUXSSM_ERR_INVALID;
  }
  if (d->potential != AUXSSM_POT_FLAT && !ys) { set_error("%s: potential needs ys", who); return AUXSSM_ERR_INVALID; }
  if (d->proposal != AUXSSM_PROP_BOOTSTRAP && (!u || !shd)) {
    set_error("%s: auxiliary proposals need u and sqrt_half_delta", who);
    return AUXSSM_ERR_INVALID;
  }
  if (d->proposal == AUXSSM_PROP_GUIDED && (!K || !cL)) {
    set_error("%s: guided proposal needs K and chol_Lam (auxssm_guided_gains)", who);
    return AUXSSM_ERR_INVALID;
  }
  if (!rng_on && d->resampling == AUXSSM_RESAMPLE_KILLING && d->T > 1 && !u_kill) {
    set_error("%s: killing resampling needs u_kill", who);
    return AUXSSM_ERR_INVALID;
  }
  if (d->dyn_time_varying && (d->proposal == AUXSSM_PROP_GUIDED || d->gradient != 0)) {
    set_error("%s: time-varying dynamics are supported for BOOTSTRAP and INDEPENDENT (gradient = 0) proposals", who);
    return AUXSSM_ERR_UNSUPPORTED;
  }
  const bool full_hist = xs_hist && lw_hist && a_hist;
  if (!full_hist && (xs_hist || lw_hist || a_hist)) {
    set_error("%s: pass all three history outputs or none", who);
    return AUXSSM_ERR_INVALID;
  }
  CsmcArgs<T> a;
  a.d = *d;
  a.x_star = x_star; a.m0 = m0; a.Lp0 = Lp0; a.F = F; a.b = b; a.Lq = Lq; a.ys = ys; a.u = u; a.shd = shd;
  a.grad = grad; a.K = K; a.cL = cL; a.normals = normals; a.u_res = u_res; a.u_kill = u_kill; a.u_bwd = u_bwd;
  a.x_out = x_out; a.anc = anc;
  a.rng.on = rng_on ? 1 : 0;
  a.rng.seed = d->rng_seed;
  a.rng.call0 = d->rng_counter;
  a.rng_counter = rng_on ? d->rng_counter_ptr : nullptr;
  int max_slots = 0;
  if (full_hist) {
    a.xs_hist = xs_hist; a.lw_hist = lw_hist; a.a_hist = a_hist; a.hist_per_chain = 1;
  } else {
    // scratch holds as many history slots as fit; the persistent grid is clamped to that
    const CsmcPlan one = make_csmc_plan(d, sizeof(T), 1);
    max_slots = one.total ? (int)(scratch_bytes / one.total) : 0;
    if (max_slots > d->n_chains) max_slots = d->n_chains;
    if (max_slots < 1 || scratch == nullptr) {
      set_error("%s: scratch too small for one history slot (%zu < %zu bytes)", who, scratch_bytes, one.total);
      return AUXSSM_ERR_SCRATCH;
    }
    const CsmcPlan p = make_csmc_plan(d, sizeof(T), max_slots);
    // make_csmc_plan aligns each region, so shrink until it really fits
    CsmcPlan q = p;
    while (q.total > scratch_bytes && max_slots > 1) q = make_csmc_plan(d, sizeof(T), --max_slots);
    if (q.total > scratch_bytes) { set_error("%s: scratch too small", who); return AUXSSM_ERR_SCRATCH; }
    unsigned char* sb = static_cast<unsigned char*>(scratch);
    a.xs_hist = reinterpret_cast<T*>(sb + q.off_xs);
    a.lw_hist = reinterpret_cast<T*>(sb + q.off_lw);
    a.a_hist = reinterpret_cast<int32_t*>(sb + q.off_a);
    a.hist_per_chain = 0;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d->proposal) {
    case AUXSSM_PROP_BOOTSTRAP: return launch_csmc_family<T, AUXSSM_PROP_BOOTSTRAP>(st, a, max_slots, nullptr);
    case AUXSSM_PROP_INDEPENDENT: return launch_csmc_family<T, AUXSSM_PROP_INDEPENDENT>(st, a, max_slots, nullptr);
    default: return launch_csmc_family<T, AUXSSM_PROP_GUIDED>(st, a, max_slots, nullptr);
  }
}

// ---- guided gains ------------------------------------------------------------------------------------

template <typename T, int D>
__global__ void __launch_bounds__(Tile<D>::NT) guided_gains_kernel(int Tn, int d, const T* __restrict__ P0,
                                                                    const T* __restrict__ Q,
                                                                    const T* __restrict__ shd, T* __restrict__ K,
                                                                    T* __restrict__ cL) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* p = reinterpret_cast<T*>(smem_raw);
  constexpr int NT = Tile<D>::NT, LD = Tile<D>::LD;
  for (int e = threadIdx.x; e < 4 * D * LD + CholWs<D>::WORDS; e += NT) p[e] = T(0);
  __syncthreads();
  T* P = p; p += D * LD;
  T* S = p; p += D * LD;
  T* Kt = p; p += D * LD;   // K^T = S^{-1} P
  T* Lam = p; p += D * LD;
  T* dg = p;
  const int t = blockIdx.x;
  const T s = shd[t];
  load_mat<T, D, false>(t == 0 ? P0 : Q, P, d, d);
  cta_sync<D>();
  AUXSSM_FOR_ROWS(i, j, d, d) {
    const T v = P[i * LD + j];
    S[i * LD + j] = v + ((i == j) ? s * s : T(0));
    Kt[i * LD + j] = v;
  }
  chol_inplace<T, D>(S, dg, d);
  cho_solve_cols<T, D>(S, Kt, d, d);                      // Kt = (P + s^2 I)^{-1} P = K^T
  mm_tn<T, D, 2>(Kt, P, Lam, P, d);                       // Lam = P - K P   (K P = (K^T)^T P)
  symmetrize<T, D>(Lam, d);
  chol_inplace<T, D>(Lam, dg, d);
  cta_sync<D>();
  T* Kg = K + (int64_t)t * d * d;
  T* Lg = cL + (int64_t)t * d * d;
  AUXSSM_FOR_ROWS(i, j, d, d) {
    Kg[i * d + j] = Kt[j * LD + i];
    T l = Lam[i * LD + j];
    if (!finite_(l)) l = (i == j) ? s : T(0);             // where(isfinite(chol), chol, scale * eye)
    Lg[i * d + j] = l;
  }
}

template <typename T, int D>
static int launch_gains(cudaStream_t st, int Tn, int d, const T* P0, const T* Q, const T* shd, T* K, T* cL) {
  const size_t smem = ((size_t)4 * D * Tile<D>::LD + CholWs<D>::WORDS) * sizeof(T);
  int rc = ensure_smem(guided_gains_kernel<T, D>, smem, "auxssm_guided_gains");
  if (rc) return rc;
  guided_gains_kernel<T, D><<<Tn, Tile<D>::NT, smem, st>>>(Tn, d, P0, Q, shd, K, cL);
  AUXSSM_CHECK_LAUNCH("auxssm_guided_gains");
  return AUXSSM_OK;
}

template <typename T>
static int gains_entry(void* stream, int32_t Tn, int32_t d, const T* P0, const T* Q, const T* shd, T* K, T* cL) {
  if (Tn < 1 || d < 1 || !P0 || !Q || !shd || !K || !cL) {
    set_error("auxssm_guided_gains: invalid argument");
    return AUXSSM_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (padded_dim(d)) {
    case 4: return launch_gains<T, 4>(st, Tn, d, P0, Q, shd, K, cL);
    case 8: return launch_gains<T, 8>(st, Tn, d, P0, Q, shd, K, cL);
    case 16: return launch_gains<T, 16>(st, Tn, d, P0, Q, shd, K, cL);
    case 32: return launch_gains<T, 32>(st, Tn, d, P0, Q, shd, K, cL);
    default: set_error("auxssm_guided_gains: d = %d exceeds 32", d); return AUXSSM_ERR_UNSUPPORTED;
  }
}

// ---- unfused primitives -------------------------------------------------------------------------------

enum { PRIM_NORMALIZE = 0, PRIM_MULTINOMIAL = 1, PRIM_KILLING = 2, PRIM_CATEGORICAL = 3 };

template <typename T, int OP>
__global__ void __launch_bounds__(1024) csmc_primitive_kernel(int N, const T* __restrict__ in, const T* __restrict__ u,
                                                              const T* __restrict__ u_kill, T* __restrict__ wout,
                                                              int32_t* __restrict__ idx) {
  __shared__ T red[64];
  __shared__ T cdf[1024];
  const int64_t row = blockIdx.x;
  const int i = threadIdx.x;
  const bool active = i < N;
  if (OP == PRIM_NORMALIZE) {
    const T lw = active ? in[row * N + i] : T(0);
    const T w = block_normalize<T>(lw, active, red);
    if (active) wout[row * N + i] = w;
    return;
  }
  const T w = active ? in[row * N + i] : T(0);
  const T cd = block_cdf<T>(w, red);
  if (active) cdf[i] = cd;
  __syncthreads();
  if (OP == PRIM_CATEGORICAL) {
    const T r = cdf[N - 1] * (T(1) - u[row]);
    int B = __syncthreads_count(active && cdf[i] < r);
    if (B > N - 1) B = N - 1;
    if (i == 0) idx[row] = B;
    return;
  }
  T wmax = T(0);
  if (OP == PRIM_KILLING) wmax = block_max<T>(w, red);
  if (active) {
    const T r = cdf[N - 1] * (T(1) - u[row * N + i]);
    int a = search_left<T>(cdf, r, N);
    if (OP == PRIM_KILLING && u_kill[row * N + i] * wmax < w) a = i;
    if (i == 0) a = 0;
    idx[row * N + i] = a;
  }
}

// Conditional systematic resampling of M weights into N indices (resamplings.py:40-86); u3 = (U, V, W) per row.
template <typename T>
__global__ void __launch_bounds__(1024) csmc_systematic_kernel(int M, int N, const T* __restrict__ w_in,
                                                               const T* __restrict__ u3, int32_t* __restrict__ idx) {
  __shared__ T red[64];
  __shared__ T cdf[1024];
  const int64_t row = blockIdx.x;
  const int i = threadIdx.x;
  const T w = (i < M) ? w_in[row * M + i] : T(0);
  const T cd = block_cdf<T>(w, red);
  if (i < M) cdf[i] = cd;
  __syncthreads();
  if (i < N) {
    const SysPlan<T> pl = systematic_plan<T>(cdf, N, u3[row * 3], u3[row * 3 + 1], u3[row * 3 + 2]);
    idx[row * N + i] = search_left<T>(cdf, pl.r(i, N), M);
  }
}

template <typename T>
static int systematic_entry(void* stream, int64_t n_rows, int32_t M, int32_t N, const T* w, const T* u3, int32_t* idx) {
  const char* who = "auxssm_resample_systematic";
  if (n_rows < 1 || M < 1 || N < 1 || !w || !u3 || !idx) { set_error("%s: invalid argument", who); return AUXSSM_ERR_INVALID; }
  if (M > 1024 || N > 1024) { set_error("%s: M = %d / N = %d exceed 1024", who, M, N); return AUXSSM_ERR_UNSUPPORTED; }
  const int n = M > N ? M : N, ntb = (n + 31) / 32 * 32;
  csmc_systematic_kernel<T><<<(unsigned)n_rows, ntb, 0, static_cast<cudaStream_t>(stream)>>>(M, N, w, u3, idx);
  AUXSSM_CHECK_LAUNCH(who);
  return AUXSSM_OK;
}

template <typename T, int OP>
static int primitive_entry(void* stream, int64_t n_rows, int32_t N, const T* in, const T* u, const T* u_kill, T* wout,
                           int32_t* idx, const char* who) {
  if (n_rows < 1 || N < 1 || !in || (OP == PRIM_NORMALIZE ? !wout : (!u || !idx)) || (OP == PRIM_KILLING && !u_kill)) {
    set_error("%s: invalid argument", who);
    return AUXSSM_ERR_INVALID;
  }
  if (N > 1024) { set_error("%s: N = %d exceeds 1024", who, N); return AUXSSM_ERR_UNSUPPORTED; }
  const int ntb = (N + 31) / 32 * 32;
  csmc_primitive_kernel<T, OP><<<(unsigned)n_rows, ntb, 0, static_cast<cudaStream_t>(stream)>>>(N, in, u, u_kill, wout,
                                                                                              idx);
  AUXSSM_CHECK_LAUNCH(who);
  return AUXSSM_OK;
}

}  // namespace auxssm

#define AUXSSM_API extern "C" __attribute__((visibility("default")))

AUXSSM_API size_t auxssm_csmc_scratch_bytes(const auxssm_csmc_desc* desc, int32_t elem_size) {
  if (auxssm::check_csmc_desc(desc, "auxssm_csmc_scratch_bytes")) return 0;
  const int slots = elem_size == 8 ? auxssm::plan_slots<double>(desc) : auxssm::plan_slots<float>(desc);
  return auxssm::make_csmc_plan(desc, (size_t)elem_size, slots).total;
}

AUXSSM_API int32_t auxssm_csmc_sweep_f32(void* stream, const auxssm_csmc_desc* desc, const float* x_star,
                                         const float* m0, const float* chol_P0, const float* F, const float* b,
                                         const float* chol_Q, const float* ys, const float* u,
                                         const float* sqrt_half_delta, const float* grad, const float* K,
                                         const float* chol_Lam, const float* normals, const float* u_resample,
                                         const float* u_kill, const float* u_backward, float* x_out,
                                         int32_t* ancestors, float* xs_hist, float* logw_hist, int32_t* anc_hist,
                                         void* scratch, size_t scratch_bytes) {
  return auxssm::csmc_entry<float>(stream, desc, x_star, m0, chol_P0, F, b, chol_Q, ys, u, sqrt_half_delta, grad, K,
                                   chol_Lam, normals, u_resample, u_kill, u_backward, x_out, ancestors, xs_hist,
                                   logw_hist, anc_hist, scratch, scratch_bytes);
}
AUXSSM_API int32_t auxssm_csmc_sweep_f64(void* stream, const auxssm_csmc_desc* desc, const double* x_star,
                                         const double* m0, const double* chol_P0, const double* F, const double* b,
                                         const double* chol_Q, const double* ys, const double* u,
                                         const double* sqrt_half_delta, const double* grad, const double* K,
                                         const double* chol_Lam, const double* normals, const double* u_resample,
                                         const double* u_kill, const double* u_backward, double* x_out,
                                         int32_t* ancestors, double* xs_hist, double* logw_hist, int32_t* anc_hist,
                                         void* scratch, size_t scratch_bytes) {
  return auxssm::csmc_entry<double>(stream, desc, x_star, m0, chol_P0, F, b, chol_Q, ys, u, sqrt_half_delta, grad, K,
                                    chol_Lam, normals, u_resample, u_kill, u_backward, x_out, ancestors, xs_hist,
                                    logw_hist, anc_hist, scratch, scratch_bytes);
}
AUXSSM_API int32_t auxssm_guided_gains_f32(void* stream, int32_t T, int32_t d, const float* P0, const float* Q,
                                           const float* sqrt_half_delta, float* K, float* chol_Lam) {
  return auxssm::gains_entry<float>(stream, T, d, P0, Q, sqrt_half_delta, K, chol_Lam);
}
AUXSSM_API int32_t auxssm_guided_gains_f64(void* stream, int32_t T, int32_t d, const double* P0, const double* Q,
                                           const double* sqrt_half_delta, double* K, double* chol_Lam) {
  return auxssm::gains_entry<double>(stream, T, d, P0, Q, sqrt_half_delta, K, chol_Lam);
}

#define AUXSSM_PRIMS(SFX, TY)                                                                                        \
  AUXSSM_API int32_t auxssm_normalize_##SFX(void* stream, int64_t n_rows, int32_t N, const TY* log_w, TY* w) {        \
    return auxssm::primitive_entry<TY, auxssm::PRIM_NORMALIZE>(stream, n_rows, N, log_w, nullptr, nullptr, w, nullptr, \
                                                               "auxssm_normalize");                                  \
  }                                                                                                                  \
  AUXSSM_API int32_t auxssm_resample_multinomial_##SFX(void* stream, int64_t n_rows, int32_t N, const TY* w,          \
                                                       const TY* u, int32_t* idx) {                                  \
    return auxssm::primitive_entry<TY, auxssm::PRIM_MULTINOMIAL>(stream, n_rows, N, w, u, nullptr, nullptr, idx,      \
                                                                 "auxssm_resample_multinomial");                     \
  }                                                                                                                  \
  AUXSSM_API int32_t auxssm_resample_killing_##SFX(void* stream, int64_t n_rows, int32_t N, const TY* w,              \
                                                   const TY* u_kill, const TY* u, int32_t* idx) {                    \
    return auxssm::primitive_entry<TY, auxssm::PRIM_KILLING>(stream, n_rows, N, w, u, u_kill, nullptr, idx,           \
                                                             "auxssm_resample_killing");                             \
  }                                                                                                                  \
  AUXSSM_API int32_t auxssm_categorical_##SFX(void* stream, int64_t n_rows, int32_t N, const TY* w, const TY* u,      \
                                              int32_t* idx) {                                                        \
    return auxssm::primitive_entry<TY, auxssm::PRIM_CATEGORICAL>(stream, n_rows, N, w, u, nullptr, nullptr, idx,      \
                                                                 "auxssm_categorical");                              \
  }
AUXSSM_PRIMS(f32, float)
AUXSSM_PRIMS(f64, double)

extern "C" __attribute__((visibility("default"))) int32_t auxssm_resample_systematic_f32(
    void* stream, int64_t n_rows, int32_t M, int32_t N, const float* w, const float* u3, int32_t* idx) {
  return auxssm::systematic_entry<float>(stream, n_rows, M, N, w, u3, idx);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_resample_systematic_f64(
    void* stream, int64_t n_rows, int32_t M, int32_t N, const double* w, const double* u3, int32_t* idx) {
  return auxssm::systematic_entry<double>(stream, n_rows, M, N, w, u3, idx);
}

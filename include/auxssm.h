/*
 * auxssm.h -- C ABI of libauxssm.so: B200 (sm_100a) kernels for the two hot paths of
 * aux-ssm-samplers (auxiliary Kalman sampler; conditional SMC particle Gibbs).
 *
 * The reference is pure Python/JAX and has no FFI of its own; each entry point below
 * replaces one *Python* seam of the reference (cited as file:line relative to the
 * reference root) and is what a `jax.ffi` custom call for that seam would bind
 * (see INTEGRATION.md for the binding stubs).
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every data pointer is CALLER-OWNED DEVICE memory;
 *    descriptors are host memory, read during the call only.
 *  - Every entry point enqueues work on `stream` (a cudaStream_t passed as void*) and
 *    returns without synchronising.  No allocation, no host<->device copies.
 *  - Return value: 0 on success, a negative AUXSSM_ERR_* otherwise; the message is
 *    available (per host thread) from auxssm_last_error().
 *  - Numerical failure (non-PD Cholesky) never errors: it produces NaN/0 the way the
 *    reference does (filtering.py:116, sampling.py:104).
 *  - Strides are in ELEMENTS.  A stride of 0 broadcasts along that axis (e.g. a
 *    time-invariant F is passed once with Fs.time = 0 instead of being tiled as in
 *    examples/stochastic_volatility/auxiliary_kalman.py:23-25).
 *  - `_f32` / `_f64` suffixes select the arithmetic type; there is no mixed precision.
 */
#ifndef AUXSSM_H
#define AUXSSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUXSSM_VERSION 100

enum {
  AUXSSM_OK = 0,
  AUXSSM_ERR_INVALID = -1,     /* bad descriptor / null pointer / unsupported size   */
  AUXSSM_ERR_SCRATCH = -2,     /* scratch buffer too small                           */
  AUXSSM_ERR_LAUNCH = -3,      /* CUDA reported an error at launch                   */
  AUXSSM_ERR_UNSUPPORTED = -4  /* valid request this build cannot serve              */
};

#define AUXSSM_MAX_DIM 64      /* max(dx, dy) supported by the kernels               */

/* Address of element (chain c, inner batch b, time t) = base + c*chain + b*batch + t*time
 * (+ row-major trailing core axes, always contiguous).                               */
typedef struct auxssm_stride {
  int64_t chain;
  int64_t batch;
  int64_t time;
} auxssm_stride;

/* flags */
#define AUXSSM_FLAG_H_IDENTITY 1u /* Hs == NULL: H_t = I (requires dy == dx), c_t = 0 if cs == NULL */
#define AUXSSM_FLAG_R_DIAGONAL 2u /* Rs holds diag(R_t): core shape [dy] instead of [dy, dy]          */

/*
 * Linear-Gaussian state-space model, mirrors the LGSSM NamedTuple
 * (_primitives/kalman/base.py:12-69):
 *   x_0 ~ N(m0, P0);  x_t = F_{t-1} x_{t-1} + b_{t-1} + N(0, Q_{t-1});  y_t = H_t x_t + c_t + N(0, R_t)
 * Core shapes: m0[dx] P0[dx,dx] Fs,Qs[T-1][dx,dx] bs[T-1][dx] Hs[T][dy,dx] Rs[T][dy,dy] cs[T][dy] ys[T][dy].
 * `n_chains` is the vmap-leading chain axis, `n_batch` the LGSSM inner batch of
 * base.py:40-49; both are independent units for the kernels.
 */
typedef struct auxssm_lgssm_desc {
  int32_t T, dx, dy;
  int32_t n_chains, n_batch;
  int32_t parallel;     /* 0: sequential in time (filtering.py:66-79); 1: parallel in time (:49-63) */
  int32_t time_chunks;  /* parallel only: number of time chunks; 0 = choose automatically           */
  uint32_t flags;
  auxssm_stride ys, m0, P0, Fs, Qs, bs, Hs, Rs, cs; /* inputs                               */
  auxssm_stride ms, Ps;                             /* filtering marginals [.., T, dx(,dx)]  */
  auxssm_stride eps, xs;                            /* N(0,I) draws and trajectories [.., T, dx] */
} auxssm_lgssm_desc;

const char* auxssm_last_error(void);
int32_t auxssm_version(void);
/* cumulative number of CUDA kernels this library has launched in the calling process */
int64_t auxssm_launch_count(void);

/* ---- Kalman filter: replaces filtering(ys, lgssm, parallel), filtering.py:18-46 ------------
 * Outputs ms, Ps (strides desc->ms, desc->Ps) and ell[n_chains] (summed over the inner batch,
 * filtering.py:43-45).  Missing observations are non-finite entries of ys (filtering.py:86-130). */
size_t auxssm_kalman_filter_scratch_bytes(const auxssm_lgssm_desc* desc, int32_t elem_size);
int32_t auxssm_kalman_filter_f32(void* stream, const auxssm_lgssm_desc* desc, const float* ys,
                                 const float* m0, const float* P0, const float* Fs, const float* Qs,
                                 const float* bs, const float* Hs, const float* Rs, const float* cs,
                                 float* ms, float* Ps, float* ell, void* scratch, size_t scratch_bytes);
int32_t auxssm_kalman_filter_f64(void* stream, const auxssm_lgssm_desc* desc, const double* ys,
                                 const double* m0, const double* P0, const double* Fs, const double* Qs,
                                 const double* bs, const double* Hs, const double* Rs, const double* cs,
                                 double* ms, double* Ps, double* ell, void* scratch, size_t scratch_bytes);

/* ---- Backward (pathwise smoothing) sampler: replaces sampling(key, ms, Ps, lgssm, parallel),
 * sampling.py:11-40.  `eps` are the N(0,I) draws the reference makes at sampling.py:128
 * (the key -> normal step stays with the caller so draws can be supplied externally).        */
size_t auxssm_kalman_sample_scratch_bytes(const auxssm_lgssm_desc* desc, int32_t elem_size);
int32_t auxssm_kalman_sample_f32(void* stream, const auxssm_lgssm_desc* desc, const float* ms,
                                 const float* Ps, const float* Fs, const float* Qs, const float* bs,
                                 const float* eps, float* xs, void* scratch, size_t scratch_bytes);
int32_t auxssm_kalman_sample_f64(void* stream, const auxssm_lgssm_desc* desc, const double* ms,
                                 const double* Ps, const double* Fs, const double* Qs, const double* bs,
                                 const double* eps, double* xs, void* scratch, size_t scratch_bytes);

/* ---- Trajectory log-densities: replaces posterior_logpdf / prior_logpdf / log_likelihood,
 * base.py:73-166.  out[n_chains] = (which&LOGLIK) log p(ys|xs) - (which&MINUS_ELL) ell + (which&PRIOR) log p(xs),
 * nansum semantics, summed over the inner batch.                                              */
#define AUXSSM_LOGPDF_PRIOR 1u
#define AUXSSM_LOGPDF_LOGLIK 2u
#define AUXSSM_LOGPDF_MINUS_ELL 4u
#define AUXSSM_LOGPDF_POSTERIOR 7u
size_t auxssm_lgssm_logpdf_scratch_bytes(const auxssm_lgssm_desc* desc, int32_t elem_size);
int32_t auxssm_lgssm_logpdf_f32(void* stream, const auxssm_lgssm_desc* desc, uint32_t which,
                                const float* ys, const float* xs, const float* ell, const float* m0,
                                const float* P0, const float* Fs, const float* Qs, const float* bs,
                                const float* Hs, const float* Rs, const float* cs, float* out,
                                void* scratch, size_t scratch_bytes);
int32_t auxssm_lgssm_logpdf_f64(void* stream, const auxssm_lgssm_desc* desc, uint32_t which,
                                const double* ys, const double* xs, const double* ell, const double* m0,
                                const double* P0, const double* Fs, const double* Qs, const double* bs,
                                const double* Hs, const double* Rs, const double* cs, double* out,
                                void* scratch, size_t scratch_bytes);

/* ---- Pieces of the auxiliary Kalman MH step, kalman/generic.py:53-106 ------------------------
 * auxiliary draw  u = x + sqrt(delta/2) * normals                     (generic.py:61)
 * accept/reject   alpha = exp(min(0, lt_prop - lt_rev + lp_rev - lp_prop - sum((x'-u)^2 - (x-u)^2)/delta));
 *                 accept = uniform < alpha; x <- accept ? x_prop : x   (generic.py:70-74, 98-106)
 * Arrays are [n_chains, n_per_chain] contiguous; delta is one value per chain (delta_stride 1)
 * or shared (delta_stride 0).                                                                   */
int32_t auxssm_auxiliary_draw_f32(void* stream, int64_t n_chains, int64_t n_per_chain, const float* x,
                                  const float* normals, const float* delta, int32_t delta_stride, float* u);
int32_t auxssm_auxiliary_draw_f64(void* stream, int64_t n_chains, int64_t n_per_chain, const double* x,
                                  const double* normals, const double* delta, int32_t delta_stride, double* u);
int32_t auxssm_kalman_mh_f32(void* stream, int64_t n_chains, int64_t n_per_chain, const float* delta,
                             int32_t delta_stride, const float* log_lgssm_prop, const float* log_lgssm_rev,
                             const float* log_target_prop, const float* log_target_rev, const float* u,
                             const float* x_prop, const float* uniforms, float* x, int32_t* accepted,
                             float* alpha);
int32_t auxssm_kalman_mh_f64(void* stream, int64_t n_chains, int64_t n_per_chain, const double* delta,
                             int32_t delta_stride, const double* log_lgssm_prop, const double* log_lgssm_rev,
                             const double* log_target_prop, const double* log_target_rev, const double* u,
                             const double* x_prop, const double* uniforms, double* x, int32_t* accepted,
                             double* alpha);

/* ---- Stochastic-volatility proposal construction ---------------------------
 * Replaces examples/stochastic_volatility/auxiliary_kalman.py:28-48 (first / second order
 * observations factories) and the potential part of log_likelihood_fn (:50-54, model.py:56-82).
 * x, u, aux_ys are [n_chains, T, d] contiguous; ys is the data [T, d] (ys_chain_stride 0) or one
 * data set per chain (ys_chain_stride T*d).  order 1: aux_ys = u + (delta/2) nan_to_num(grad),
 * Rs[n_chains, d, d] = (delta/2) I (use Rs.time = 0).  order 2: Omega = (-hess + 2I/delta)^{-1},
 * aux_ys = Omega (2u/delta + grad - hess x), Rs[n_chains, T, d, d] = Omega (dense).
 * dense_R = 0 writes only the diagonals (Rs[n_chains, 1 or T, d], for AUXSSM_FLAG_R_DIAGONAL).   */
int32_t auxssm_sv_obs_factory_f32(void* stream, int32_t order, int64_t n_chains, int64_t T, int32_t d,
                                  const float* ys, int64_t ys_chain_stride, const float* x, const float* u,
                                  const float* delta, int32_t delta_stride, float* aux_ys, float* Rs,
                                  int32_t dense_R);
int32_t auxssm_sv_obs_factory_f64(void* stream, int32_t order, int64_t n_chains, int64_t T, int32_t d,
                                  const double* ys, int64_t ys_chain_stride, const double* x, const double* u,
                                  const double* delta, int32_t delta_stride, double* aux_ys, double* Rs,
                                  int32_t dense_R);
/* out[c] (+)= sum_{t,i} nan_to_num(log N(ys[t,i]; 0, exp(x[c,t,i]/2)))   (model.py:56-76)          */
int32_t auxssm_sv_log_potential_f32(void* stream, int64_t n_chains, int64_t n_per_chain, const float* ys,
                                    int64_t ys_chain_stride, const float* x, float* out, int32_t accumulate);
int32_t auxssm_sv_log_potential_f64(void* stream, int64_t n_chains, int64_t n_per_chain, const double* ys,
                                    int64_t ys_chain_stride, const double* x, double* out, int32_t accumulate);

#ifdef __cplusplus
}
#endif
#endif /* AUXSSM_H */

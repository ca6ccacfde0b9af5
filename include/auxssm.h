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

#define AUXSSM_VERSION 200

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
/* HINT: every F_t is probably diagonal (layout unchanged, [dx, dx]).  The library verifies it ON THE DEVICE on
 * every call (no host synchronisation) and takes the structured kernels only when it holds; with a dense F the
 * general kernels run and the results are the same as without the flag. */
#define AUXSSM_FLAG_F_IS_DIAGONAL 4u

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
/* ';'-separated names of the kernels launched by the calling host thread since the last reset (diagnostics and
 * tests: which code path served a call).  A structured kernel and the general kernel launched behind it under the
 * diagonal-F gate both appear; exactly one of them does the work. */
const char* auxssm_dispatch_log(int32_t reset);

/* ---- Kalman filter: replaces filtering(ys, lgssm, parallel), filtering.py:18-46 ------------
 * Outputs ms, Ps (strides desc->ms, desc->Ps) and ell[n_chains] (summed over the inner batch,
 * filtering.py:43-45).  Missing observations are non-finite entries of ys (filtering.py:86-130).
 * ms == Ps == NULL is an ELL-ONLY call: the marginals are not stored (the reverse proposal of
 * kalman/generic.py:67 uses the likelihood only).  The scratch buffer is always required (>= 256 bytes). */
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

/* ---- Conditional SMC particle Gibbs sweep: replaces _primitives/csmc/csmc.py:16-149 ---------------
 * (get_kernel/kernel :16-66, _csmc :69-107, _backward_scanning_pass :110-124,
 * _backward_sampling_pass :127-149), the conditional resamplers of resamplings.py:14-37, normalize
 * (math/utils.py:24-39) and the auxiliary wrappers csmc/generic.py:56-72, csmc/independent.py:144-268,
 * examples/stochastic_volatility/auxiliary_guided_csmc.py:29-167.
 *
 * The reference calls user Python callables (M0.sample, G0, Mt.sample, Gt, Pt.logpdf) inside the time
 * loop; a fused kernel needs a closed set of model families instead:
 *   prior      x_0 ~ N(m0, P0),  x_t ~ N(F x_{t-1} + b, Q)   (time-invariant, chol factors supplied)
 *   potential  FLAT 0 | GAUSSIAN sum_i log N(y_ti; x_i, sig) | SV sum_i log N(y_ti; 0, exp(x_i/2))
 *   proposal   BOOTSTRAP    propose from the prior, weight = potential
 *              INDEPENDENT  N(u_t + s_t^2 grad_t, s_t^2 I), weight = log prior transition + potential
 *                           (+ log N(x;u,s) - log N(x;mean,s) when `gradient`)      independent.py:144-268
 *                           gradient = 1 is the reference to the letter: per-particle correction at t = 0
 *                           (GradientAuxiliaryG0 :173-190); for t >= 1 GradientAuxiliaryGt (:252-268) sums the
 *                           correction over all particles, a per-step constant that cancels in every normalisation
 *                           and is therefore not added.  gradient = 2 corrects every particle at every step.
 *              GUIDED       N(mu_t, Lambda_t), mu_t = xpred + K_t (u_t - xpred); K_t, chol Lambda_t from
 *                           auxssm_guided_gains_*                        auxiliary_guided_csmc.py:90-167
 * Anything else must use the unfused primitives below, one call per time step.
 *
 * Draws are inputs (parity mode): normals[n_chains,T,N,d] (csmc.py:74,91), u_resample[n_chains,T-1,N]
 * (resamplings.py:35; SYSTEMATIC: [n_chains,T-1,3] = the (U, V, W) of resamplings.py:65),
 * u_kill[n_chains,T-1,N] (killing only), u_backward[n_chains,T] (csmc.py:111 /
 * :131,138).  Categorical draws follow jax.random.choice: idx = first j with cdf[j] >= cdf[N-1](1-U),
 * with the CDF accumulated in the library's canonical order.
 * Outputs: x_out[n_chains,T,d], ancestors[n_chains,T] (updated = ancestors != 0, csmc.py:59).
 * Optional (may be NULL) full history: xs_hist[n_chains,T,N,d], logw_hist[n_chains,T,N],
 * anc_hist[n_chains,T-1,N].  N <= 1024, d <= 32.                                                      */
enum { AUXSSM_PROP_BOOTSTRAP = 0, AUXSSM_PROP_INDEPENDENT = 1, AUXSSM_PROP_GUIDED = 2 };
enum { AUXSSM_POT_FLAT = 0, AUXSSM_POT_GAUSSIAN = 1, AUXSSM_POT_SV = 2 };
enum { AUXSSM_RESAMPLE_MULTINOMIAL = 0, AUXSSM_RESAMPLE_KILLING = 1, AUXSSM_RESAMPLE_SYSTEMATIC = 2 };

typedef struct auxssm_csmc_desc {
  int32_t T, N, d, n_chains;
  int32_t proposal, potential, resampling;
  int32_t backward;         /* 0: ancestor tracing; 1: backward sampling with Pt = the prior transition */
  int32_t gradient;         /* INDEPENDENT only: 0 none, 1 reference semantics, 2 per-particle at every step (above) */
  int64_t ys_chain_stride;  /* 0: one data set ys[T,d] shared by all chains; T*d: one per chain          */
  int64_t gains_time_stride;/* GUIDED: d*d if K / cholLam are [T,d,d], 0 if time-invariant [d,d]         */
  double sig;               /* GAUSSIAN potential: observation std                                       */
  /* In-kernel draws.  When normals, u_resample, u_kill and u_backward are ALL NULL the sweep generates them itself
   * with counter-based Philox4x32-10 (csrc/philox.cuh): the logical arrays documented above are reproduced element
   * for element as auxssm_philox_{normal,uniform}_* would fill them with (rng_seed, call), call = rng_counter + 0
   * (normals), + 1 (u_resample), + 2 (u_kill, killing only), then u_backward.  rng_counter_ptr (device memory, may
   * be NULL) overrides rng_counter with the value stored there when the kernel runs (CUDA-graph replays; advance it
   * with auxssm_philox_advance).  No draw ever touches HBM: 2N(d+1)s fewer bytes per (time step x chain).        */
  uint64_t rng_seed;
  uint64_t rng_counter;
  const uint64_t* rng_counter_ptr;
  /* Time-varying prior dynamics -- the reference scans the `Mt.params` pytree over t (csmc.py:103).  When
   * dyn_time_varying != 0: F is [T-1,d,d], b is [T-1,d], chol_Q is [T-1,d,d]; entry t-1 is the transition INTO time
   * step t.  BOOTSTRAP / INDEPENDENT (gradient = 0) proposals; served by the one-thread-per-particle kernel.       */
  int32_t dyn_time_varying;
  /* Backward sampling with Pt != Mt (csmc.py:47-50, :127-149): a linear-Gaussian Pt given by device pointers of the
   * call's dtype, laid out like F, b, chol_Q (time-varying together with them).  All NULL: Pt is Mt.               */
  const void* pt_F;
  const void* pt_b;
  const void* pt_chol_Q;
} auxssm_csmc_desc;

size_t auxssm_csmc_scratch_bytes(const auxssm_csmc_desc* desc, int32_t elem_size);
int32_t auxssm_csmc_sweep_f32(void* stream, const auxssm_csmc_desc* desc, const float* x_star,
                              const float* m0, const float* chol_P0, const float* F, const float* b,
                              const float* chol_Q, const float* ys, const float* u, const float* sqrt_half_delta,
                              const float* grad, const float* K, const float* chol_Lam, const float* normals,
                              const float* u_resample, const float* u_kill, const float* u_backward,
                              float* x_out, int32_t* ancestors, float* xs_hist, float* logw_hist,
                              int32_t* anc_hist, void* scratch, size_t scratch_bytes);
int32_t auxssm_csmc_sweep_f64(void* stream, const auxssm_csmc_desc* desc, const double* x_star,
                              const double* m0, const double* chol_P0, const double* F, const double* b,
                              const double* chol_Q, const double* ys, const double* u,
                              const double* sqrt_half_delta, const double* grad, const double* K,
                              const double* chol_Lam, const double* normals, const double* u_resample,
                              const double* u_kill, const double* u_backward, double* x_out, int32_t* ancestors,
                              double* xs_hist, double* logw_hist, int32_t* anc_hist, void* scratch,
                              size_t scratch_bytes);

/* ---- Parallel-in-time conditional SMC: replaces csmc/independent.py:78-118 (get_independent_kernel(...,
 * parallel=True)) and _primitives/csmc/pit/{csmc.py:68-114, operator.py:39-149, dc_map.py:73-123} for the closed
 * family above with INDEPENDENT proposals  x_t^n ~ N(u_t + s_t^2 grad_t, s_t^2 I)  (grad used when desc->gradient,
 * which also switches the initial weights to log N(x; u_t, s_t) - log N(x; mean_t, s_t), pit/csmc.py:83-85).
 * Uses desc->{T, N, d, n_chains, potential, gradient, sig, ys_chain_stride, rng_*}; N <= 128, d <= 32, T >= 2.
 * Draws (parity mode): normals[n_chains,T,N,d] (pit/csmc.py:75) and uniforms[n_chains,T,N]: row t serves the stitch
 * whose right-hand block starts at time t (keys_b[0], operator.py:78,81; the last level draws ONE pair from
 * uniforms[.,t,0]).  Both NULL: generated in the kernels with Philox (call = rng_counter + 0 / + 1).
 * log2(T) + 1 launches: every boundary (t-1, t) is stitched once, N^2 weights in shared memory, canonical
 * log-sum-exp / CDF order, blocks of trajectories are index rows (`origins`, operator.py:97-99), never
 * copied.  Outputs x_out[n_chains,T,d], ancestors[n_chains,T] (updated = ancestors != 0, pit/csmc.py:58).        */
size_t auxssm_pit_csmc_scratch_bytes(const auxssm_csmc_desc* desc, int32_t elem_size);
int32_t auxssm_pit_csmc_sweep_f32(void* stream, const auxssm_csmc_desc* desc, const float* x_star, const float* m0,
                                  const float* chol_P0, const float* F, const float* b, const float* chol_Q,
                                  const float* ys, const float* u, const float* sqrt_half_delta, const float* grad,
                                  const float* normals, const float* uniforms, float* x_out, int32_t* ancestors,
                                  void* scratch, size_t scratch_bytes);
int32_t auxssm_pit_csmc_sweep_f64(void* stream, const auxssm_csmc_desc* desc, const double* x_star, const double* m0,
                                  const double* chol_P0, const double* F, const double* b, const double* chol_Q,
                                  const double* ys, const double* u, const double* sqrt_half_delta,
                                  const double* grad, const double* normals, const double* uniforms, double* x_out,
                                  int32_t* ancestors, void* scratch, size_t scratch_bytes);

/* Gains of the guided proposal for every time step (auxiliary_guided_csmc.py:155-157, :44-46, :104-106):
 * K_t = P (P + s_t^2 I)^{-1}, chol_Lam_t = chol(sym(P - K_t P)) with non-finite entries replaced by
 * s_t I; P = P0 for t = 0 and Q otherwise.  sqrt_half_delta[T]; outputs [T,d,d].                       */
int32_t auxssm_guided_gains_f32(void* stream, int32_t T, int32_t d, const float* P0, const float* Q,
                                const float* sqrt_half_delta, float* K, float* chol_Lam);
int32_t auxssm_guided_gains_f64(void* stream, int32_t T, int32_t d, const double* P0, const double* Q,
                                const double* sqrt_half_delta, double* K, double* chol_Lam);

/* ---- Unfused cSMC primitives (for models outside the closed set), rows are independent ------------
 * normalize      w = exp(lw - logsumexp(lw))                                   math/utils.py:24-39
 * multinomial    idx[r, :] conditional multinomial, idx[r, 0] = 0              resamplings.py:14-37
 * killing        idx = i if u_kill * max w < w_i else multinomial draw; idx[r, 0] = 0  (not in the reference)
 * categorical    one draw per row: jax.random.choice(key, N, p=w, shape=())    csmc.py:111,131,138
 * systematic     conditional systematic resampling (Chopin & Singh Alg. 4) of M weights into N indices from the
 *                three uniforms u3[r] = (U, V, W) of resamplings.py:65; idx[r, 0] = 0       resamplings.py:40-86
 * Arrays are [n_rows, N] contiguous, N <= 1024.                                                       */
int32_t auxssm_normalize_f32(void* stream, int64_t n_rows, int32_t N, const float* log_w, float* w);
int32_t auxssm_normalize_f64(void* stream, int64_t n_rows, int32_t N, const double* log_w, double* w);
int32_t auxssm_resample_multinomial_f32(void* stream, int64_t n_rows, int32_t N, const float* w,
                                        const float* u, int32_t* idx);
int32_t auxssm_resample_multinomial_f64(void* stream, int64_t n_rows, int32_t N, const double* w,
                                        const double* u, int32_t* idx);
int32_t auxssm_resample_killing_f32(void* stream, int64_t n_rows, int32_t N, const float* w,
                                    const float* u_kill, const float* u, int32_t* idx);
int32_t auxssm_resample_killing_f64(void* stream, int64_t n_rows, int32_t N, const double* w,
                                    const double* u_kill, const double* u, int32_t* idx);
int32_t auxssm_resample_systematic_f32(void* stream, int64_t n_rows, int32_t M, int32_t N, const float* w,
                                       const float* u3, int32_t* idx);
int32_t auxssm_resample_systematic_f64(void* stream, int64_t n_rows, int32_t M, int32_t N, const double* w,
                                       const double* u3, int32_t* idx);
int32_t auxssm_categorical_f32(void* stream, int64_t n_rows, int32_t N, const float* w, const float* u,
                               int32_t* idx);
int32_t auxssm_categorical_f64(void* stream, int64_t n_rows, int32_t N, const double* w, const double* u,
                               int32_t* idx);

/* ---- Fused auxiliary-Kalman MH sweep: replaces kernel(key, state, delta) of kalman/generic.py:53-106 in ONE kernel
 * launch for a latent-Gaussian target  pi(x) ~ LGSSM prior (m0, P0, F_t, Q_t, b_t; strides in desc->lg, time-varying
 * or stride 0) x prod_t exp(g(x_t; ys_t)),  g = FLAT | GAUSSIAN (sig) | SV  (AUXSSM_POT_*), with the first / second
 * order observation factories of examples/stochastic_volatility/auxiliary_kalman.py:28-48.  Narrow states only
 * (d <= 4; AUXSSM_ERR_UNSUPPORTED otherwise: compose the entry points above).  One thread per time step; the
 * filter's associative scan (filtering.py:49-63), the sampler's reverse scan (sampling.py:28-32) and the reverse
 * proposal's filter are single-pass scans: warp shuffles inside a warp, a warp of warp aggregates inside a CTA,
 * decoupled look-back across the CTAs a long chain is split over (all CTAs must be co-resident).
 * x, normals_aux, normals_eps: [n_chains, T, d]; uniforms: [n_chains]; delta: [n_chains] (delta_stride 1) or [1];
 * outputs x_out [n_chains, T, d], accepted [n_chains], alpha [n_chains] (may be NULL).                          */
typedef struct auxssm_sweep_desc {
  auxssm_lgssm_desc lg;     /* T, dx, n_chains and the strides of m0, P0, Fs, Qs, bs are used                       */
  int32_t order;            /* 1 or 2                                                                             */
  int32_t potential;        /* AUXSSM_POT_*                                                                       */
  int32_t delta_stride;     /* 0 or 1                                                                             */
  double sig;               /* GAUSSIAN potential                                                                 */
  int64_t ys_chain_stride;  /* 0: ys[T,d] shared; T*d: per chain                                                  */
} auxssm_sweep_desc;
size_t auxssm_kalman_sweep_scratch_bytes(const auxssm_sweep_desc* desc, int32_t elem_size);
int32_t auxssm_kalman_sweep_f32(void* stream, const auxssm_sweep_desc* desc, const float* x, const float* m0,
                                const float* P0, const float* Fs, const float* Qs, const float* bs, const float* ys,
                                const float* normals_aux, const float* normals_eps, const float* uniforms,
                                const float* delta, float* x_out, int32_t* accepted, float* alpha, void* scratch,
                                size_t scratch_bytes);
int32_t auxssm_kalman_sweep_f64(void* stream, const auxssm_sweep_desc* desc, const double* x, const double* m0,
                                const double* P0, const double* Fs, const double* Qs, const double* bs,
                                const double* ys, const double* normals_aux, const double* normals_eps,
                                const double* uniforms, const double* delta, double* x_out, int32_t* accepted,
                                double* alpha, void* scratch, size_t scratch_bytes);

/* ---- Host-side pieces of the samplers as kernels (no eager tensor arithmetic on the sweep path) ------------------
 * grad_log_target  gradient at u[n_chains,T,d] of log M0(u_0) + G0(u_0) + sum_t [Gt(u_t) + log Mt(u_t | u_{t-1})]
 *                  for the closed family set -- what the reference obtains with jax.grad(_log_pdf),
 *                  csmc/independent.py:62,121-137.  Uses desc->{T,d,n_chains,potential,sig,ys_chain_stride}.
 * cholesky         n lower Cholesky factors [n,d,d] with jnp.linalg.cholesky semantics (all-NaN factor on failure;
 *                  d = 1 is a plain sqrt), the factorisations of the model constructors (auxiliary_csmc.py:18-19).
 * delta_adaptation delta * exp(adaptation_rate (acc - target)), clipped to [min, max]            common.py:4-32
 *                  (acc_stride 0: one acceptance rate shared by all n entries).
 * mcmc_stats       running statistics of the outer loop, examples/stochastic_volatility/experiment.py:98-127:
 *                  stats[c] = {n, acceptance mean, windowed acceptance (EMA with weight beta), mean squared jump};
 *                  `updated` is int32 [n_chains, per_step].
 * philox           counter-based Philox4x32-10 uniforms in (0,1) (fp32: (0,1] -- ((w >> 8) + 0.5) / 2^24 rounds to 1.0f
 *                  when the 24 bits are all ones; consumers clamp) / Box-Muller normals, always finite (stand-in for jax.random):
 *                  element stream = f(seed, *counter); *counter (device memory, may be NULL = 0) is incremented by
 *                  one after the fill, so a captured CUDA graph draws fresh numbers on every replay.  Restated bit
 *                  for bit (uniforms) in oracle/philox.py.                                                        */
int32_t auxssm_csmc_grad_log_target_f32(void* stream, const auxssm_csmc_desc* desc, const float* u, const float* m0,
                                        const float* chol_P0, const float* F, const float* b, const float* chol_Q,
                                        const float* ys, float* grad);
int32_t auxssm_csmc_grad_log_target_f64(void* stream, const auxssm_csmc_desc* desc, const double* u, const double* m0,
                                        const double* chol_P0, const double* F, const double* b, const double* chol_Q,
                                        const double* ys, double* grad);
int32_t auxssm_cholesky_f32(void* stream, int64_t n, int32_t d, const float* A, float* L);
int32_t auxssm_cholesky_f64(void* stream, int64_t n, int32_t d, const double* A, double* L);
int32_t auxssm_delta_adaptation_f32(void* stream, int64_t n, const float* delta, const float* acc_rate,
                                    int32_t acc_stride, double target_rate, double adaptation_rate, double min_delta,
                                    double max_delta, float* delta_out);
int32_t auxssm_delta_adaptation_f64(void* stream, int64_t n, const double* delta, const double* acc_rate,
                                    int32_t acc_stride, double target_rate, double adaptation_rate, double min_delta,
                                    double max_delta, double* delta_out);
int32_t auxssm_mcmc_stats_f32(void* stream, int32_t n_chains, int64_t n_per_chain, int32_t per_step, const float* x_old,
                              const float* x_new, const int32_t* updated, double beta, float* stats);
int32_t auxssm_mcmc_stats_f64(void* stream, int32_t n_chains, int64_t n_per_chain, int32_t per_step,
                              const double* x_old, const double* x_new, const int32_t* updated, double beta,
                              double* stats);
/* autocov: per-chain mean[n_chains] and biased autocovariances acov[n_chains, n] (double) of x[n_chains, n], the
 * quantities examples/rare_event/ess.py:69-79 forms with an FFT (n <= ~28 000: one chain is held in shared memory). */
int32_t auxssm_autocov_f32(void* stream, int32_t n_chains, int64_t n, const float* x, double* mean, double* acov);
int32_t auxssm_autocov_f64(void* stream, int32_t n_chains, int64_t n, const double* x, double* mean, double* acov);
/* ---- Sigma-point linearisation: replaces _primitives/linearisation.py:110-139
 * (_generic_sigma_points / _cov behind `cubature` :78-107 and `gauss_hermite` :48-75), batched over n items
 * (chain x time step), split at the seam where the reference calls the user's `mean` / `cov` callables:
 *   sigma_points  chol[n,dx,dx] = cholesky(P_star[n,dx,dx]);  points[n,n_pts,dx] = x_star + (chol xi)^T, xi[dx,n_pts]
 *   (caller)      f_pts[n,n_pts,dy] = mean(points),  v_pts = cov(points)
 *   sigma_fit     F[n,dy,dx], Q[n,dy,dy], b[n,dy] from the weights w[n_pts] and the values at the points.
 *                 v_pts element (item i, point p) starts at v_pts + i*v_item_stride + p*v_point_stride
 *                 (strides 0 = a covariance shared by all items / all points).        dx, dy <= 32.             */
int32_t auxssm_sigma_points_f32(void* stream, int64_t n, int32_t dx, int32_t n_pts, const float* x_star,
                                const float* P_star, const float* xi, float* points, float* chol);
int32_t auxssm_sigma_points_f64(void* stream, int64_t n, int32_t dx, int32_t n_pts, const double* x_star,
                                const double* P_star, const double* xi, double* points, double* chol);
int32_t auxssm_sigma_fit_f32(void* stream, int64_t n, int32_t dx, int32_t dy, int32_t n_pts, const float* w,
                             const float* points, const float* x_star, const float* chol, const float* f_pts,
                             const float* v_pts, int64_t v_item_stride, int64_t v_point_stride, float* F, float* Q,
                             float* b);
int32_t auxssm_sigma_fit_f64(void* stream, int64_t n, int32_t dx, int32_t dy, int32_t n_pts, const double* w,
                             const double* points, const double* x_star, const double* chol, const double* f_pts,
                             const double* v_pts, int64_t v_item_stride, int64_t v_point_stride, double* F, double* Q,
                             double* b);
int32_t auxssm_philox_uniform_f32(void* stream, uint64_t seed, uint64_t* counter, int64_t n, float* out);
int32_t auxssm_philox_uniform_f64(void* stream, uint64_t seed, uint64_t* counter, int64_t n, double* out);
int32_t auxssm_philox_normal_f32(void* stream, uint64_t seed, uint64_t* counter, int64_t n, float* out);
int32_t auxssm_philox_normal_f64(void* stream, uint64_t seed, uint64_t* counter, int64_t n, double* out);
int32_t auxssm_philox_advance(void* stream, uint64_t* counter, uint64_t by); /* *counter += by, on the stream */

#ifdef __cplusplus
}
#endif
#endif /* AUXSSM_H */

// Elementwise / reduction pieces of the auxiliary Kalman MH step.
//
// Replaces, in aux_samplers/kalman/generic.py of the reference:
//   :61      u = x + sqrt(delta / 2) * normal(key, x.shape)
//   :98-106  _get_alpha   (incl. the sum((x'-u)^2 - (x-u)^2) / delta correction)
//   :73-74   accept = bernoulli(key, alpha)  [== uniform < alpha];  x = select(accept, x_prop, x)
// One CTA per chain for the accept step (block reduction over the T*d trajectory entries, fixed
// order => deterministic), grid-stride vectorised loop for the auxiliary draw.
#include "common.cuh"
#include "smem_linalg.cuh"  // ldv / stv (vector accesses)

namespace auxssm {

template <typename T>
__global__ void auxiliary_draw_kernel(int64_t n_chains, int64_t n_per_chain, const T* __restrict__ x,
                                      const T* __restrict__ normals, const T* __restrict__ delta, int delta_stride,
                                      T* __restrict__ u) {
  const int64_t total = n_chains * n_per_chain;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e / n_per_chain;
    const T s = sqrt(T(0.5) * delta[c * delta_stride]);
    u[e] = x[e] + s * normals[e];
  }
}

// Long chains (the aux-Kalman sweep: n_per_chain = T d): blockIdx.y walks the chains, so the scale is computed once per
// thread instead of a 64-bit division and a square root per entry, and a thread moves VEC consecutive entries.
template <typename T, int VEC>
__global__ void auxiliary_draw_rows_kernel(int64_t n_chains, int64_t n_per_chain, const T* __restrict__ x,
                                           const T* __restrict__ normals, const T* __restrict__ delta,
                                           int delta_stride, T* __restrict__ u) {
  for (int64_t c = blockIdx.y; c < n_chains; c += gridDim.y) {
    const T s = sqrt(T(0.5) * delta[c * delta_stride]);
    const int64_t base = c * n_per_chain;
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; e < n_per_chain;
         e += (int64_t)gridDim.x * blockDim.x * VEC) {
      T xv[VEC], nv[VEC], uv[VEC];
      ldv<T, VEC>(xv, x + base + e);
      ldv<T, VEC>(nv, normals + base + e);
#pragma unroll
      for (int q = 0; q < VEC; ++q) uv[q] = xv[q] + s * nv[q];
      stv<T, VEC>(u + base + e, uv);
    }
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) kalman_mh_kernel(int64_t n_per_chain, const T* __restrict__ delta,
                                                       int delta_stride, const T* __restrict__ lp_prop,
                                                       const T* __restrict__ lp_rev, const T* __restrict__ lt_prop,
                                                       const T* __restrict__ lt_rev, const T* __restrict__ u,
                                                       const T* __restrict__ x_prop, const T* __restrict__ uniforms,
                                                       T* __restrict__ x, int32_t* __restrict__ accepted,
                                                       T* __restrict__ alpha_out) {
  __shared__ T red[NT / 32];
  __shared__ int acc_flag;
  const int64_t c = blockIdx.x;
  const int64_t base = c * n_per_chain;
  T part = T(0);
  for (int64_t e = threadIdx.x; e < n_per_chain; e += NT) {
    const T uu = u[base + e];
    const T a = x_prop[base + e] - uu, b = x[base + e] - uu;
    part += a * a - b * b;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    T corr = T(0);
    for (int q = 0; q < NT / 32; ++q) corr += red[q];
    corr /= delta[c * delta_stride];
    T log_alpha = lt_prop[c] - lt_rev[c] + lp_rev[c] - lp_prop[c] - corr;
    const T alpha = exp(log_alpha < T(0) ? log_alpha : T(0));
    const int acc = uniforms[c] < alpha ? 1 : 0;  // NaN alpha => reject, as bernoulli(key, nan) does
    accepted[c] = acc;
    if (alpha_out) alpha_out[c] = alpha;
    acc_flag = acc;
  }
  __syncthreads();
  if (acc_flag) {
    for (int64_t e = threadIdx.x; e < n_per_chain; e += NT) x[base + e] = x_prop[base + e];
  }
}

template <typename T>
static int aux_draw_entry(void* stream, int64_t n_chains, int64_t n_per_chain, const T* x, const T* normals,
                          const T* delta, int32_t delta_stride, T* u) {
  if (!x || !normals || !delta || !u || n_chains < 1 || n_per_chain < 1 || (delta_stride != 0 && delta_stride != 1)) {
    set_error("auxssm_auxiliary_draw: invalid argument");
    return AUXSSM_ERR_INVALID;
  }
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (n_per_chain >= 4096 && n_per_chain % 4 == 0 && al16(x) && al16(normals) && al16(u)) {
    const unsigned gx = (unsigned)((n_per_chain / 4 + 255) / 256), gy = (unsigned)(n_chains < 65535 ? n_chains : 65535);
    auxiliary_draw_rows_kernel<T, 4><<<dim3(gx, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_chains, n_per_chain, x, normals, delta, delta_stride, u);
    AUXSSM_CHECK_LAUNCH("auxssm_auxiliary_draw");
    return AUXSSM_OK;
  }
  const int64_t total = n_chains * n_per_chain;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = 16L * device_sm_count();
  if (blocks > cap) blocks = cap;
  auxiliary_draw_kernel<T><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(n_chains, n_per_chain, x,
                                                                                           normals, delta,
                                                                                           delta_stride, u);
  AUXSSM_CHECK_LAUNCH("auxssm_auxiliary_draw");
  return AUXSSM_OK;
}

template <typename T>
static int mh_entry(void* stream, int64_t n_chains, int64_t n_per_chain, const T* delta, int32_t delta_stride,
                    const T* lp_prop, const T* lp_rev, const T* lt_prop, const T* lt_rev, const T* u, const T* x_prop,
                    const T* uniforms, T* x, int32_t* accepted, T* alpha) {
  if (!delta || !lp_prop || !lp_rev || !lt_prop || !lt_rev || !u || !x_prop || !uniforms || !x || !accepted ||
      n_chains < 1 || n_per_chain < 1 || (delta_stride != 0 && delta_stride != 1)) {
    set_error("auxssm_kalman_mh: invalid argument");
    return AUXSSM_ERR_INVALID;
  }
  kalman_mh_kernel<T, 256><<<(unsigned)n_chains, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n_per_chain, delta, delta_stride, lp_prop, lp_rev, lt_prop, lt_rev, u, x_prop, uniforms, x, accepted, alpha);
  AUXSSM_CHECK_LAUNCH("auxssm_kalman_mh");
  return AUXSSM_OK;
}

}  // namespace auxssm

extern "C" __attribute__((visibility("default"))) int32_t auxssm_auxiliary_draw_f32(void* stream, int64_t n_chains, int64_t n_per_chain, const float* x,
                                             const float* normals, const float* delta, int32_t delta_stride,
                                             float* u) {
  return auxssm::aux_draw_entry<float>(stream, n_chains, n_per_chain, x, normals, delta, delta_stride, u);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_auxiliary_draw_f64(void* stream, int64_t n_chains, int64_t n_per_chain, const double* x,
                                             const double* normals, const double* delta, int32_t delta_stride,
                                             double* u) {
  return auxssm::aux_draw_entry<double>(stream, n_chains, n_per_chain, x, normals, delta, delta_stride, u);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_mh_f32(void* stream, int64_t n_chains, int64_t n_per_chain, const float* delta,
                                        int32_t delta_stride, const float* log_lgssm_prop, const float* log_lgssm_rev,
                                        const float* log_target_prop, const float* log_target_rev, const float* u,
                                        const float* x_prop, const float* uniforms, float* x, int32_t* accepted,
                                        float* alpha) {
  return auxssm::mh_entry<float>(stream, n_chains, n_per_chain, delta, delta_stride, log_lgssm_prop, log_lgssm_rev,
                                 log_target_prop, log_target_rev, u, x_prop, uniforms, x, accepted, alpha);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_kalman_mh_f64(void* stream, int64_t n_chains, int64_t n_per_chain, const double* delta,
                                        int32_t delta_stride, const double* log_lgssm_prop,
                                        const double* log_lgssm_rev, const double* log_target_prop,
                                        const double* log_target_rev, const double* u, const double* x_prop,
                                        const double* uniforms, double* x, int32_t* accepted, double* alpha) {
  return auxssm::mh_entry<double>(stream, n_chains, n_per_chain, delta, delta_stride, log_lgssm_prop, log_lgssm_rev,
                                  log_target_prop, log_target_rev, u, x_prop, uniforms, x, accepted, alpha);
}

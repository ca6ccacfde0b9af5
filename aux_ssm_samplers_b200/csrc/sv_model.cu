// Proposal construction for the multivariate stochastic-volatility model:
// the step immediately before the Kalman filter on every sweep, fused into one pass over (chain, t, dim).
//
// Replaces, in the reference's examples/stochastic_volatility:
//   model.py:56-82            log_potential / its gradient and (diagonal) Hessian,
//                             g(x, y) = log N(y; 0, exp(x/2)^2) = -y^2 e^{-x}/2 - x/2 - log(2 pi)/2
//   auxiliary_kalman.py:28-35 first-order observations factory   ys' = u + (delta/2) grad,  R = (delta/2) I
//   auxiliary_kalman.py:37-48 second-order observations factory  Omega = (-hess + 2 I/delta)^{-1},
//                             ys' = Omega (2u/delta + grad - hess x),  R_t = Omega_t  (diagonal, stored dense
//                             [.., d, d] exactly like the reference materialises it)
//   auxiliary_kalman.py:50-54 the potential part of log_likelihood_fn (sum of g over t and dims).
#include "common.cuh"
#include "smem_linalg.cuh"  // ldv / stv (vector accesses)

namespace auxssm {

template <typename T> __device__ __forceinline__ T nan_to_num_(T v);
template <> __device__ __forceinline__ float nan_to_num_<float>(float v) {
  if (v != v) return 0.f;
  if (isinf(v)) return v > 0 ? 3.402823466e+38f : -3.402823466e+38f;
  return v;
}
template <> __device__ __forceinline__ double nan_to_num_<double>(double v) {
  if (v != v) return 0.0;
  if (isinf(v)) return v > 0 ? 1.7976931348623157e+308 : -1.7976931348623157e+308;
  return v;
}

// One element of ys' (and of the diagonal of Omega_t, second order).
template <typename T>
__device__ __forceinline__ void sv_obs_elem(int order, T dl, T y, T xv, T uv, T& aux, T& omega) {
  const T w = T(0.5) * y * y * exp(-xv);  // y^2 e^{-x} / 2
  const T grad = w - T(0.5);
  const T hess = -w;
  if (order == 1) {
    aux = uv + T(0.5) * dl * nan_to_num_(grad);
    omega = T(0);
  } else {
    omega = T(1) / (-hess + T(2) / dl);
    aux = omega * (T(2) * uv / dl + grad - hess * xv);
  }
}

// ys' (and diag R_t): blockIdx.y walks the chains, a thread owns VEC consecutive (t, dim) entries of one chain -- no
// 64-bit division per entry, 16-byte accesses when the chain length and the bases allow (VEC = 4); the dense R_t
// block of the reference layout is written by sv_second_order_R_kernel.
template <typename T, int VEC>
__global__ void sv_obs_factory_kernel(int order, int64_t n_chains, int64_t per, const T* __restrict__ ys,
                                      int64_t ys_chain_stride, const T* __restrict__ x, const T* __restrict__ u,
                                      const T* __restrict__ delta, int delta_stride, T* __restrict__ aux_ys,
                                      T* __restrict__ Rdiag) {
  for (int64_t c = blockIdx.y; c < n_chains; c += gridDim.y) {
    const T dl = delta[c * delta_stride];
    const T* yc = ys + c * ys_chain_stride;
    const T *xc = x + c * per, *uc = u + c * per;
    T* ac = aux_ys + c * per;
    T* rc = Rdiag ? Rdiag + c * per : nullptr;
    for (int64_t td = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; td < per;
         td += (int64_t)gridDim.x * blockDim.x * VEC) {
      T yv[VEC], xv[VEC], uv[VEC], av[VEC], ov[VEC];
      ldv<T, VEC>(yv, yc + td);
      ldv<T, VEC>(xv, xc + td);
      ldv<T, VEC>(uv, uc + td);
#pragma unroll
      for (int q = 0; q < VEC; ++q) sv_obs_elem<T>(order, dl, yv[q], xv[q], uv[q], av[q], ov[q]);
      stv<T, VEC>(ac + td, av);
      if (order == 2 && rc) stv<T, VEC>(rc + td, ov);  // diagonal layout: R[c, t, i]
    }
  }
}

// R[c, t] = diag(Omega_t), dense and written with consecutive threads on consecutive addresses.
template <typename T>
__global__ void sv_second_order_R_kernel(int64_t n_chains, int64_t Tn, int d, const T* __restrict__ ys,
                                         int64_t ys_chain_stride, const T* __restrict__ x,
                                         const T* __restrict__ delta, int delta_stride, T* __restrict__ Rs) {
  const int64_t dd = (int64_t)d * d, total = n_chains * Tn * dd;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ct = e / dd;  // c * Tn + t
    const int r = (int)(e - ct * dd);
    const int i = r / d, j = r - i * d;
    T v = T(0);
    if (i == j) {
      const int64_t c = ct / Tn, t = ct - c * Tn;
      const T y = ys[c * ys_chain_stride + t * d + i];
      const T w = T(0.5) * y * y * exp(-x[ct * d + i]);
      v = T(1) / (w + T(2) / delta[c * delta_stride]);
    }
    Rs[e] = v;
  }
}

// R[c] = (delta_c / 2) I   (first order; one block per chain, time stride 0); dense d x d or diagonal d
template <typename T>
__global__ void sv_first_order_R_kernel(int64_t n_chains, int d, const T* __restrict__ delta, int delta_stride,
                                        T* __restrict__ Rs, int dense) {
  const int64_t per = dense ? (int64_t)d * d : d, total = n_chains * per;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e / per;
    const int r = (int)(e % per);
    Rs[e] = (!dense || r / d == r % d) ? T(0.5) * delta[c * delta_stride] : T(0);
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) sv_log_potential_kernel(int64_t n_per_chain, const T* __restrict__ ys,
                                                              int64_t ys_chain_stride, const T* __restrict__ x,
                                                              T* __restrict__ out, int accumulate) {
  __shared__ T red[NT / 32];
  const int64_t c = blockIdx.x;
  T part = T(0);
  for (int64_t e = threadIdx.x; e < n_per_chain; e += NT) {
    const T y = ys[c * ys_chain_stride + e], xv = x[c * n_per_chain + e];
    const T val = T(-0.5) * y * y * exp(-xv) - T(0.5) * xv - T(0.91893853320467274178);
    part += nan_to_num_(val);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
    for (int q = 0; q < NT / 32; ++q) s += red[q];
    out[c] = accumulate ? out[c] + s : s;
  }
}

template <typename T>
static int sv_factory_entry(void* stream, int32_t order, int64_t n_chains, int64_t Tn, int32_t d, const T* ys,
                            int64_t ys_chain_stride, const T* x, const T* u, const T* delta, int32_t delta_stride,
                            T* aux_ys, T* Rs, int32_t dense_R) {
  if ((order != 1 && order != 2) || n_chains < 1 || Tn < 1 || d < 1 || !ys || !x || !u || !delta || !aux_ys || !Rs ||
      (delta_stride != 0 && delta_stride != 1)) {
    set_error("auxssm_sv_obs_factory: invalid argument");
    return AUXSSM_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t per = Tn * d, total = n_chains * per;
  const int64_t cap = 32L * device_sm_count();
  T* rdiag = (order == 2 && !dense_R) ? Rs : nullptr;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  const bool vec4 = per % 4 == 0 && ys_chain_stride % 4 == 0 && al16(ys) && al16(x) && al16(u) && al16(aux_ys) &&
                    (rdiag == nullptr || al16(rdiag));
  const unsigned gy = (unsigned)(n_chains < 65535 ? n_chains : 65535);
  if (vec4) {
    const unsigned gx = (unsigned)((per / 4 + 255) / 256);
    sv_obs_factory_kernel<T, 4><<<dim3(gx, gy), 256, 0, st>>>(order, n_chains, per, ys, ys_chain_stride, x, u, delta,
                                                              delta_stride, aux_ys, rdiag);
  } else {
    const unsigned gx = (unsigned)((per + 255) / 256);
    sv_obs_factory_kernel<T, 1><<<dim3(gx, gy), 256, 0, st>>>(order, n_chains, per, ys, ys_chain_stride, x, u, delta,
                                                              delta_stride, aux_ys, rdiag);
  }
  AUXSSM_CHECK_LAUNCH("auxssm_sv_obs_factory");
  if (order == 1) {
    const int64_t tot = n_chains * (dense_R ? (int64_t)d * d : d);
    sv_first_order_R_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n_chains, d, delta, delta_stride, Rs,
                                                                            dense_R);
    AUXSSM_CHECK_LAUNCH("auxssm_sv_obs_factory (R)");
  } else if (dense_R) {
    int64_t rb = (total * d + 255) / 256;
    if (rb > 64 * cap) rb = 64 * cap;
    sv_second_order_R_kernel<T><<<(unsigned)rb, 256, 0, st>>>(n_chains, Tn, d, ys, ys_chain_stride, x, delta,
                                                              delta_stride, Rs);
    AUXSSM_CHECK_LAUNCH("auxssm_sv_obs_factory (R)");
  }
  return AUXSSM_OK;
}

template <typename T>
static int sv_potential_entry(void* stream, int64_t n_chains, int64_t n_per_chain, const T* ys,
                              int64_t ys_chain_stride, const T* x, T* out, int32_t accumulate) {
  if (n_chains < 1 || n_per_chain < 1 || !ys || !x || !out) {
    set_error("auxssm_sv_log_potential: invalid argument");
    return AUXSSM_ERR_INVALID;
  }
  sv_log_potential_kernel<T, 256><<<(unsigned)n_chains, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n_per_chain, ys, ys_chain_stride, x, out, accumulate);
  AUXSSM_CHECK_LAUNCH("auxssm_sv_log_potential");
  return AUXSSM_OK;
}

}  // namespace auxssm

extern "C" __attribute__((visibility("default"))) int32_t auxssm_sv_obs_factory_f32(void* stream, int32_t order, int64_t n_chains, int64_t T, int32_t d,
                                             const float* ys, int64_t ys_chain_stride, const float* x, const float* u,
                                             const float* delta, int32_t delta_stride, float* aux_ys, float* Rs,
                                             int32_t dense_R) {
  return auxssm::sv_factory_entry<float>(stream, order, n_chains, T, d, ys, ys_chain_stride, x, u, delta, delta_stride,
                                         aux_ys, Rs, dense_R);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_sv_obs_factory_f64(void* stream, int32_t order, int64_t n_chains, int64_t T, int32_t d,
                                             const double* ys, int64_t ys_chain_stride, const double* x,
                                             const double* u, const double* delta, int32_t delta_stride,
                                             double* aux_ys, double* Rs, int32_t dense_R) {
  return auxssm::sv_factory_entry<double>(stream, order, n_chains, T, d, ys, ys_chain_stride, x, u, delta,
                                          delta_stride, aux_ys, Rs, dense_R);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_sv_log_potential_f32(void* stream, int64_t n_chains, int64_t n_per_chain, const float* ys,
                                               int64_t ys_chain_stride, const float* x, float* out,
                                               int32_t accumulate) {
  return auxssm::sv_potential_entry<float>(stream, n_chains, n_per_chain, ys, ys_chain_stride, x, out, accumulate);
}
extern "C" __attribute__((visibility("default"))) int32_t auxssm_sv_log_potential_f64(void* stream, int64_t n_chains, int64_t n_per_chain, const double* ys,
                                               int64_t ys_chain_stride, const double* x, double* out,
                                               int32_t accumulate) {
  return auxssm::sv_potential_entry<double>(stream, n_chains, n_per_chain, ys, ys_chain_stride, x, out, accumulate);
}

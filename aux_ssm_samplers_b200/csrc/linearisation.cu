// Sigma-point (statistical) linearisation of conditional moments, batched over (chain, time step).
//
// Replaces aux_samplers/_primitives/linearisation.py: _generic_sigma_points (:110-133), _cov (:136-139), used by
// `cubature` (:78-107) and `gauss_hermite` (:48-75).  The reference evaluates the user's `mean` / `cov` callables at the
// sigma points in the middle of the routine; a CUDA kernel cannot call Python, so the routine is split at that seam:
//   sigma_points_kernel  chol = cholesky(P*),  points_p = x* + chol xi_p                          (:111-114)
//   (caller)             f_p = mean(points_p, params),  V_p = cov(points_p, params)               (:116, :122)
//   sigma_fit_kernel     m_f = sum_p w_p f_p;  Psi = sum_p w_p (points_p - x*)(f_p - m_f)^T;  F = (cho_solve(chol, Psi))^T;
//                        Q = sum_p w_p (f_p - m_f)(f_p - m_f)^T - (F chol)(F chol)^T + sum_p w_p V_p;  b = m_f - F x*
// One CTA of 128 threads per item; all matrices of an item live in shared memory (dx, dy <= 32).  Host-side factory
// work (once per sweep per time step), not a hot loop: plain loops, no tuning.
#include "common.cuh"

namespace auxssm {

constexpr int kLinThreads = 128;
constexpr int kLinMaxD = 32;

// in-place lower Cholesky of A [d, d] (row-major, shared); jnp.linalg.cholesky semantics: NaN on failure
template <typename T> __device__ __forceinline__ void lin_chol(T* A, int d) {
  const int tid = threadIdx.x;
  for (int k = 0; k < d; ++k) {
    __syncthreads();
    const T lkk = sqrt(A[k * d + k]);
    __syncthreads();
    if (tid == 0) A[k * d + k] = lkk;
    for (int i = k + 1 + tid; i < d; i += kLinThreads) A[i * d + k] = A[i * d + k] / lkk;
    __syncthreads();
    for (int e = tid; e < d * d; e += kLinThreads) {
      const int i = e / d, j = e - i * d;
      if (i > k && j > k && j <= i) A[i * d + j] -= A[i * d + k] * A[j * d + k];
    }
  }
  __syncthreads();
  for (int e = tid; e < d * d; e += kLinThreads) {
    const int i = e / d, j = e - i * d;
    if (j > i) A[e] = T(0);
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kLinThreads) sigma_points_kernel(int dx, int n_pts, const T* __restrict__ x_star,
                                                                    const T* __restrict__ P_star, const T* __restrict__ xi,
                                                                    T* __restrict__ points, T* __restrict__ chol) {
  __shared__ T L[kLinMaxD * kLinMaxD];
  __shared__ T xs[kLinMaxD];
  const int64_t n = blockIdx.x;
  const int tid = threadIdx.x;
  for (int e = tid; e < dx * dx; e += kLinThreads) L[e] = P_star[n * dx * dx + e];
  if (tid < dx) xs[tid] = x_star[n * dx + tid];
  lin_chol<T>(L, dx);
  for (int e = tid; e < dx * dx; e += kLinThreads) chol[n * dx * dx + e] = L[e];
  for (int e = tid; e < n_pts * dx; e += kLinThreads) {
    const int p = e / dx, i = e - p * dx;
    T acc = T(0);
    for (int j = 0; j <= i; ++j) acc += L[i * dx + j] * xi[j * n_pts + p];
    points[(n * n_pts + p) * dx + i] = xs[i] + acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(kLinThreads) sigma_fit_kernel(int dx, int dy, int n_pts, const T* __restrict__ w,
                                                                 const T* __restrict__ points, const T* __restrict__ x_star,
                                                                 const T* __restrict__ chol, const T* __restrict__ f_pts,
                                                                 const T* __restrict__ v_pts, int64_t v_item_stride,
                                                                 int64_t v_point_stride, T* __restrict__ F, T* __restrict__ Q,
                                                                 T* __restrict__ b) {
  __shared__ T L[kLinMaxD * kLinMaxD], Psi[kLinMaxD * kLinMaxD], Tm[kLinMaxD * kLinMaxD];
  __shared__ T mf[kLinMaxD], xs[kLinMaxD];
  const int64_t n = blockIdx.x;
  const int tid = threadIdx.x;
  const T* pts = points + n * n_pts * dx;
  const T* fp = f_pts + n * n_pts * dy;
  for (int e = tid; e < dx * dx; e += kLinThreads) L[e] = chol[n * dx * dx + e];
  if (tid < dx) xs[tid] = x_star[n * dx + tid];
  if (tid < dy) {
    T acc = T(0);
    for (int p = 0; p < n_pts; ++p) acc += w[p] * fp[p * dy + tid];
    mf[tid] = acc;
  }
  __syncthreads();
  // Psi [dx, dy]
  for (int e = tid; e < dx * dy; e += kLinThreads) {
    const int i = e / dy, j = e - i * dy;
    T acc = T(0);
    for (int p = 0; p < n_pts; ++p) acc += (pts[p * dx + i] - xs[i]) * w[p] * (fp[p * dy + j] - mf[j]);
    Psi[i * dy + j] = acc;
  }
  __syncthreads();
  // Y = (L L^T)^{-1} Psi, column j by thread j;  F = Y^T  [dy, dx]
  if (tid < dy) {
    const int j = tid;
    for (int i = 0; i < dx; ++i) {
      T acc = Psi[i * dy + j];
      for (int k = 0; k < i; ++k) acc -= L[i * dx + k] * Psi[k * dy + j];
      Psi[i * dy + j] = acc / L[i * dx + i];
    }
    for (int i = dx - 1; i >= 0; --i) {
      T acc = Psi[i * dy + j];
      for (int k = i + 1; k < dx; ++k) acc -= L[k * dx + i] * Psi[k * dy + j];
      Psi[i * dy + j] = acc / L[i * dx + i];
    }
  }
  __syncthreads();
  // temp = F chol  [dy, dx];  F out;  b = m_f - F x*
  for (int e = tid; e < dy * dx; e += kLinThreads) {
    const int a = e / dx, k = e - a * dx;
    T acc = T(0);
    for (int i = k; i < dx; ++i) acc += Psi[i * dy + a] * L[i * dx + k];
    Tm[a * dx + k] = acc;
    F[n * dy * dx + e] = Psi[k * dy + a];
  }
  if (tid < dy) {
    T acc = T(0);
    for (int i = 0; i < dx; ++i) acc += Psi[i * dy + tid] * xs[i];
    b[n * dy + tid] = mf[tid] - acc;
  }
  __syncthreads();
  // Q = Phi - temp temp^T + v_f
  const T* vp = v_pts + n * v_item_stride;
  for (int e = tid; e < dy * dy; e += kLinThreads) {
    const int a = e / dy, c = e - a * dy;
    T phi = T(0), vf = T(0), tt = T(0);
    for (int p = 0; p < n_pts; ++p) {
      phi += (fp[p * dy + a] - mf[a]) * w[p] * (fp[p * dy + c] - mf[c]);
      vf += w[p] * vp[p * v_point_stride + e];
    }
    for (int k = 0; k < dx; ++k) tt += Tm[a * dx + k] * Tm[c * dx + k];
    Q[n * dy * dy + e] = phi - tt + vf;
  }
}

static int check_lin(int64_t n, int dx, int dy, int n_pts, const char* who) {
  if (n < 0 || dx < 1 || dy < 1 || n_pts < 1) { set_error("%s: bad sizes", who); return AUXSSM_ERR_INVALID; }
  if (dx > kLinMaxD || dy > kLinMaxD) { set_error("%s: dx, dy <= %d (got %d, %d)", who, kLinMaxD, dx, dy); return AUXSSM_ERR_UNSUPPORTED; }
  return AUXSSM_OK;
}

template <typename T>
static int sigma_points_entry(void* stream, int64_t n, int dx, int n_pts, const T* x_star, const T* P_star, const T* xi,
                              T* points, T* chol) {
  const char* who = "auxssm_sigma_points";
  int rc = check_lin(n, dx, 1, n_pts, who);
  if (rc) return rc;
  if (!x_star || !P_star || !xi || !points || !chol) { set_error("%s: null pointer", who); return AUXSSM_ERR_INVALID; }
  if (n == 0) return AUXSSM_OK;
  sigma_points_kernel<T><<<(unsigned)n, kLinThreads, 0, static_cast<cudaStream_t>(stream)>>>(dx, n_pts, x_star, P_star, xi, points, chol);
  AUXSSM_CHECK_LAUNCH("auxssm_sigma_points");
  return AUXSSM_OK;
}

template <typename T>
static int sigma_fit_entry(void* stream, int64_t n, int dx, int dy, int n_pts, const T* w, const T* points, const T* x_star,
                           const T* chol, const T* f_pts, const T* v_pts, int64_t v_item_stride, int64_t v_point_stride,
                           T* F, T* Q, T* b) {
  const char* who = "auxssm_sigma_fit";
  int rc = check_lin(n, dx, dy, n_pts, who);
  if (rc) return rc;
  if (!w || !points || !x_star || !chol || !f_pts || !v_pts || !F || !Q || !b) { set_error("%s: null pointer", who); return AUXSSM_ERR_INVALID; }
  if (n == 0) return AUXSSM_OK;
  sigma_fit_kernel<T><<<(unsigned)n, kLinThreads, 0, static_cast<cudaStream_t>(stream)>>>(dx, dy, n_pts, w, points, x_star, chol, f_pts,
                                                                                         v_pts, v_item_stride, v_point_stride, F, Q, b);
  AUXSSM_CHECK_LAUNCH("auxssm_sigma_fit");
  return AUXSSM_OK;
}

}  // namespace auxssm

#define AUXSSM_EXPORT extern "C" __attribute__((visibility("default")))
AUXSSM_EXPORT int32_t auxssm_sigma_points_f32(void* stream, int64_t n, int32_t dx, int32_t n_pts, const float* x_star,
                                              const float* P_star, const float* xi, float* points, float* chol) {
  return auxssm::sigma_points_entry<float>(stream, n, dx, n_pts, x_star, P_star, xi, points, chol);
}
AUXSSM_EXPORT int32_t auxssm_sigma_points_f64(void* stream, int64_t n, int32_t dx, int32_t n_pts, const double* x_star,
                                              const double* P_star, const double* xi, double* points, double* chol) {
  return auxssm::sigma_points_entry<double>(stream, n, dx, n_pts, x_star, P_star, xi, points, chol);
}
AUXSSM_EXPORT int32_t auxssm_sigma_fit_f32(void* stream, int64_t n, int32_t dx, int32_t dy, int32_t n_pts, const float* w,
                                           const float* points, const float* x_star, const float* chol, const float* f_pts,
                                           const float* v_pts, int64_t v_item_stride, int64_t v_point_stride, float* F,
                                           float* Q, float* b) {
  return auxssm::sigma_fit_entry<float>(stream, n, dx, dy, n_pts, w, points, x_star, chol, f_pts, v_pts, v_item_stride,
                                        v_point_stride, F, Q, b);
}
AUXSSM_EXPORT int32_t auxssm_sigma_fit_f64(void* stream, int64_t n, int32_t dx, int32_t dy, int32_t n_pts, const double* w,
                                           const double* points, const double* x_star, const double* chol,
                                           const double* f_pts, const double* v_pts, int64_t v_item_stride,
                                           int64_t v_point_stride, double* F, double* Q, double* b) {
  return auxssm::sigma_fit_entry<double>(stream, n, dx, dy, n_pts, w, points, x_star, chol, f_pts, v_pts, v_item_stride,
                                         v_point_stride, F, Q, b);
}

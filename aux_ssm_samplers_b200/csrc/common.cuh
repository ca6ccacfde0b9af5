// Shared host/device helpers for libauxssm: error reporting, descriptor checks, dispatch on D.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/auxssm.h"

namespace auxssm {

void set_error(const char* fmt, ...);
void count_launch();
void note_dispatch(const char* who);  // appends to the calling thread's dispatch log (auxssm_dispatch_log)

inline int padded_dim(int d) {
  if (d <= 4) return 4;
  if (d <= 8) return 8;
  if (d <= 16) return 16;
  if (d <= 32) return 32;
  if (d <= 64) return 64;
  return -1;
}

int device_sm_count();
size_t device_max_smem();

// Resolve the number of time chunks actually used for a descriptor (policy shared by the
// scratch-size queries and the launches).
int resolve_time_chunks(const auxssm_lgssm_desc* d);

int check_desc(const auxssm_lgssm_desc* d, const char* who);

__host__ __device__ inline int64_t unit_off(const auxssm_stride& s, int c, int b) {
  return (int64_t)c * s.chain + (int64_t)b * s.batch;
}

#define AUXSSM_CHECK_LAUNCH(who)                                              \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      auxssm::set_error("%s: CUDA launch failed: %s", who, cudaGetErrorString(e__)); \
      return AUXSSM_ERR_LAUNCH;                                               \
    }                                                                         \
    auxssm::count_launch();                                                   \
    auxssm::note_dispatch(who);                                               \
  } while (0)

// ---- AUXSSM_FLAG_F_IS_DIAGONAL is a HINT, verified on the device on every call -------------------------------
// `count_offdiag_kernel` counts the non-zero off-diagonal entries of every F the call can read into gate[0]
// (zeroed by a cudaMemsetAsync on the same stream).  The structured kernels (which read diag(F) only) start with
// `if (*gate != 0) return;`, the general kernels launched right behind them with `if (*gate == 0) return;`:
// exactly one of the two does the work, so a wrong hint costs time, never correctness, and nothing synchronises.
template <typename T>
__global__ void count_offdiag_kernel(const T* __restrict__ F, auxssm_stride st, int d, int n_chains, int n_batch,
                                     int* gate) {
  const int nc = st.chain != 0 ? n_chains : 1, nb = st.batch != 0 ? n_batch : 1;
  const int64_t total = (int64_t)nc * nb * d * d;
  int bad = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = e / (d * d);
    const int k = (int)(e - u * d * d), r = k / d, c = k - r * d;
    const T v = F[(u / nb) * st.chain + (u % nb) * st.batch + k];
    if (r != c && v != T(0)) bad = 1;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicAdd(gate, 1);
}

template <typename T>
inline int launch_count_offdiag(cudaStream_t st, const T* F, const auxssm_lgssm_desc& d, int* gate) {
  if (cudaMemsetAsync(gate, 0, sizeof(int), st) != cudaSuccess) {
    set_error("auxssm: cudaMemsetAsync(gate) failed");
    return AUXSSM_ERR_LAUNCH;
  }
  const int nc = d.Fs.chain != 0 ? d.n_chains : 1, nb = d.Fs.batch != 0 ? d.n_batch : 1;
  const int64_t total = (int64_t)nc * nb * d.dx * d.dx;
  int grid = (int)((total + 255) / 256);
  if (grid > 592) grid = 592;
  count_offdiag_kernel<T><<<grid, 256, 0, st>>>(F, d.Fs, d.dx, d.n_chains, d.n_batch, gate);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("auxssm: count_offdiag launch failed: %s", cudaGetErrorString(e));
    return AUXSSM_ERR_LAUNCH;
  }
  count_launch();
  return AUXSSM_OK;
}

constexpr size_t kGateBytes = 256;  // scratch header holding the gate word

template <typename K> inline int ensure_smem(K kernel, size_t bytes, const char* who) {
  if (bytes > device_max_smem()) {
    set_error("%s: needs %zu bytes of shared memory per CTA, device allows %zu", who, bytes, device_max_smem());
    return AUXSSM_ERR_UNSUPPORTED;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", who, bytes, cudaGetErrorString(e));
      return AUXSSM_ERR_LAUNCH;
    }
  }
  return AUXSSM_OK;
}

}  // namespace auxssm

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
  } while (0)

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

// Host-side dispatch of the wide-state cSMC kernel (csmc_wide.cuh) on (D, NPAD) for one proposal family.
#pragma once
#include "csmc_launch.cuh"
#include "csmc_wide.cuh"
#include "csmc_multi.cuh"

namespace auxssm {

// one CTA of 256 threads per resident chain
template <typename T, int D, int PROP, int NPAD>
static int launch_csmc_wide(cudaStream_t st, CsmcArgs<T>& a, int max_slots, int* slots_used) {
  const size_t smem = csmc_wide_smem_elems<T, D>(NPAD) * sizeof(T);
  auto kern = csmc_wide_kernel<T, D, PROP, NPAD>;
  int rc = ensure_smem(kern, smem, "auxssm_csmc_sweep");
  if (rc) return rc;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WideCfg<T, D>::NT, smem);
  if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  long grid = (long)per_sm * device_sm_count();
  if (grid > a.d.n_chains) grid = a.d.n_chains;
  if (!a.hist_per_chain && grid > max_slots) grid = max_slots;
  if (grid < 1) grid = 1;
  if (slots_used) *slots_used = (int)grid;
  if (st == (cudaStream_t)-1) return AUXSSM_OK;  // planning only
  kern<<<(unsigned)grid, WideCfg<T, D>::NT, smem, st>>>(a);
  AUXSSM_CHECK_LAUNCH(NPAD > 128 ? "auxssm_csmc_sweep (csmc_wide_kernel, NPAD 256)" : "auxssm_csmc_sweep (csmc_wide_kernel, NPAD 128)");
  return AUXSSM_OK;
}

template <typename T, int D> static bool wide_fits(int NPAD) {
  return csmc_wide_smem_elems<T, D>(NPAD) * sizeof(T) <= device_max_smem();
}

// four particles per thread for narrow states and many particles (csmc_multi.cuh)
template <typename T, int D, int PROP>
static int launch_csmc_multi(cudaStream_t st, CsmcArgs<T>& a, int max_slots, int* slots_used) {
  const int N = a.d.N;
  const int ntb = (N / 4 + 31) / 32 * 32;
  const size_t smem = csmc_multi_smem_elems<D>(N, a.d.d, PROP) * sizeof(T);
  auto kern = csmc_multi_kernel<T, D, PROP>;
  int rc = ensure_smem(kern, smem, "auxssm_csmc_sweep");
  if (rc) return rc;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ntb, smem);
  if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  long grid = (long)per_sm * device_sm_count();
  if (grid > a.d.n_chains) grid = a.d.n_chains;
  if (!a.hist_per_chain && grid > max_slots) grid = max_slots;
  if (grid < 1) grid = 1;
  if (slots_used) *slots_used = (int)grid;
  if (st == (cudaStream_t)-1) return AUXSSM_OK;  // planning only
  kern<<<(unsigned)grid, ntb, smem, st>>>(a);
  AUXSSM_CHECK_LAUNCH("auxssm_csmc_sweep (csmc_multi_kernel)");
  return AUXSSM_OK;
}

template <typename T, int PROP>
int launch_csmc_wide_family(cudaStream_t st, CsmcArgs<T>& a, int max_slots, int* slots_used) {
  const int d = a.d.d, N = a.d.N;
  if (d <= 4 && N > 256 && N % 4 == 0) {
    return d <= 1 ? launch_csmc_multi<T, 1, PROP>(st, a, max_slots, slots_used)
                  : launch_csmc_multi<T, 4, PROP>(st, a, max_slots, slots_used);
  }
  if (d <= 4 || d > 32 || N > 256) return AUXSSM_WIDE_NA;
  const int npad = N <= 128 ? 128 : 256;
#define AUXSSM_CSMC_WIDE(DD)                                                                  \
  if (wide_fits<T, DD>(npad))                                                                 \
    return npad == 128 ? launch_csmc_wide<T, DD, PROP, 128>(st, a, max_slots, slots_used)     \
                       : launch_csmc_wide<T, DD, PROP, 256>(st, a, max_slots, slots_used)
  if (d <= 8) { AUXSSM_CSMC_WIDE(8); } else { AUXSSM_CSMC_WIDE(32); }
#undef AUXSSM_CSMC_WIDE
  return AUXSSM_WIDE_NA;
}

}  // namespace auxssm

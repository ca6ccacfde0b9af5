// Host-side dispatch of the fused cSMC sweep on (D, MAXT) for one proposal family; each family is
// instantiated in its own translation unit so the families compile in parallel.
#pragma once
#include "csmc_kernel.cuh"

namespace auxssm {

template <typename T, int PROP> int launch_csmc_family(cudaStream_t st, CsmcArgs<T>& a, int max_slots, int* slots_used);

template <typename T, int D, int PROP, int MAXT>
static int launch_csmc_one(cudaStream_t st, CsmcArgs<T>& a, int max_slots, int* slots_used) {
  const int N = a.d.N;
  const int ntb = (N + 31) / 32 * 32;
  const size_t smem = csmc_smem_elems<D>(N, a.d.d, PROP) * sizeof(T);
  auto kern = csmc_sweep_kernel<T, D, PROP, MAXT>;
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
  AUXSSM_CHECK_LAUNCH(MAXT > 256 ? "auxssm_csmc_sweep (csmc_sweep_kernel, <= 1024 threads)"
                                 : "auxssm_csmc_sweep (csmc_sweep_kernel, <= 256 threads)");
  return AUXSSM_OK;
}

// Alternative work decompositions instantiated in csmc_wide_*.cu: the particle-matrix kernel for wide states
// (csmc_wide.cuh: 4 < d <= 32, N <= 256) and the four-particles-per-thread kernel (csmc_multi.cuh: d <= 4,
// N > 256, N % 4 == 0).  Returns AUXSSM_WIDE_NA when the shape is in neither domain.
enum { AUXSSM_WIDE_NA = 1 };
template <typename T, int PROP> int launch_csmc_wide_family(cudaStream_t st, CsmcArgs<T>& a, int max_slots, int* slots_used);

template <typename T, int PROP>
int launch_csmc_family(cudaStream_t st, CsmcArgs<T>& a, int max_slots, int* slots_used) {
  const int d = a.d.d, N = a.d.N;
  const bool big = N > 256;
  const bool general_dynamics = a.d.dyn_time_varying != 0 || a.d.pt_F || a.d.pt_b || a.d.pt_chol_Q;
  if (!general_dynamics) {  // the two re-decomposed kernels hold a time-invariant model with Pt = Mt
    const int rc = launch_csmc_wide_family<T, PROP>(st, a, max_slots, slots_used);
    if (rc != AUXSSM_WIDE_NA) return rc;
  }
#define AUXSSM_CSMC_CASE(DD)                                                                  \
  return big ? launch_csmc_one<T, DD, PROP, 1024>(st, a, max_slots, slots_used)               \
             : launch_csmc_one<T, DD, PROP, 256>(st, a, max_slots, slots_used)
  if (d <= 1) { AUXSSM_CSMC_CASE(1); }
  if (d <= 4) { AUXSSM_CSMC_CASE(4); }
  if (d <= 8) { AUXSSM_CSMC_CASE(8); }
  if (d <= 32) { AUXSSM_CSMC_CASE(32); }
#undef AUXSSM_CSMC_CASE
  set_error("auxssm_csmc_sweep: d = %d exceeds 32", d);
  return AUXSSM_ERR_UNSUPPORTED;
}

}  // namespace auxssm

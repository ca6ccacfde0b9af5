// Wide-state cSMC sweep, BOOTSTRAP proposal family (see csmc_wide.cuh).
#include "csmc_wide_launch.cuh"

namespace auxssm {
template int launch_csmc_wide_family<float, AUXSSM_PROP_BOOTSTRAP>(cudaStream_t, CsmcArgs<float>&, int, int*);
template int launch_csmc_wide_family<double, AUXSSM_PROP_BOOTSTRAP>(cudaStream_t, CsmcArgs<double>&, int, int*);
}  // namespace auxssm

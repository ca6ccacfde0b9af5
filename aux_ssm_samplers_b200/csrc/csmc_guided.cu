// Fused cSMC sweep, GUIDED proposal family (see csmc_kernel.cuh).
#include "csmc_launch.cuh"

namespace auxssm {
template int launch_csmc_family<float, AUXSSM_PROP_GUIDED>(cudaStream_t, CsmcArgs<float>&, int, int*);
template int launch_csmc_family<double, AUXSSM_PROP_GUIDED>(cudaStream_t, CsmcArgs<double>&, int, int*);
}  // namespace auxssm

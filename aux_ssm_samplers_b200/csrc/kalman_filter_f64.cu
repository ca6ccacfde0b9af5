// fp64 instantiations of kalman_filter.cu (separate translation unit: parallel build).
#define AUXSSM_ONLY_F64
#include "kalman_filter.cu"

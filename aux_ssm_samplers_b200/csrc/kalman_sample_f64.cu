// fp64 instantiations of kalman_sample.cu (separate translation unit: parallel build).
#define AUXSSM_ONLY_F64
#include "kalman_sample.cu"

// fp64 instantiations of lgssm_logpdf.cu (separate translation unit: parallel build).
#define AUXSSM_ONLY_F64
#include "lgssm_logpdf.cu"

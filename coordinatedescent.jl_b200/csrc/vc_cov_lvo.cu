// vc_cov_lvo.cu — instances of vc_cov_kernel (vc_cov_kernel.cuh) with LVO = true, ring depth 4
#include "vc_cov_kernel.cuh"

VC_DEFINE_PICK(vc_cov_pick_lvo, true, 4)

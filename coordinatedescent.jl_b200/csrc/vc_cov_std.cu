// vc_cov_std.cu — instances of vc_cov_kernel (vc_cov_kernel.cuh) for locpolyl1: LVO = false, ring depth 4
#include "vc_cov_kernel.cuh"

VC_DEFINE_PICK(vc_cov_pick_std, false, 4)

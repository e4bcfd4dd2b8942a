// ffb_rd.h -- host-side entry points of the dual-tile engine's translation unit (ffb_rd.cu), called by the C ABI
// functions in ffb_kernels.cu.  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"

namespace ffb {
// bytes of global scratch the dual-tile kernels need for a field (all SMs, both groups)
FFB_HIDDEN size_t rd_scratch_bytes(int state_dim, int cond_dim);
// 0 on success; the kernels pick where the state slots / stage input live from the shared memory left beside
// the two A_lo images and the weight ring
FFB_HIDDEN int rd_launch_eval(const FieldDev& fd, const ffb_eval_args& a, cudaStream_t st);
FFB_HIDDEN int rd_launch_dopri5(const FieldDev& fd, const ffb_dopri5_args& a, cudaStream_t st);
FFB_HIDDEN int rd_launch_fixed(const FieldDev& fd, const ffb_fixed_args& a, cudaStream_t st);
}  // namespace ffb

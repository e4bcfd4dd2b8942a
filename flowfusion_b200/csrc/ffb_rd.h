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
// true when the field's stage input and conditional fit shared memory beside the two A_lo images and the weight ring
FFB_HIDDEN bool rd_dopri5_fits(const FieldDev& fd);
// 0 on success; the kernel picks where the state slots live from the shared memory that is left
FFB_HIDDEN int rd_launch_dopri5(const FieldDev& fd, const ffb_dopri5_args& a, cudaStream_t st);
}  // namespace ffb

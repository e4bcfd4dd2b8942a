// ffb_wide.h -- host-side entry points of the wide engine's translation unit (ffb_wide.cu), called by the C ABI
// functions in ffb_kernels.cu.  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"

namespace ffb {
// 0 on success.  The launchers decide where the state slots live (shared memory when the block still fits) and fail with
// FFB_ERR_ARG when the field does not fit an SM's shared memory at all.
FFB_HIDDEN int wide_launch_field_eval(FieldDev fd, const ffb_eval_args& a, cudaStream_t st);
FFB_HIDDEN int wide_launch_dopri5(FieldDev fd, const ffb_dopri5_args& a, cudaStream_t st);
FFB_HIDDEN int wide_launch_fixed(FieldDev fd, const ffb_fixed_args& a, cudaStream_t st);
// two 32-row halves per pass (EngineWideT<2, false>, SiLU): translation unit ffb_wide2.cu; `smem` and fd.slots_smem are
// chosen by the caller (ffb_wide.cu: wd_pick_smem)
FFB_HIDDEN int wide2_launch_dopri5(const FieldDev& fd, const ffb_dopri5_args& a, size_t smem, cudaStream_t st);
FFB_HIDDEN int wide2_launch_fixed(const FieldDev& fd, const ffb_fixed_args& a, size_t smem, cudaStream_t st);
}  // namespace ffb

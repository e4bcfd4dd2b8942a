// ffb_common.cuh -- what the translation units of libffb200.so share: the layout of the FFB_NPART partial sums, two
// device helpers of the dopri5 tail, and the host-side error / launch-count plumbing (defined in ffb_kernels.cu).
#pragma once
#include <string>

#include "ffb200.h"

// indices into the per-tile partial sums (ffb200.h: FFB_NPART doubles per tile)
enum { P_X_Y = 0, P_X_F = 1, P_X_DF = 2, P_X_ERR = 3, P_LP_Y = 4, P_LP_F = 5, P_LP_DF = 6, P_LP_ERR = 7,
       P_C_Y = 8, P_NONFINITE = 9 };

__device__ __forceinline__ bool is_finite_f(float x) { return fabsf(x) <= 3.402823466e38f; }

// torchdiffeq's dense-output polynomial (interp.py): coefficients then Horner-like evaluation,
// every product and sum rounded separately as the eager PyTorch ops do
__device__ __forceinline__ float dense_output(float y0, float y1, float ymid, float f0, float f1, float dt, float x) {
  const float A = __fadd_rn(__fsub_rn(__fmul_rn(2.0f * dt, __fsub_rn(f1, f0)), __fmul_rn(8.0f, __fadd_rn(y1, y0))),
                            __fmul_rn(16.0f, ymid));
  const float B = __fsub_rn(
      __fadd_rn(__fadd_rn(__fmul_rn(dt, __fsub_rn(__fmul_rn(5.0f, f0), __fmul_rn(3.0f, f1))), __fmul_rn(18.0f, y0)),
                __fmul_rn(14.0f, y1)),
      __fmul_rn(32.0f, ymid));
  const float C = __fadd_rn(
      __fsub_rn(__fsub_rn(__fmul_rn(dt, __fsub_rn(f1, __fmul_rn(4.0f, f0))), __fmul_rn(11.0f, y0)),
                __fmul_rn(5.0f, y1)),
      __fmul_rn(16.0f, ymid));
  const float D = __fmul_rn(dt, f0);
  float total = __fadd_rn(y0, __fmul_rn(x, D));
  float xp = __fmul_rn(x, x);
  total = __fadd_rn(total, __fmul_rn(xp, C));
  xp = __fmul_rn(xp, x);
  total = __fadd_rn(total, __fmul_rn(xp, B));
  xp = __fmul_rn(xp, x);
  total = __fadd_rn(total, __fmul_rn(xp, A));
  return total;
}

#define FFB_HIDDEN __attribute__((visibility("hidden")))
FFB_HIDDEN int ffb_fail(int code, const std::string& msg);   // stores the thread's ffb_last_error() text, returns code
FFB_HIDDEN void ffb_count_launches(int n);                    // ffb_launch_count()
FFB_HIDDEN int ffb_num_sms();

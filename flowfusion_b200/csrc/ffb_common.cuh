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

// idx / D for D in 1..128 and idx < 2^25 without the ~25-instruction integer division sequence
__device__ __forceinline__ int fast_div(int idx, int D) {
  return D == 1 ? idx : (int)__umulhi((unsigned)idx, (unsigned)((0x100000000ull + (unsigned)D - 1u) / (unsigned)D));
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (throughput-mode noise of the Euler-Maruyama kernel)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// 4 standard normals for (global row, step, group of 4 columns)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t offset, int64_t grow, int step, int grp) {
  const uint64_t c1 = offset + (uint64_t)(uint32_t)step;
  uint4 ctr = make_uint4((uint32_t)grow, (uint32_t)((uint64_t)grow >> 32) ^ ((uint32_t)grp << 8), (uint32_t)c1,
                         (uint32_t)(c1 >> 32));
  const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.x + 1.0f) * k, u1 = (float)r.y * k;   // (0,1], [0,1)
  const float u2 = ((float)r.z + 1.0f) * k, u3 = (float)r.w * k;
  const float r0 = sqrtf(-2.0f * __logf(fminf(u0, 1.0f))), r1 = sqrtf(-2.0f * __logf(fminf(u2, 1.0f)));
  float s0, c0, s1, c1f;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1f);
  return make_float4(r0 * c0, r0 * s0, r1 * c1f, r1 * s1);
}


// network evaluations per step of a fixed-grid method
__device__ __forceinline__ int evals_per_step(int method) {
  return method == FFB_M_RK4 ? 4 : (method == FFB_M_MIDPOINT ? 2 : (method == FFB_M_LEAPFROG ? 3 : 1));
}

#define FFB_HIDDEN __attribute__((visibility("hidden")))
FFB_HIDDEN int ffb_fail(int code, const std::string& msg);   // stores the thread's ffb_last_error() text, returns code
FFB_HIDDEN void ffb_count_launches(int n);                    // ffb_launch_count()
FFB_HIDDEN int ffb_num_sms();

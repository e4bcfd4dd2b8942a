// ffb_control.cuh -- torchdiffeq's adaptive dopri5 loop as a device-side controller.
//
// One turn = what flowfusion_b200/solver.py::dopri5 does on the host between two attempted steps
// (SURVEY.md section 8c, T3-T12): RMS / mixed error norm from the FP64 partial sums, accept test, order-5
// step-size controller, step_t clipping, FP32 stage times (with the one-ulp perturbation of the alpha = 1
// stages), the per-evaluation scalars of the next attempt (time features and SDE coefficients in the FP32 op
// order of diffusion.py:109-110, 1061, 1112, 1131, 1152-1156, 866, 885-887, 1288-1297, 1340, flow.py:112-115,
// symplectic.py:103) and the dt-scaled tableau rows.  The code is __host__ __device__: k_dopri5_control runs
// it in one thread between two attempt kernels, ffb_dopri5_control_host runs the same statements on the CPU
// for the GPU-less tests.  FP32 products and sums are rounded one by one (no FMA contraction), as eager
// PyTorch / numpy scalar ops are; sin / cos / exp / pow are the CUDA (or libm) FP32 functions, which may differ
// from PyTorch's CPU vector math in the last ulp (tests/test_gpu_control.py bounds the difference).
#pragma once
#include <math.h>
#include <stdint.h>

#include "ffb200.h"

#ifdef __CUDACC__
#define FFB_HD __host__ __device__ __forceinline__
#else
#define FFB_HD inline
#endif

namespace ffbctl {

enum { CP_X_ERR = 3, CP_LP_ERR = 7, CP_NONFINITE = 9 };   // indices into the FFB_NPART sums (ffb_kernels.cu)

FFB_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
FFB_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
FFB_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
FFB_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
FFB_HD float fsqrt(float a) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// beta(t) = beta_min + (beta_max - beta_min) * (t / T)                       diffusion.py:1061
FFB_HD float beta_of(const ffb_time_program& p, float t) { return fadd(p.beta_min, fmul(p.beta_diff, fdiv(t, p.T))); }
// lc(t) = 0.5 (beta_max - beta_min) t^2 / T + beta_min t                     diffusion.py:1152-1154
FFB_HD float log_coeff(const ffb_time_program& p, float t) {
  return fadd(fdiv(fmul(p.half_beta_diff, fmul(t, t)), p.T), fmul(p.beta_min, t));
}

// ffb_eval_scalars of one evaluation at user time t
FFB_HD void program_row(const ffb_time_program& p, float t, float sign, ffb_eval_scalars* out) {
  for (int k = 0; k < FFB_MAX_TFEAT; ++k) out->tfeat[k] = 0.0f;
  if (p.time_features == FFB_PROG_RAW_T) {
    out->tfeat[0] = t;
  } else {
    for (int k = 0; k < p.n_freq; ++k) {
      const float proj = fmul(fmul(fmul(t, p.W[k]), 2.0f), p.pi);
      out->tfeat[k] = sinf(proj);
      out->tfeat[p.n_freq + k] = cosf(proj);
    }
  }
  float a = 0.0f, g = 0.0f, sigma = 1.0f;
  if (p.sde == FFB_SDE_VP || p.sde == FFB_SDE_SUBVP) {
    const float beta = beta_of(p, t);
    a = fmul(-0.5f, beta);
    if (p.sde == FFB_SDE_VP) {
      g = fsqrt(beta);
      if (p.use_sigma) sigma = fsqrt(fsub(1.0f, expf(-log_coeff(p, t))));
    } else {
      const float inner = fsub(fmul(p.m2_beta_min, t), fdiv(fmul(p.beta_diff, fmul(t, t)), p.T));
      g = fsqrt(fmul(beta, fsub(1.0f, expf(inner))));
      if (p.use_sigma) sigma = fsub(1.0f, expf(-log_coeff(p, t)));
    }
  } else if (p.sde == FFB_SDE_VE) {
    const float s = fmul(p.sigma_min, powf(p.sigma_ratio, fdiv(t, p.T)));
    g = fmul(s, p.ve_gfac);
    if (p.use_sigma) sigma = s;
  }
  const float g2 = fmul(g, g);
  out->a = a;
  out->c = (p.sde == FFB_SDE_NONE) ? 0.0f : (p.sde_mode ? g2 : fmul(0.5f, g2));
  out->sigma = sigma;
  out->sign = sign;
}

FFB_HD double nan64() { return (double)NAN; }

// sqrt(mean) of a sum of squares, rounded to FP32 as torch's .sqrt() of an FP32 mean would be stored
FFB_HD float rms32(double sumsq, int64_t n) {
  if (n == 0) return NAN;
  return (float)sqrt(sumsq / (double)n);
}

// the attempt that starts at ctl.t with ctl.dt_next: stage times, scalars, dt-scaled tableau (solver.py loop head)
FFB_HD void prepare_attempt(const ffb_dopri5_ctl_params& p, ffb_dopri5_ctl& c) {
  if (!(c.n_attempts < p.max_num_steps)) { c.done = FFB_CTL_MAX_STEPS; return; }
  const double t = c.t, dt = c.dt_next;
  if (!(t + dt > t)) { c.done = FFB_CTL_DT_UNDERFLOW; return; }
  double t1 = t + dt, dts = dt;
  int on_grid = 0;
  if (p.n_grid > 0) {
    const double nxt = p.grid[c.grid_idx];
    on_grid = (t < nxt && nxt < t + dt) ? 1 : 0;
    if (on_grid) { t1 = nxt; dts = t1 - t; }
  }
  const float t0_32 = (float)t, dt_32 = (float)dts, t1_32 = (float)t1;
  const float sign = p.reverse ? -1.0f : 1.0f;
  for (int i = 0; i < 6; ++i) {
    const float ts = (p.alpha[i] == 1.0f) ? nextafterf(t1_32, fsub(t1_32, 1.0f)) : fadd(t0_32, fmul(p.alpha[i], dt_32));
    program_row(p.prog, p.reverse ? -ts : ts, sign, &c.ev[i]);
    for (int j = 0; j < 6; ++j) c.cb[i][j] = fmul(p.beta[i][j], dt_32);
  }
  for (int j = 0; j < 7; ++j) { c.ce[j] = fmul(dt_32, p.c_err[j]); c.cm[j] = fmul(dt_32, p.c_mid[j]); }
  const int final = !(p.t_end > t1);
  c.final = final;
  c.x_interp = final ? (float)((p.t_end - t) / (t1 - t)) : 0.0f;
  c.dt = dt_32;
  c.cur_t1 = t1; c.cur_dt = dts; c.cur_on_grid = on_grid;
}

// judge the attempt whose sums arrived, then prepare the next one (solver.py loop tail)
FFB_HD void after_attempt(const ffb_dopri5_ctl_params& p, const double* s, ffb_dopri5_ctl& c) {
  if (c.done != FFB_CTL_RUNNING) return;
  const int idx = c.n_attempts;
  c.n_attempts = idx + 1;
  if (s[CP_NONFINITE] > 0.0) { c.done = FFB_CTL_NONFINITE; return; }
  // mixed norm = Python's max() over (x, cond, lp) in tuple order; a zero-derivative conditional has zero error
  float m = rms32(s[CP_X_ERR], p.n_x);
  if (p.n_cond > 0) { const float v = rms32(0.0, p.n_cond); if (v > m) m = v; }
  if (p.n_lp > 0) { const float v = rms32(s[CP_LP_ERR], p.n_lp); if (v > m) m = v; }
  const float ratio = fabsf(m);
  const double dts = c.cur_dt;
  bool accept = ratio <= 1.0f;
  if (dts > p.max_step) accept = false;
  if (dts <= p.min_step) accept = true;
  if (idx < FFB_CTL_HIST) { c.hist_dt[idx] = dts; c.hist_ratio[idx] = ratio; c.hist_accept[idx] = accept ? 1 : 0; }
  if (accept) {
    c.n_accepted += 1;
    c.cur ^= 1;
    if (c.cur_on_grid && c.grid_idx != p.n_grid - 1) c.grid_idx += 1;
    c.t = c.cur_t1;
    if (c.final) { c.done = FFB_CTL_FINISHED; return; }
  } else {
    c.n_rejected += 1;
  }
  double nxt;
  if (ratio != ratio) {
    nxt = nan64();
  } else if (ratio == 0.0f) {
    nxt = dts * p.ifactor;
  } else {
    const double dfac = (ratio < 1.0f) ? 1.0 : p.dfactor;
    const double r = (double)ratio;
    double fac = p.safety / pow(r, 0.2);
    if (fac < dfac) fac = dfac;
    if (fac > p.ifactor) fac = p.ifactor;
    nxt = dts * fac;
  }
  if (nxt == nxt) { if (nxt < p.min_step) nxt = p.min_step; if (nxt > p.max_step) nxt = p.max_step; }
  c.dt_next = nxt;
  prepare_attempt(p, c);
}

// tell the host which turn has been taken and whether the solve goes on (pinned host memory, no stream operation)
FFB_HD void notify_host(ffb_dopri5_ctl& c) {
  c.n_turns += 1;
  if (c.notify) {
    volatile int32_t* slot = c.notify + ((c.n_turns - 1) & (FFB_CTL_NOTIFY_SLOTS - 1));
    *slot = (int32_t)(((uint32_t)c.n_turns << 8) | ((uint32_t)c.done & 0xffu));
#ifdef __CUDA_ARCH__
    __threadfence_system();
#endif
  }
}

FFB_HD void control_turn(const ffb_dopri5_ctl_params& p, const double* sums, ffb_dopri5_ctl& c, int after) {
  if (after) { after_attempt(p, sums, c); notify_host(c); }
  else if (c.done == FFB_CTL_RUNNING) prepare_attempt(p, c);
}

}  // namespace ffbctl

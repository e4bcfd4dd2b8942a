// ffb_kernels.cu -- integrator kernels and the C ABI of libffb200.so (see include/ffb200.h).
//
// Kernels (all persistent: grid = min(#tiles, #SMs), one 288-thread CTA per SM):
//   k_field_eval   one evaluation of the vector field (+ torchdiffeq initial-step norms)
//   k_dopri5       one attempted Dormand-Prince step: 6 fused evaluations, FSAL, error norm
//                  partials in FP64, dense-output interpolant at t_end
//   k_fixed        whole fixed-grid trajectory on-chip: euler / midpoint / rk4(3/8) /
//                  Euler-Maruyama / leapfrog
//   k_reduce, k_gauss_logprob, k_philox, k_ffma_peak, k_pack_*   small helpers
//
// Reference semantics followed (paths relative to the reference checkout):
//   torchdiffeq dopri5 / fixed-grid drivers  -> restated in oracle/torchdiffeq/_solver.py
//   diffusion.py:258-279 (PF-ODE drift), :543-562 (Euler-Maruyama), :483-503 / :327-334 (trace)
//   flow.py:109-166, :553-652 ; symplectic.py:99-123, :186-201
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"
#include "ffb_engine_tc.cuh"
#include "ffb_control.cuh"
#include "ffb_rd.h"

using namespace ffb;

// =============================================================================================
// small device helpers
// =============================================================================================

// k-major tile buffer <- row-major global rows [row0, row0+nv); rows nv..S-1 are zero filled
__device__ __forceinline__ void load_rows(float* dst, const float* __restrict__ src, int64_t row0, int nv, int S,
                                          int D, int tid) {
  const float* __restrict__ base = src + row0 * D;
#pragma unroll 4
  for (int idx = tid; idx < S * D; idx += NCOMP) {
    const int r = fast_div(idx, D), d = idx - r * D;
    dst[d * LDA + r] = (r < nv) ? base[idx] : 0.0f;
  }
}
__device__ __forceinline__ void store_rows(float* __restrict__ dst, const float* src, int64_t row0, int nv, int D,
                                           int tid) {
  float* __restrict__ base = dst + row0 * D;
#pragma unroll 4
  for (int idx = tid; idx < nv * D; idx += NCOMP) {
    const int r = fast_div(idx, D), d = idx - r * D;
    base[idx] = src[d * LDA + r];
  }
}

// deterministic block reduction of NV doubles per thread -> out[q] (thread 0 writes)
template <int NV, class CTX>
__device__ __forceinline__ void block_reduce_store(CTX& cx, double (&v)[NV], double* out, const int (&slot)[NV]) {
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
  }
  if (cx.lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) cx.red()[cx.warp * FFB_NPART + q] = v[q];
  }
  bar_compute();
  if (cx.tid == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double s = 0.0;
      for (int w = 0; w < NCOMP / 32; ++w) s += cx.red()[w * FFB_NPART + q];
      out[slot[q]] = s;
    }
  }
  bar_compute();
}

// =============================================================================================
// k_field_eval
// =============================================================================================
template <class ENG, bool SS>
__global__ void __launch_bounds__(ENG::NTHR, 1) k_field_eval(const __grid_constant__ FieldDev f,
        const __grid_constant__ ffb_eval_args a, const int64_t ntiles) {
  typename ENG::Ctx cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch));
  const int S = cx.S, SD = cx.SD, CD = cx.CD;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * S;
    const int nv = (int)min((int64_t)S, a.batch - row0);
    float* Y0 = slot_ptr_t<SS>(cx, SLOT_Y0);
    float* FB = slot_ptr_t<SS>(cx, 1);
    if (!cx.producer) {
      load_rows(Y0, a.y, row0, nv, S, SD, cx.tid);
      if (a.fbase) load_rows(FB, a.fbase, row0, nv, S, SD, cx.tid);
      if (CD) load_rows(cx.condb(), a.cond, row0, nv, S, CD, cx.tid);
      if (f.div_mode == FFB_DIV_HUTCH) load_rows(cx.prb(), a.probes, row0, nv, S, SD, cx.tid);
      bar_compute();
      for (int d = cx.tid >> 7, r = cx.tid & (TM - 1); d < SD && r < S; d += NCOMP / TM) {
        const int e = d * LDA + r;
        cx.ycur()[e] = a.fbase ? __fadd_rn(Y0[e], __fmul_rn(a.h, FB[e])) : Y0[e];
      }
      bar_compute();
    }
    ENG::template eval<SS>(cx, f, a.ev, 0);
    if (!cx.producer) {
      const float* F = slot_ptr_t<SS>(cx, 0);
      if (a.f) store_rows(a.f, F, row0, nv, SD, cx.tid);
      if (a.dlp && cx.T > 0)
        for (int s = cx.tid; s < nv; s += NCOMP) a.dlp[row0 + s] = cx.klp()[s];
      if (a.norms) {
        double v[6] = {0, 0, 0, 0, 0, 0};   // x_y, x_f, x_df, lp_f, lp_df, c_y
        for (int d = cx.tid >> 7, r = cx.tid & (TM - 1); d < SD && r < S; d += NCOMP / TM) {
          if (r >= nv) continue;
          const int e = d * LDA + r;
          const float y0 = Y0[e];
          const float sc = __fadd_rn(a.atol, __fmul_rn(fabsf(y0), a.rtol));
          if (a.norms == 1) {
            const float q0 = __fdiv_rn(y0, sc), q1 = __fdiv_rn(F[e], sc);
            v[0] += (double)q0 * q0;
            v[1] += (double)q1 * q1;
          } else {
            const float q2 = __fdiv_rn(__fsub_rn(F[e], FB[e]), sc);
            v[2] += (double)q2 * q2;
          }
        }
        if (cx.T > 0) {
          for (int s = cx.tid; s < nv; s += NCOMP) {
            if (a.norms == 1) {
              const float q = __fdiv_rn(cx.klp()[s], a.atol);
              v[3] += (double)q * q;
            } else {
              const float q = __fdiv_rn(__fsub_rn(cx.klp()[s], a.dlpbase[row0 + s]), a.atol);
              v[4] += (double)q * q;
            }
          }
        }
        if (a.cond_in_state && a.norms == 1) {
          const float* cs = a.cond_state ? a.cond_state : a.cond;
          for (int idx = cx.tid; idx < CD * nv; idx += NCOMP) {
            const float c = cs[row0 * CD + idx];
            const float q = __fdiv_rn(c, __fadd_rn(a.atol, __fmul_rn(fabsf(c), a.rtol)));
            v[5] += (double)q * q;
          }
        }
        const int slot[6] = {P_X_Y, P_X_F, P_X_DF, P_LP_F, P_LP_DF, P_C_Y};
        block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot);
      }
      bar_compute();
    }
  }
  ENG::fini(cx);
}

// =============================================================================================
// k_dopri5: one attempted step
// =============================================================================================
template <class ENG, bool SS>
__global__ void __launch_bounds__(ENG::NTHR, 1) k_dopri5(const __grid_constant__ FieldDev f,
        const __grid_constant__ ffb_dopri5_args a, const int64_t ntiles) {
  typename ENG::Ctx cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch));
  const int S = cx.S, SD = cx.SD, CD = cx.CD;
  const bool prob = cx.T > 0;
  float* LP0 = cx.klp() + NSLOT * TM;
  float* LPC = cx.klp() + (NSLOT + 1) * TM;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * S;
    const int nv = (int)min((int64_t)S, a.batch - row0);
    float* Y0 = slot_ptr_t<SS>(cx, SLOT_Y0);
    double nonfinite = 0.0;
    if (!cx.producer) {
      load_rows(Y0, a.y0, row0, nv, S, SD, cx.tid);
      load_rows(slot_ptr_t<SS>(cx, 0), a.f0, row0, nv, S, SD, cx.tid);
      if (CD) load_rows(cx.condb(), a.cond, row0, nv, S, CD, cx.tid);
      if (f.div_mode == FFB_DIV_HUTCH) load_rows(cx.prb(), a.probes, row0, nv, S, SD, cx.tid);
      if (prob)
        for (int s = cx.tid; s < S; s += NCOMP) {
          LP0[s] = (s < nv) ? a.lp0[row0 + s] : 0.0f;
          cx.klp()[s] = (s < nv) ? a.dlp0[row0 + s] : 0.0f;
          if (!is_finite_f(LP0[s])) nonfinite += 1.0;
        }
      bar_compute();
    }
    for (int i = 1; i <= 6; ++i) {
      if (!cx.producer) {
        for (int d = cx.tid >> 7, r = cx.tid & (TM - 1); d < SD && r < S; d += NCOMP / TM) {
          const int e = d * LDA + r;
          float kv[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) kv[j] = (j < i) ? slot_ptr_t<SS>(cx, j)[e] : 0.0f;
          const float y0 = Y0[e];
          float acc = __fmul_rn(kv[0], a.cb[i - 1][0]);
#pragma unroll
          for (int j = 1; j < 6; ++j) if (j < i) acc = fmaf(kv[j], a.cb[i - 1][j], acc);
          if (i == 1 && !is_finite_f(y0)) nonfinite += 1.0;
          cx.ycur()[e] = __fadd_rn(y0, acc);
        }
        bar_compute();
      }
      ENG::template eval<SS>(cx, f, a.ev[i - 1], i);
    }
    if (!cx.producer) {
      // cx.ycur() now holds y1 (FSAL: the 7th stage input), slot 6 holds f1
      double v[3] = {0.0, 0.0, nonfinite};
      float* stage_out = cx.stage_buf();   // free between evaluations: staging for the interpolant
      for (int d = cx.tid >> 7, r = cx.tid & (TM - 1); d < SD && r < S; d += NCOMP / TM) {
        if (r >= nv) continue;
        const int e = d * LDA + r;
        const float y0 = Y0[e], y1 = cx.ycur()[e];
        float kv[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) kv[j] = slot_ptr_t<SS>(cx, j)[e];
        float err = __fmul_rn(kv[0], a.ce[0]);
#pragma unroll
        for (int j = 1; j < 7; ++j) err = fmaf(kv[j], a.ce[j], err);
        const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(y0), fabsf(y1))));
        const float q = __fdiv_rn(err, tol);
        v[0] += (double)q * q;
        if (a.final) {
          float mid = __fmul_rn(kv[0], a.cm[0]);
#pragma unroll
          for (int j = 1; j < 7; ++j) mid = fmaf(kv[j], a.cm[j], mid);
          stage_out[e] = dense_output(y0, y1, __fadd_rn(y0, mid), kv[0], kv[6], a.dt, a.x_interp);
        }
      }
      if (prob) {
        for (int s = cx.tid; s < nv; s += NCOMP) {
          const float l0 = LP0[s];
          float acc = __fmul_rn(cx.klp()[s], a.cb[5][0]);
          for (int j = 1; j < 6; ++j) acc = fmaf(cx.klp()[j * TM + s], a.cb[5][j], acc);
          const float l1 = __fadd_rn(l0, acc);
          float err = __fmul_rn(cx.klp()[s], a.ce[0]);
          for (int j = 1; j < 7; ++j) err = fmaf(cx.klp()[j * TM + s], a.ce[j], err);
          const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(l0), fabsf(l1))));
          const float q = __fdiv_rn(err, tol);
          v[1] += (double)q * q;
          a.lp1[row0 + s] = l1;
          a.dlp1[row0 + s] = cx.klp()[6 * TM + s];
          if (a.final) {
            float mid = __fmul_rn(cx.klp()[s], a.cm[0]);
            for (int j = 1; j < 7; ++j) mid = fmaf(cx.klp()[j * TM + s], a.cm[j], mid);
            a.lp_out[row0 + s] = dense_output(l0, l1, __fadd_rn(l0, mid), cx.klp()[s], cx.klp()[6 * TM + s], a.dt,
                                              a.x_interp);
          }
        }
      }
      (void)LPC;
      bar_compute();
      store_rows(a.y1, cx.ycur(), row0, nv, SD, cx.tid);
      store_rows(a.f1, slot_ptr_t<SS>(cx, 6), row0, nv, SD, cx.tid);
      if (a.final) store_rows(a.y_out, stage_out, row0, nv, SD, cx.tid);
      const int slot[3] = {P_X_ERR, P_LP_ERR, P_NONFINITE};
      block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot);
    }
  }
  ENG::fini(cx);
}

// =============================================================================================
// k_fixed: fixed-grid integrators, the whole trajectory of a tile on-chip
// =============================================================================================

template <class ENG, bool SS>
__global__ void __launch_bounds__(ENG::NTHR, 1) k_fixed(const __grid_constant__ FieldDev f,
        const __grid_constant__ ffb_fixed_args a, const int64_t ntiles) {
  typename ENG::Ctx cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch));
  const int S = cx.S, SD = cx.SD, CD = cx.CD;
  const bool prob = cx.T > 0;
  const int nev = evals_per_step(a.method);
  const float third = (float)(1.0 / 3.0);
  float* LPC = cx.klp() + (NSLOT + 1) * TM;  // running lp
  // column ranges produced by the two networks of a symplectic field (leapfrog only)
  const int q_lo = f.out_off[0], q_hi = f.out_off[0] + f.net[0].N[f.net[0].n_layers - 1];
  const int p_lo = f.out_off[1], p_hi = f.out_off[1] + f.net[1].N[f.net[1].n_layers > 0 ? f.net[1].n_layers - 1 : 0];
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * S;
    const int nv = (int)min((int64_t)S, a.batch - row0);
    float* Y0 = slot_ptr_t<SS>(cx, SLOT_Y0);
    const float* K1 = slot_ptr_t<SS>(cx, 0);
    const float* K2 = slot_ptr_t<SS>(cx, 1);
    const float* K3 = slot_ptr_t<SS>(cx, 2);
    const float* K4 = slot_ptr_t<SS>(cx, 3);
    int nan_step = 0x7fffffff;            // first Euler-Maruyama step of this thread's rows that produced a NaN
    if (!cx.producer) {
      load_rows(cx.ycur(), a.x0, row0, nv, S, SD, cx.tid);
      if (CD) load_rows(cx.condb(), a.cond, row0, nv, S, CD, cx.tid);
      if (f.div_mode == FFB_DIV_HUTCH) load_rows(cx.prb(), a.probes, row0, nv, S, SD, cx.tid);
      if (prob)
        for (int s = cx.tid; s < S; s += NCOMP) LPC[s] = (s < nv && a.lp0) ? a.lp0[row0 + s] : 0.0f;
      bar_compute();
    }
    for (int step = 0; step < a.nsteps; ++step) {
      const float* st = a.step_table + (size_t)step * FFB_STEP_STRIDE;
      const ffb_eval_scalars* ev = a.ev_table + (size_t)step * nev;
      const float dt = st[0], half = st[3];
      for (int e = 0; e < nev; ++e) {
        // ---- which networks this evaluation runs, and where the derivative goes -----------------
        unsigned mask = 3u;
        int dst = e;
        if (a.method == FFB_M_LEAPFROG) {         // e0: dp/dt(q, t0) [first step only], e1: dq/dt, e2: dp/dt
          if (e == 0 && step > 0) mask = 0u;      // reuse the previous step's closing kick
          else mask = (e == 1) ? 1u : 2u;
          dst = (e == 1) ? 0 : 1;
        }
        if (mask) ENG::template eval<SS>(cx, f, ev[e], dst, mask);
        if (cx.producer) continue;
        // ---- stage algebra after evaluation e (op order of torchdiffeq's step functions) ---------
        for (int d = cx.tid >> 7, r = cx.tid & (TM - 1); d < SD && r < S; d += NCOMP / TM) {
          const int i = d * LDA + r;
          float* y = cx.ycur();
          switch (a.method) {
            case FFB_M_EULER:
              y[i] = __fadd_rn(y[i], __fmul_rn(dt, K1[i]));
              break;
            case FFB_M_MIDPOINT:
              if (e == 0) { const float y0 = y[i]; Y0[i] = y0; y[i] = __fadd_rn(y0, __fmul_rn(K1[i], half)); }
              else y[i] = __fadd_rn(Y0[i], __fmul_rn(dt, K2[i]));
              break;
            case FFB_M_RK4:
              if (e == 0) { const float y0 = y[i]; Y0[i] = y0; y[i] = __fadd_rn(y0, __fmul_rn(__fmul_rn(dt, K1[i]), third)); }
              else if (e == 1) y[i] = __fadd_rn(Y0[i], __fmul_rn(dt, __fsub_rn(K2[i], __fmul_rn(K1[i], third))));
              else if (e == 2) y[i] = __fadd_rn(Y0[i], __fmul_rn(dt, __fadd_rn(__fsub_rn(K1[i], K2[i]), K3[i])));
              else {
                const float sum = __fadd_rn(__fadd_rn(K1[i], __fmul_rn(3.0f, __fadd_rn(K2[i], K3[i]))), K4[i]);
                y[i] = __fadd_rn(Y0[i], __fmul_rn(__fmul_rn(sum, dt), 0.125f));
              }
              break;
            case FFB_M_LEAPFROG:
              if (e == 1) { if (d >= q_lo && d < q_hi) y[i] = __fadd_rn(y[i], __fmul_rn(dt, K1[i])); }
              else if (d >= p_lo && d < p_hi) y[i] = __fadd_rn(y[i], __fmul_rn(half, K2[i]));
              break;
            default: break;   // EM handled below (its noise is indexed row-major)
          }
        }
        if (prob) {
          for (int s = cx.tid; s < S; s += NCOMP) {
            const float* kl = cx.klp();
            if (a.method == FFB_M_EULER) LPC[s] = __fadd_rn(LPC[s], __fmul_rn(dt, kl[s]));
            else if (a.method == FFB_M_MIDPOINT && e == 1) LPC[s] = __fadd_rn(LPC[s], __fmul_rn(dt, kl[TM + s]));
            else if (a.method == FFB_M_RK4 && e == 3) {
              const float sum = __fadd_rn(__fadd_rn(kl[s], __fmul_rn(3.0f, __fadd_rn(kl[TM + s], kl[2 * TM + s]))), kl[3 * TM + s]);
              LPC[s] = __fadd_rn(LPC[s], __fmul_rn(__fmul_rn(sum, dt), 0.125f));
            }
          }
        }
        if (a.method == FFB_M_EM) {
          // diffusion.py:552-559: f = drift - g^2 score (ev.c = g^2); x_mean = x + f dt; x = x_mean + g dw
          const float g = st[1], sq = st[2];
          float* y = cx.ycur();
          if (a.noise) {
            for (int idx = cx.tid; idx < SD * S; idx += NCOMP) {
              const int r = fast_div(idx, SD), d = idx - r * SD;
              if (r >= nv) continue;
              const int i = d * LDA + r;
              const float xm = __fadd_rn(y[i], __fmul_rn(K1[i], dt));
              const float dw = __fmul_rn(a.noise[((size_t)step * a.batch + row0 + r) * SD + d], sq);
              const float xn = __fadd_rn(xm, __fmul_rn(g, dw));
              Y0[i] = xm;
              y[i] = xn;
              if (xn != xn) nan_step = min(nan_step, step);
            }
          } else {
            const int ng = (SD + 3) >> 2;
            for (int idx = cx.tid; idx < ng * S; idx += NCOMP) {
              const int grp = fast_div(idx, S), r = idx - grp * S;
              if (r >= nv) continue;
              const float4 z = philox_normal4(a.philox_seed, a.philox_offset, a.row_offset + row0 + r, step, grp);
              const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int d = grp * 4 + q;
                if (d >= SD) break;
                const int i = d * LDA + r;
                const float xm = __fadd_rn(y[i], __fmul_rn(K1[i], dt));
                const float xn = __fadd_rn(xm, __fmul_rn(g, __fmul_rn(zz[q], sq)));
                Y0[i] = xm;
                y[i] = xn;
                if (xn != xn) nan_step = min(nan_step, step);
              }
            }
          }
        }
        bar_compute();
      }
    }
    if (!cx.producer) {
      store_rows(a.x_out, (a.method == FFB_M_EM) ? Y0 : cx.ycur(), row0, nv, SD, cx.tid);
      if (prob && a.lp_out)
        for (int s = cx.tid; s < nv; s += NCOMP) a.lp_out[row0 + s] = LPC[s];
      if (nan_step != 0x7fffffff) { atomicOr(a.status, FFB_ST_NAN_SAMPLE); atomicMin(a.status + 1, nan_step); }
      bar_compute();
    }
  }
  ENG::fini(cx);
}

#include "ffb_kernels_rr.cuh"
#include "ffb_engine_rrt.cuh"

// =============================================================================================
// helpers
// =============================================================================================
// sums[q] = sum over tiles of partials[tile][q], in a fixed order (deterministic).  Thread i owns quantity q = i & 15
// of the tiles (i >> 4) + 64 k: a warp reads two whole 128-byte rows per load and keeps 8 loads in flight (the former
// one-warp-per-quantity loop was a chain of dependent L2 reads: 80 us for the 7 813 tiles of 1 M rows).
__device__ __forceinline__ void reduce_partials_block(const double* __restrict__ partials, int64_t ntiles, double* sh /*[1024]*/,
                                                      double* out /*[FFB_NPART], shared or global*/) {
  const int q = threadIdx.x & (FFB_NPART - 1);
  const int64_t r = threadIdx.x >> 4;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t t = r;
  for (; t + 7 * 64 < ntiles; t += 8 * 64) {
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] += partials[(t + 64 * u) * FFB_NPART + q];
  }
  for (; t < ntiles; t += 64) s[0] += partials[t * FFB_NPART + q];
  sh[threadIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (threadIdx.x < FFB_NPART) {
    double a = 0.0;
    for (int k = 0; k < 64; ++k) a += sh[k * FFB_NPART + threadIdx.x];
    out[threadIdx.x] = a;
  }
}
__global__ void __launch_bounds__(1024) k_reduce(const double* __restrict__ partials, int64_t ntiles, double* __restrict__ sums) {
  __shared__ double sh[1024];
  reduce_partials_block(partials, ntiles, sh, sums);
}

__global__ void k_gauss_logprob(const float* __restrict__ x, const float* __restrict__ add, float* __restrict__ out,
                                int64_t batch, int dim, float sigma, float log_norm) {
  // one warp per row: sum_d(-(x^2)/(2 sigma^2) - log sigma - 0.5 log 2pi)  (torch Normal.log_prob order)
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= batch) return;
  const float var2 = 2.0f * sigma * sigma;
  float s = 0.0f;
  for (int d = lane; d < dim; d += 32) {
    const float v = x[row * dim + d];
    s += -(v * v) / var2 - log_norm;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s + (add ? add[row] : 0.0f);
}

__global__ void k_philox(float* __restrict__ out, int64_t batch, int dim, uint64_t seed, uint64_t offset, int step,
                         int64_t row_offset) {
  const int ng = (dim + 3) >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * ng) return;
  const int64_t r = idx / ng;
  const int grp = (int)(idx - r * ng);
  const float4 z = philox_normal4(seed, offset, row_offset + r, step, grp);
  const float zz[4] = {z.x, z.y, z.z, z.w};
  for (int q = 0; q < 4; ++q)
    if (grp * 4 + q < dim) out[r * dim + grp * 4 + q] = zz[q];
}

__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters) {
  float2 acc[16];
  const float2 a = make_float2(1.0f + 1e-7f * threadIdx.x, 1.0f - 1e-7f * threadIdx.x);
  const float2 b = make_float2(1e-9f, -1e-9f);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2((float)i, (float)-i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;
}

// weight packing -------------------------------------------------------------------------------
// real column n of packed column p for a layer with padded width Np
__host__ __device__ inline int packed_to_real(int p, int Np) {
  const int C = Np / 16;
  if (C == 8) {
    const int g = p / 64, rem = p % 64, tx = rem / 4, jj = rem % 4;
    return (g * 4 + jj) * 16 + tx;
  }
  const int tx = p / C, j = p % C;
  return j * 16 + tx;
}
// dst[k][p] = W[n(p)][src_col(k)], zero outside; rows_map: k -> source column (or -1)
__global__ void k_pack_weight(const float* __restrict__ W, int in_features, int out_features, float* __restrict__ dst,
                              int K, int Np, int x_col, int x_dim, int c_col, int c_dim, int layer0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * Np) return;
  const int k = idx / Np, p = idx - k * Np;
  const int n = packed_to_real(p, Np);
  int col = -1;
  if (layer0) {
    if (k < x_dim) col = x_col + k;
    else if (k < x_dim + c_dim) col = c_col + (k - x_dim);
  } else if (k < in_features) {
    col = k;
  }
  dst[idx] = (n < out_features && col >= 0) ? W[(size_t)n * in_features + col] : 0.0f;
}
__global__ void k_pack_time(const float* __restrict__ W, int in_features, int out_features, float* __restrict__ dst,
                            int t_dim, int Np, int t_col) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= t_dim * Np) return;
  const int j = idx / Np, p = idx - j * Np;
  const int n = packed_to_real(p, Np);
  dst[idx] = (n < out_features) ? W[(size_t)n * in_features + t_col + j] : 0.0f;
}
__global__ void k_pack_bias(const float* __restrict__ b, int out_features, float* __restrict__ dst, int Np) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np) return;
  const int n = packed_to_real(p, Np);
  dst[p] = (n < out_features) ? b[n] : 0.0f;
}

// ---- tensor-core image: per chunk of 32 k-rows [W_hi | W_lo], each in the canonical no-swizzle
// K-major UMMA layout: element (n, k) at kc*(Np*16 B) + (n/8)*128 B + (n%8)*16 B + (k%4)*4 B, kc = k/4
__global__ void k_pack_weight_tc(const float* __restrict__ W, int in_features, int out_features,
                                 float* __restrict__ dst, int K, int Np, int x_col, int x_dim, int c_col, int c_dim,
                                 int layer0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * Np) return;
  const int k = idx / Np, n = idx - k * Np;
  int col = -1;
  if (layer0) {
    if (k < x_dim) col = x_col + k;
    else if (k < x_dim + c_dim) col = c_col + (k - x_dim);
  } else if (k < in_features) {
    col = k;
  }
  const float w = (n < out_features && col >= 0) ? W[(size_t)n * in_features + col] : 0.0f;
  uint32_t hi, lo;
  tf32_split_rn(w, hi, lo);
  const int c = k / KC, kk = k - c * KC, rows = min(KC, K - c * KC);
  const size_t base = (size_t)c * 2 * KC * Np;
  const size_t off = (size_t)(kk >> 2) * (Np * 4) + (size_t)(n >> 3) * 32 + (size_t)(n & 7) * 4 + (kk & 3);
  dst[base + off] = __uint_as_float(hi);
  dst[base + (size_t)rows * Np + off] = __uint_as_float(lo);
}
__global__ void k_pack_time_tc(const float* __restrict__ W, int in_features, int out_features, float* __restrict__ dst,
                               int t_dim, int Np, int t_col) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= t_dim * Np) return;
  const int j = idx / Np, n = idx - j * Np;
  dst[idx] = (n < out_features) ? W[(size_t)n * in_features + t_col + j] : 0.0f;
}
__global__ void k_pack_bias_tc(const float* __restrict__ b, int out_features, float* __restrict__ dst, int Np) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Np) return;
  dst[n] = (n < out_features) ? b[n] : 0.0f;
}

// =============================================================================================
// device-side dopri5 controller (ffb_control.cuh): one thread, between two attempt kernels
// =============================================================================================
__global__ void k_dopri5_control(const __grid_constant__ ffb_dopri5_ctl_params p, const double* __restrict__ sums,
                                 ffb_dopri5_ctl* __restrict__ ctl, int after) {
  if (threadIdx.x == 0 && blockIdx.x == 0) ffbctl::control_turn(p, sums, *ctl, after);
}
// single-GPU shortcut: the tile reduction and the controller turn in one launch (sums is still written, for the caller)
__global__ void __launch_bounds__(1024) k_dopri5_reduce_control(const __grid_constant__ ffb_dopri5_ctl_params p,
        const double* __restrict__ partials, int64_t ntiles, double* __restrict__ sums, ffb_dopri5_ctl* __restrict__ ctl) {
  __shared__ double sh[1024];
  __shared__ double tot[FFB_NPART];
  if (ctl->done != FFB_CTL_RUNNING) {                  // uniform: nothing in flight to judge, only the turn is counted
    if (threadIdx.x == 0) ffbctl::notify_host(*ctl);
    return;
  }
  reduce_partials_block(partials, ntiles, sh, tot);
  __syncthreads();
  if (threadIdx.x < FFB_NPART) sums[threadIdx.x] = tot[threadIdx.x];
  if (threadIdx.x == 0) ffbctl::control_turn(p, tot, *ctl, 1);
}
__global__ void k_time_program(const __grid_constant__ ffb_time_program prog, const float* __restrict__ times, int n,
                               float sign, ffb_eval_scalars* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ffbctl::program_row(prog, times[i], sign, out + i);
}

// =============================================================================================
// host side: C ABI
// =============================================================================================
struct ffb_net {
  NetDev dev;   // FP32 FFMA2 engine image
  NetDev tc;    // tensor-core (3xTF32) engine image
  std::vector<void*> allocs;
  int64_t flops;
};

// engine selection: 1 = tensor cores (default: dual-tile engine for fields without tangent rows, tangent-row engine
// for the log-likelihood paths), 3 = the same with the single-tile chunk-pipelined engine instead of the dual-tile one
// (FFB_ENGINE=rr), 4 = the same with the dual-tile engine for every dopri5 attempt it can hold, whatever the batch
// (FFB_ENGINE=rd; tests), 2 = the older whole-layer hand-off tile engine (FFB_ENGINE=tc_tile), 0 = FP32 FFMA2 (FFB_ENGINE=ffma)
static int g_engine = -1;
static int engine() {
  if (g_engine < 0) {
    const char* e = getenv("FFB_ENGINE");
    if (e && (!strcmp(e, "ffma") || !strcmp(e, "0"))) g_engine = 0;
    else if (e && (!strcmp(e, "tc_tile") || !strcmp(e, "2"))) g_engine = 2;
    else if (e && (!strcmp(e, "rr") || !strcmp(e, "3"))) g_engine = 3;
    else if (e && (!strcmp(e, "rd") || !strcmp(e, "4"))) g_engine = 4;
    else g_engine = 1;
  }
  return g_engine;
}
static bool chunk_engines() { return engine() == 1 || engine() == 3 || engine() == 4; }
// debug: timeline trace buffer (2 * 4096 int64) or NULL to disable
extern "C" int ffb_debug_trace(long long* buf) {
  int zero = 0;
  cudaMemcpyToSymbol(ffb::g_trace, &buf, sizeof(buf));
  cudaMemcpyToSymbol(ffb::g_trace_pos, &zero, sizeof(zero));
  return 0;
}
extern "C" int ffb_set_engine(int e) { g_engine = (e == 2 || e == 3 || e == 4) ? e : (e ? 1 : 0); return g_engine; }
extern "C" int ffb_get_engine(void) { return engine(); }

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return fail(FFB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
  } while (0)

// helpers shared with the other translation unit (csrc/ffb_staged.cu), declared in ffb_common.cuh
int ffb_fail(int code, const std::string& msg) { return fail(code, msg); }
void ffb_count_launches(int n) { g_launches += n; }

extern "C" int ffb_abi_version(void) { return FFB_ABI_VERSION; }
extern "C" const char* ffb_last_error(void) { return g_err.c_str(); }
extern "C" int64_t ffb_launch_count(void) { return g_launches.load(); }

extern "C" int ffb_device_info(int32_t* sm_count, int32_t* smem_optin, int32_t* cc_major, int32_t* cc_minor,
                               int32_t* clock_khz) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  int v = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem_optin) *smem_optin = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major) *cc_major = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
  if (cc_minor) *cc_minor = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev));
  if (clock_khz) *clock_khz = v;
  return FFB_OK;
}

static int pad_width(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }

extern "C" int ffb_net_create(const ffb_net_desc* d, void* stream_, ffb_net** out) {
  if (!d || !out) return fail(FFB_ERR_ARG, "ffb_net_create: null argument");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (d->n_layers < 1 || d->n_layers > FFB_MAX_LAYERS) return fail(FFB_ERR_ARG, "ffb_net_create: 1..8 Linear layers supported");
  if (d->t_dim < 0 || d->t_dim > FFB_MAX_TFEAT) return fail(FFB_ERR_ARG, "ffb_net_create: at most 32 time-feature columns");
  if (d->x_dim < 1 || d->x_dim + d->c_dim + d->t_dim != d->in_features)
    return fail(FFB_ERR_ARG, "ffb_net_create: x_dim + c_dim + t_dim must equal in_features");
  if (d->x_dim + d->c_dim > FFB_MAX_WIDTH) return fail(FFB_ERR_ARG, "ffb_net_create: state + conditional columns exceed 128");
  for (int l = 0; l < d->n_layers; ++l)
    if (d->widths[l] < 1 || d->widths[l] > FFB_MAX_WIDTH)
      return fail(FFB_ERR_ARG, "ffb_net_create: layer widths must be in 1..128");
  if (d->activation < FFB_ACT_SILU || d->activation > FFB_ACT_GELU) return fail(FFB_ERR_ARG, "ffb_net_create: unknown activation");
  ffb_net* net = new ffb_net();
  NetDev& nd = net->dev;
  memset(&nd, 0, sizeof(nd));
  nd.n_layers = d->n_layers;
  nd.t_dim = d->t_dim; nd.x_dim = d->x_dim; nd.c_dim = d->c_dim;
  nd.act = d->activation;
  net->flops = 0;
  int in_f = d->in_features;
  auto cleanup = [&]() { for (void* p : net->allocs) cudaFree(p); delete net; };
  for (int l = 0; l < d->n_layers; ++l) {
    const int N = d->widths[l], Np = pad_width(N);
    const int K = (l == 0) ? ((d->x_dim + d->c_dim + 3) & ~3) : nd.Np[l - 1];
    nd.K[l] = K; nd.N[l] = N; nd.Np[l] = Np;
    net->flops += 2LL * in_f * N;
    float *w = nullptr, *b = nullptr;
    if (cudaMalloc(&w, sizeof(float) * K * Np) != cudaSuccess || cudaMalloc(&b, sizeof(float) * Np) != cudaSuccess) {
      cleanup();
      return fail(FFB_ERR_CUDA, "ffb_net_create: cudaMalloc failed");
    }
    net->allocs.push_back(w); net->allocs.push_back(b);
    k_pack_weight<<<(K * Np + 255) / 256, 256, 0, stream>>>(d->weight[l], in_f, N, w, K, Np, d->x_col, d->x_dim,
                                                           d->c_col, d->c_dim, l == 0);
    k_pack_bias<<<1, 128, 0, stream>>>(d->bias[l], N, b, Np);
    g_launches += 2;
    nd.W[l] = w; nd.b[l] = b;
    if (l == 0) {
      float* wt = nullptr;
      const int td = d->t_dim > 0 ? d->t_dim : 1;
      if (cudaMalloc(&wt, sizeof(float) * td * Np) != cudaSuccess) { cleanup(); return fail(FFB_ERR_CUDA, "cudaMalloc"); }
      net->allocs.push_back(wt);
      if (d->t_dim > 0) {
        k_pack_time<<<(d->t_dim * Np + 255) / 256, 256, 0, stream>>>(d->weight[0], in_f, N, wt, d->t_dim, Np, d->t_col);
        g_launches += 1;
      }
      nd.Wt = wt;
    }
    in_f = N;
  }
  // ---- tensor-core image ---------------------------------------------------------------------
  NetDev& nt = net->tc;
  memset(&nt, 0, sizeof(nt));
  nt.n_layers = d->n_layers;
  nt.t_dim = d->t_dim; nt.x_dim = d->x_dim; nt.c_dim = d->c_dim;
  nt.act = d->activation;
  in_f = d->in_features;
  for (int l = 0; l < d->n_layers; ++l) {
    const int N = d->widths[l], Np = (N + 31) & ~31;
    const int K = (l == 0) ? ((d->x_dim + d->c_dim + 7) & ~7) : nt.Np[l - 1];
    nt.K[l] = K; nt.N[l] = N; nt.Np[l] = Np;
    float *w = nullptr, *b = nullptr;
    if (cudaMalloc(&w, sizeof(float) * 2 * K * Np) != cudaSuccess || cudaMalloc(&b, sizeof(float) * Np) != cudaSuccess) {
      cleanup();
      return fail(FFB_ERR_CUDA, "ffb_net_create: cudaMalloc failed");
    }
    net->allocs.push_back(w); net->allocs.push_back(b);
    k_pack_weight_tc<<<(K * Np + 255) / 256, 256, 0, stream>>>(d->weight[l], in_f, N, w, K, Np, d->x_col, d->x_dim,
                                                              d->c_col, d->c_dim, l == 0);
    k_pack_bias_tc<<<1, 128, 0, stream>>>(d->bias[l], N, b, Np);
    g_launches += 2;
    nt.W[l] = w; nt.b[l] = b;
    if (l == 0) {
      float* wt = nullptr;
      const int td = d->t_dim > 0 ? d->t_dim : 1;
      if (cudaMalloc(&wt, sizeof(float) * td * Np) != cudaSuccess) { cleanup(); return fail(FFB_ERR_CUDA, "cudaMalloc"); }
      net->allocs.push_back(wt);
      if (d->t_dim > 0) {
        k_pack_time_tc<<<(d->t_dim * Np + 255) / 256, 256, 0, stream>>>(d->weight[0], in_f, N, wt, d->t_dim, Np, d->t_col);
        g_launches += 1;
      }
      nt.Wt = wt;
    }
    in_f = N;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { cleanup(); return fail(FFB_ERR_CUDA, std::string("ffb_net_create: ") + cudaGetErrorString(e)); }
  *out = net;
  return FFB_OK;
}

extern "C" void ffb_net_destroy(ffb_net* net) {
  if (!net) return;
  for (void* p : net->allocs) cudaFree(p);
  delete net;
}

extern "C" int64_t ffb_net_flops(const ffb_net* net) { return net ? net->flops : 0; }

// ---- field validation / conversion -----------------------------------------------------------
static int tangents_of(const ffb_field* f) {
  return f->div_mode == FFB_DIV_EXACT ? f->net[0]->dev.x_dim : (f->div_mode == FFB_DIV_HUTCH ? 1 : 0);
}

static size_t field_smem(const FieldDev& fd, int T, int slots) {
  return engine() ? smem_layout_tc(fd.state_dim, fd.cond_dim, T, fd.div_mode == FFB_DIV_HUTCH, slots, fd.n_calls, field_tdim(fd), nullptr)
                  : smem_layout(fd.state_dim, fd.cond_dim, T, fd.div_mode == FFB_DIV_HUTCH, slots, nullptr);
}

static int make_field(const ffb_field* f, FieldDev* out) {
  if (!f) return fail(FFB_ERR_ARG, "null field");
  if (f->n_calls < 1 || f->n_calls > 2) return fail(FFB_ERR_ARG, "field: n_calls must be 1 or 2");
  if (f->state_dim < 1 || f->state_dim > FFB_MAX_STATE) return fail(FFB_ERR_ARG, "field: state_dim must be in 1..128");
  if (f->cond_dim < 0 || f->cond_dim > FFB_MAX_WIDTH) return fail(FFB_ERR_ARG, "field: bad cond_dim");
  memset(out, 0, sizeof(*out));
  out->n_calls = f->n_calls;
  for (int c = 0; c < f->n_calls; ++c) {
    if (!f->net[c]) return fail(FFB_ERR_ARG, "field: null net");
    const NetDev& nd = f->net[c]->dev;
    const int dout = nd.N[nd.n_layers - 1];
    if (nd.c_dim != f->cond_dim) return fail(FFB_ERR_ARG, "field: net conditional width differs from field cond_dim");
    if (f->in_off[c] < 0 || f->in_off[c] + nd.x_dim > f->state_dim) return fail(FFB_ERR_ARG, "field: input block outside the state");
    if (f->out_off[c] < 0 || f->out_off[c] + dout > f->state_dim) return fail(FFB_ERR_ARG, "field: output block outside the state");
    out->net[c] = engine() ? f->net[c]->tc : nd;
    out->in_off[c] = f->in_off[c];
    out->out_off[c] = f->out_off[c];
    out->out_sign[c] = f->out_sign[c];
  }
  if (f->div_mode != FFB_DIV_NONE) {
    const NetDev& nd = f->net[0]->dev;
    if (f->n_calls != 1 || nd.x_dim != f->state_dim || nd.N[nd.n_layers - 1] != f->state_dim)
      return fail(FFB_ERR_ARG, "field: divergence needs a single network mapping the state to itself");
    if (f->div_mode == FFB_DIV_EXACT && 1 + nd.x_dim > TM)
      return fail(FFB_ERR_ARG, "field: exact trace supports at most 127 state columns");
  }
  out->state_dim = f->state_dim; out->cond_dim = f->cond_dim; out->kind = f->kind;
  out->use_sigma = f->use_sigma; out->has_drift = f->has_drift; out->div_mode = f->div_mode;
  // keep the Y0 / K1..K7 slots in shared memory whenever the tile still fits in one SM
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t with_slots = field_smem(*out, tangents_of(f), 1);
  out->slots_smem = (optin > 0 && with_slots <= (size_t)optin) ? 1 : 0;
  return FFB_OK;
}

static int smem_optin();
// Tangent-row engine plan: samples per tile (<= what the 128 rows hold, reduced until the per-sample state fits
// shared memory) and the dynamic shared-memory size; 0 samples = the field cannot run on it.
static int rrt_plan(const ffb_field* f, int state_dim, int cond_dim, int tdim, int nslot, int nbeff, size_t* smem_out) {
  if (f->div_mode == FFB_DIV_NONE) return 0;
  int S = rrt_rowmap(tangents_of(f), -1, nullptr);
  for (; S >= 1; --S) {
    const size_t smem = smem_layout_rrt(state_dim, cond_dim, rrt_ld(S), f->div_mode == FFB_DIV_HUTCH, nslot, tdim, nbeff, nullptr);
    if (smem <= (size_t)smem_optin()) { if (smem_out) *smem_out = smem; return S; }
  }
  return 0;
}
static int field_tdim_host(const ffb_field* f) {
  int t = f->net[0]->tc.t_dim;
  if (f->n_calls > 1 && f->net[1] && f->net[1]->tc.t_dim > t) t = f->net[1]->tc.t_dim;
  return t;
}
// Tiles of `batch` rows: the caller sizes `partials` with this, so it is the maximum over the tile shapes
// the field's kernels use (dense rows for the tile engine, quarter-packed samples for the tangent engine).
extern "C" int64_t ffb_num_tiles(const ffb_field* f, int64_t batch) {
  if (!f || !f->net[0]) return -1;
  const int S = TM / (1 + tangents_of(f));
  int64_t n = (batch + S - 1) / S;
  if (f->div_mode != FFB_DIV_NONE) {
    const int td = field_tdim_host(f);
    const int configs[2][2] = {{NSLOT, 6}, {3, 1}};          // dopri5 attempt, single evaluation
    for (auto& c : configs) {
      const int St = rrt_plan(f, f->state_dim, f->cond_dim, td, c[0], c[1], nullptr);
      if (St > 0) n = std::max<int64_t>(n, (batch + St - 1) / St);
    }
  }
  return n;
}

static int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n > 0 ? n : 148;
}
int ffb_num_sms() { return num_sms(); }

extern "C" size_t ffb_scratch_bytes(const ffb_field* f) {
  if (!f) return 0;
  return std::max((size_t)num_sms() * NSLOT * f->state_dim * LDA * sizeof(float), rd_scratch_bytes(f->state_dim, f->cond_dim));
}

template <typename Kern, typename Args>
static int launch_tiles(Kern kern, int nthr, const char* name, const ffb_field* f, const FieldDev& fd, const Args& a,
                        int64_t batch, cudaStream_t stream) {
  const int T = tangents_of(f);
  for (int c = 0; c < fd.n_calls; ++c)
    if (fd.net[c].act != FFB_ACT_SILU)
      return fail(FFB_ERR_ARG, std::string(name) + ": non-SiLU activations run on the chunk-pipelined tensor-core engines only "
                  "(not on FFB_ENGINE=ffma / tc_tile, fixed grids with a divergence, or samples wider than a tile)");
  const size_t smem = field_smem(fd, T, fd.slots_smem);
  int dev = 0, optin = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ((int)smem > optin)
    return fail(FFB_ERR_ARG, std::string(name) + ": tile needs " + std::to_string(smem) + " B of shared memory, device allows " + std::to_string(optin));
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int Sd = TM / (1 + T);
  const int64_t ntiles = (batch + Sd - 1) / Sd;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  kern<<<grid, nthr, smem, stream>>>(fd, a, ntiles);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

// ---- row-resident tensor-core kernels (fields without tangent rows) -----------------------------
// FFB_ENGINE=tc_tile (or ffb_set_engine(2)) keeps the older whole-layer hand-off tile engine for A/B runs
static bool use_rr(const FieldDev& fd) { return chunk_engines() && fd.div_mode == FFB_DIV_NONE; }
// The dual-tile engine (two resident tiles per SM) takes the dopri5 attempts of fields without tangent rows when there
// are at least two tiles per SM (below that the single-tile engine spreads the tiles over more SMs) and the stage input
// fits shared memory beside its operand images.  Both engines issue the same MMAs in the same order: same bits.
static bool use_rd(const FieldDev& fd, int64_t batch) {
  if (fd.div_mode != FFB_DIV_NONE || !(engine() == 1 || engine() == 4) || !rd_dopri5_fits(fd)) return false;
  return engine() == 4 || (batch + TM - 1) / TM >= 2 * (int64_t)num_sms();
}
// true when a network of the field has a non-SiLU activation (selects the kernels with the run-time dispatch)
static bool gen_act(const FieldDev& fd) {
  for (int c = 0; c < fd.n_calls; ++c) if (fd.net[c].act != FFB_ACT_SILU) return true;
  return false;
}

static int smem_optin() {
  static int v = 0;
  if (!v) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); }
  return v;
}
// decide where the state slots live (shared memory when the tile still fits) and return the block size
static size_t rr_pick_smem(FieldDev* fd, int nslot, int nbeff) {
  const int td = field_tdim(*fd);
  const size_t with_slots = smem_layout_rr(fd->state_dim, fd->cond_dim, nslot, fd->n_calls, td, nbeff, nullptr);
  fd->slots_smem = (with_slots <= (size_t)smem_optin()) ? 1 : 0;
  return fd->slots_smem ? with_slots : smem_layout_rr(fd->state_dim, fd->cond_dim, 0, fd->n_calls, td, nbeff, nullptr);
}
template <typename Kern, typename Args>
static int launch_rr(Kern kern, size_t smem, const char* name, const FieldDev& fd, const Args& a, int64_t batch,
                     cudaStream_t stream) {
  if ((int)smem > smem_optin())
    return fail(FFB_ERR_ARG, std::string(name) + ": tile needs " + std::to_string(smem) + " B of shared memory, device allows " + std::to_string(smem_optin()));
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (batch + TM - 1) / TM;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  kern<<<grid, RR_NTHR, smem, stream>>>(fd, a, ntiles);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

// ---- tangent-row engine (log-likelihood paths) ------------------------------------------------------------
// returns 0 when the field cannot run on it; otherwise sets fd->rrt_cap and returns the shared-memory size
static size_t rrt_smem(const ffb_field* f, FieldDev* fd, int nslot, int nbeff) {
  if (!chunk_engines() || fd->div_mode == FFB_DIV_NONE) return 0;
  size_t smem = 0;
  const int S = rrt_plan(f, fd->state_dim, fd->cond_dim, field_tdim(*fd), nslot, nbeff, &smem);
  if (S <= 0) return 0;
  fd->rrt_cap = S;
  return smem;
}
template <typename Kern, typename Args>
static int launch_rrt(Kern kern, size_t smem, const char* name, const FieldDev& fd, const Args& a, int64_t batch,
                      cudaStream_t stream) {
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int S = fd.rrt_cap;
  const int64_t ntiles = (batch + S - 1) / S;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  kern<<<grid, RR_NTHR, smem, stream>>>(fd, a, ntiles);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  (void)name;
  return FFB_OK;
}

extern "C" int ffb_field_eval(const ffb_field* f, const ffb_eval_args* a, void* stream) {
  FieldDev fd;
  int rc = make_field(f, &fd);
  if (rc) return rc;
  if (!a || !a->y || !a->scratch) return fail(FFB_ERR_ARG, "ffb_field_eval: y and scratch are required");
  if (fd.cond_dim && !a->cond) return fail(FFB_ERR_ARG, "ffb_field_eval: cond is required");
  if (fd.div_mode == FFB_DIV_HUTCH && !a->probes) return fail(FFB_ERR_ARG, "ffb_field_eval: probes are required");
  if (a->norms && !a->partials) return fail(FFB_ERR_ARG, "ffb_field_eval: partials buffer required for norms");
  if (a->norms == 2 && (!a->fbase || (fd.div_mode != FFB_DIV_NONE && !a->dlpbase)))
    return fail(FFB_ERR_ARG, "ffb_field_eval: norms=2 needs fbase (and dlpbase with a divergence)");
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  if (a->jac && fd.div_mode != FFB_DIV_EXACT) return fail(FFB_ERR_ARG, "ffb_field_eval: jac needs div_mode FFB_DIV_EXACT");
  if (const size_t smt = rrt_smem(f, &fd, 3, 1))
    return gen_act(fd) ? launch_rrt(k_field_eval_rrt<true>, smt, "ffb_field_eval", fd, *a, a->batch, st_)
                       : launch_rrt(k_field_eval_rrt<false>, smt, "ffb_field_eval", fd, *a, a->batch, st_);
  if (a->jac) return fail(FFB_ERR_ARG, "ffb_field_eval: jac is written by the tangent-row tensor-core engine only");
  if (use_rr(fd)) {
    const size_t smem = rr_pick_smem(&fd, 3, 1);
    if (gen_act(fd)) {
      if (fd.slots_smem) return launch_rr(k_field_eval_rr<true, true>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
      return launch_rr(k_field_eval_rr<false, true>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
    }
    if (fd.slots_smem) return launch_rr(k_field_eval_rr<true, false>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
    return launch_rr(k_field_eval_rr<false, false>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
  }
  if (engine()) {
    if (fd.slots_smem) return launch_tiles(k_field_eval<EngineTC, true>, EngineTC::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
    return launch_tiles(k_field_eval<EngineTC, false>, EngineTC::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
  }
  if (fd.slots_smem) return launch_tiles(k_field_eval<EngineFFMA, true>, EngineFFMA::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
  return launch_tiles(k_field_eval<EngineFFMA, false>, EngineFFMA::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
}

extern "C" int ffb_dopri5_attempt(const ffb_field* f, const ffb_dopri5_args* a, void* stream) {
  FieldDev fd;
  int rc = make_field(f, &fd);
  if (rc) return rc;
  if (!a || !a->y0 || !a->f0 || !a->y1 || !a->f1 || !a->partials || !a->scratch)
    return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: y0, f0, y1, f1, partials, scratch are required");
  if (fd.div_mode != FFB_DIV_NONE && (!a->lp0 || !a->dlp0 || !a->lp1 || !a->dlp1))
    return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: log-det buffers are required with a divergence");
  if ((a->final || a->ctl) && (!a->y_out || (fd.div_mode != FFB_DIV_NONE && !a->lp_out)))
    return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: final step needs y_out (and lp_out)");
  if (fd.cond_dim && !a->cond) return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: cond is required");
  if (fd.div_mode == FFB_DIV_HUTCH && !a->probes) return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: probes are required");
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  const bool dyn = a->ctl != nullptr;
  if (const size_t smt = rrt_smem(f, &fd, NSLOT, 6)) {
    if (dyn) return gen_act(fd) ? launch_rrt(k_dopri5_rrt<true, true>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_)
                                : launch_rrt(k_dopri5_rrt<false, true>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_);
    return gen_act(fd) ? launch_rrt(k_dopri5_rrt<true, false>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_)
                       : launch_rrt(k_dopri5_rrt<false, false>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_);
  }
  if (use_rd(fd, a->batch)) return rd_launch_dopri5(fd, *a, st_);
  if (use_rr(fd)) {
    const size_t smem = rr_pick_smem(&fd, NSLOT, 6);
#define FFB_RR_DOPRI5(SS_, GEN_)                                                                                      \
  (dyn ? launch_rr(k_dopri5_rr<SS_, GEN_, true>, smem, "ffb_dopri5_attempt", fd, *a, a->batch, st_)                    \
       : launch_rr(k_dopri5_rr<SS_, GEN_, false>, smem, "ffb_dopri5_attempt", fd, *a, a->batch, st_))
    if (gen_act(fd)) return fd.slots_smem ? FFB_RR_DOPRI5(true, true) : FFB_RR_DOPRI5(false, true);
    return fd.slots_smem ? FFB_RR_DOPRI5(true, false) : FFB_RR_DOPRI5(false, false);
#undef FFB_RR_DOPRI5
  }
  if (dyn) return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: args->ctl needs the chunk-pipelined tensor-core engines (ffb_dopri5_ctl_supported)");
  if (engine()) {
    if (fd.slots_smem) return launch_tiles(k_dopri5<EngineTC, true>, EngineTC::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
    return launch_tiles(k_dopri5<EngineTC, false>, EngineTC::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
  }
  if (fd.slots_smem) return launch_tiles(k_dopri5<EngineFFMA, true>, EngineFFMA::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
  return launch_tiles(k_dopri5<EngineFFMA, false>, EngineFFMA::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
}

extern "C" int ffb_integrate_fixed(const ffb_field* f, const ffb_fixed_args* a, void* stream) {
  FieldDev fd;
  int rc = make_field(f, &fd);
  if (rc) return rc;
  if (!a || !a->x0 || !a->x_out || !a->step_table || !a->ev_table || !a->scratch || !a->status)
    return fail(FFB_ERR_ARG, "ffb_integrate_fixed: x0, x_out, step_table, ev_table, scratch, status are required");
  if (a->method < FFB_M_EULER || a->method > FFB_M_LEAPFROG) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: unknown method");
  if (a->method == FFB_M_LEAPFROG && fd.n_calls != 2) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: leapfrog needs a 2-network (q,p) field");
  if ((a->method == FFB_M_EM || a->method == FFB_M_LEAPFROG) && fd.div_mode != FFB_DIV_NONE)
    return fail(FFB_ERR_ARG, "ffb_integrate_fixed: no divergence with this method");
  if (fd.cond_dim && !a->cond) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: cond is required");
  if (fd.div_mode == FFB_DIV_HUTCH && !a->probes) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: probes are required");
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  if (const size_t smt = rrt_smem(f, &fd, rr_fixed_slots(a->method), 1))
    return gen_act(fd) ? launch_rrt(k_fixed_rrt<true>, smt, "ffb_integrate_fixed", fd, *a, a->batch, st_)
                       : launch_rrt(k_fixed_rrt<false>, smt, "ffb_integrate_fixed", fd, *a, a->batch, st_);
  if (use_rr(fd)) {
    const size_t smem = rr_pick_smem(&fd, rr_fixed_slots(a->method), 8);
    if (gen_act(fd)) {
      if (fd.slots_smem) return launch_rr(k_fixed_rr<true, true>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
      return launch_rr(k_fixed_rr<false, true>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
    }
    if (fd.slots_smem) return launch_rr(k_fixed_rr<true, false>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
    return launch_rr(k_fixed_rr<false, false>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
  }
  if (engine()) {
    if (fd.slots_smem) return launch_tiles(k_fixed<EngineTC, true>, EngineTC::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
    return launch_tiles(k_fixed<EngineTC, false>, EngineTC::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
  }
  if (fd.slots_smem) return launch_tiles(k_fixed<EngineFFMA, true>, EngineFFMA::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
  return launch_tiles(k_fixed<EngineFFMA, false>, EngineFFMA::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
}

extern "C" int ffb_dopri5_ctl_supported(const ffb_field* f) {
  FieldDev fd;
  if (make_field(f, &fd)) return 0;
  if (rrt_smem(f, &fd, NSLOT, 6)) return 1;
  return use_rr(fd) ? 1 : 0;
}

extern "C" int ffb_dopri5_control(const ffb_dopri5_ctl_params* p, double* sums, const double* partials, int64_t n_tiles,
                                  ffb_dopri5_ctl* ctl, int32_t after_attempt, void* stream) {
  if (!p || !ctl || (after_attempt && !sums)) return fail(FFB_ERR_ARG, "ffb_dopri5_control: null argument");
  if (p->n_grid < 0 || p->n_grid > FFB_CTL_MAX_GRID) return fail(FFB_ERR_ARG, "ffb_dopri5_control: at most 16 step_t points");
  if (p->prog.n_freq < 0 || p->prog.n_freq > FFB_MAX_FREQ) return fail(FFB_ERR_ARG, "ffb_dopri5_control: bad n_freq");
  if (after_attempt && partials)
    k_dopri5_reduce_control<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, partials, n_tiles, sums, ctl);
  else
    k_dopri5_control<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, sums, ctl, after_attempt);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_dopri5_control_host(const ffb_dopri5_ctl_params* p, const double* sums, ffb_dopri5_ctl* ctl,
                                       int32_t after_attempt) {
  if (!p || !ctl || (after_attempt && !sums)) return fail(FFB_ERR_ARG, "ffb_dopri5_control_host: null argument");
  if (p->n_grid < 0 || p->n_grid > FFB_CTL_MAX_GRID) return fail(FFB_ERR_ARG, "ffb_dopri5_control_host: at most 16 step_t points");
  if (p->prog.n_freq < 0 || p->prog.n_freq > FFB_MAX_FREQ) return fail(FFB_ERR_ARG, "ffb_dopri5_control_host: bad n_freq");
  ffbctl::control_turn(*p, sums, *ctl, after_attempt);
  return FFB_OK;
}

extern "C" int ffb_time_program_rows(const ffb_time_program* prog, const float* times, int32_t n, float sign,
                                     ffb_eval_scalars* out, int32_t on_device) {
  if (!prog || !times || !out || n < 0) return fail(FFB_ERR_ARG, "ffb_time_program_rows: bad argument");
  if (prog->n_freq < 0 || prog->n_freq > FFB_MAX_FREQ) return fail(FFB_ERR_ARG, "ffb_time_program_rows: bad n_freq");
  if (n == 0) return FFB_OK;
  if (!on_device) {
    for (int i = 0; i < n; ++i) ffbctl::program_row(*prog, times[i], sign, out + i);
    return FFB_OK;
  }
  float* dt = nullptr; ffb_eval_scalars* dout = nullptr;
  CUDA_TRY(cudaMalloc(&dt, sizeof(float) * n));
  if (cudaMalloc(&dout, sizeof(ffb_eval_scalars) * n) != cudaSuccess) { cudaFree(dt); return fail(FFB_ERR_CUDA, "cudaMalloc"); }
  cudaMemcpy(dt, times, sizeof(float) * n, cudaMemcpyHostToDevice);
  k_time_program<<<(n + 127) / 128, 128>>>(*prog, dt, n, sign, dout);
  g_launches += 1;
  cudaError_t e = cudaMemcpy(out, dout, sizeof(ffb_eval_scalars) * n, cudaMemcpyDeviceToHost);
  cudaFree(dt); cudaFree(dout);
  if (e != cudaSuccess) return fail(FFB_ERR_CUDA, std::string("ffb_time_program_rows: ") + cudaGetErrorString(e));
  return FFB_OK;
}

extern "C" int ffb_reduce_partials(const double* partials, int64_t n_tiles, double* sums, void* stream) {
  if (!partials || !sums) return fail(FFB_ERR_ARG, "ffb_reduce_partials: null argument");
  k_reduce<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partials, n_tiles, sums);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_gaussian_logprob(const float* x, const float* add, float* out, int64_t batch, int32_t dim,
                                    float sigma, void* stream) {
  if (!x || !out || dim < 1) return fail(FFB_ERR_ARG, "ffb_gaussian_logprob: bad argument");
  if (batch == 0) return FFB_OK;
  const float log_norm = logf(sigma) + 0.918938533204672742f;   // log sigma + log sqrt(2 pi)
  const int wpb = 8;
  k_gauss_logprob<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, add, out, batch, dim, sigma, log_norm);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_philox_normal(float* out, int64_t batch, int32_t dim, uint64_t seed, uint64_t offset, int32_t step,
                                 int64_t row_offset, void* stream) {
  if (!out || dim < 1) return fail(FFB_ERR_ARG, "ffb_philox_normal: bad argument");
  const int64_t n = batch * ((dim + 3) / 4);
  if (n == 0) return FFB_OK;
  k_philox<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, batch, dim, seed,
                                                                                             offset, step, row_offset);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_ffma_peak(int32_t iters, float* tflops, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  float* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 4));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int grid = num_sms() * 8;
  k_ffma_peak<<<grid, 256, 0, stream>>>(d, iters);   // warm-up
  CUDA_TRY(cudaEventRecord(e0, stream));
  k_ffma_peak<<<grid, 256, 0, stream>>>(d, iters);
  CUDA_TRY(cudaEventRecord(e1, stream));
  g_launches += 2;
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  const double flop = (double)grid * 256.0 * (double)iters * 16.0 * 4.0;   // 16 FFMA2 = 32 FMA = 64 FLOP
  if (tflops) *tflops = (float)(flop / (ms * 1e-3) / 1e12);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return FFB_OK;
}


// ffb_kernels_generic.cuh -- the engine-generic integrator kernels: k_field_eval, k_dopri5, k_fixed are templates over a
// tile engine (EngineFFMA, EngineTC in ffb_kernels.cu; EngineWide in ffb_wide.cu) that provides Ctx / init / fini / eval.
// Included by every translation unit that instantiates them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"

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

// ffb_kernels_rr.cuh -- integrator kernels on the row-resident tensor-core engine (ffb_engine_rr.cuh):
// fields without tangent rows (div_mode == FFB_DIV_NONE).  Included by ffb_kernels.cu.
//
//   k_dopri5_rr   one attempted Dormand-Prince step (6 fused evaluations, FSAL, FP64 error partials,
//                 dense output at t_end) -- same contract as k_dopri5
//   k_fixed_rr    whole fixed-grid trajectory on-chip -- same contract as k_fixed
//
// Persistent: grid = min(#tiles, #SMs), one 576-thread CTA per SM (16 epilogue warps, a weight loader
// warp, an MMA warp).  Between the tile load and the tile store no CTA-wide barrier is executed.
#pragma once
#include "ffb_engine_rr.cuh"

namespace ffb {

template <int NT>
__device__ __forceinline__ void load_rows_t(float* dst, const float* __restrict__ src, int64_t row0, int nv, int S,
                                            int D, int tid) {
  const float* __restrict__ base = src + row0 * D;
#pragma unroll 4
  for (int idx = tid; idx < S * D; idx += NT) {
    const int r = fast_div(idx, D), d = idx - r * D;
    dst[d * LDA + r] = (r < nv) ? base[idx] : 0.0f;
  }
}
template <int NT>
__device__ __forceinline__ void store_rows_t(float* __restrict__ dst, const float* src, int64_t row0, int nv, int D,
                                             int tid) {
  float* __restrict__ base = dst + row0 * D;
#pragma unroll 4
  for (int idx = tid; idx < nv * D; idx += NT) {
    const int r = fast_div(idx, D), d = idx - r * D;
    base[idx] = src[d * LDA + r];
  }
}

template <int NV>
__device__ __forceinline__ void rr_block_reduce_store(CtxR& cx, double (&v)[NV], double* out, const int (&slot)[NV]) {
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
  }
  if (cx.lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) cx.red()[cx.cwarp * FFB_NPART + q] = v[q];
  }
  rr_bar();
  if (cx.tid == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double s = 0.0;
      for (int w = 0; w < RR_NCOMP / 32; ++w) s += cx.red()[w * FFB_NPART + q];
      out[slot[q]] = s;
    }
  }
  rr_bar();
}

// state slots a fixed-grid method needs: K_e in slot e, Y0 (when the method keeps it) in the last slot
__host__ __device__ inline int rr_fixed_slots(int method) {
  return method == FFB_M_RK4 ? 5 : (method == FFB_M_MIDPOINT ? 3 : (method == FFB_M_EULER ? 1 : 2));
}

}  // namespace ffb

// =============================================================================================
// k_field_eval_rr: one evaluation (+ the norms of torchdiffeq's initial-step heuristic)
// =============================================================================================
template <bool SS, bool GEN>
__global__ void __launch_bounds__(ffb::RR_NTHR, 1) k_field_eval_rr(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_eval_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENG = EngineRR_<GEN>;
  CtxR cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch), 3, 1);
  const int SD = cx.SD, CD = cx.CD;
  if (!cx.producer) {
    ENG::prep_beff(cx, f, a.ev.tfeat, cx.beff());
    rr_bar();
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    const int nv = (int)min((int64_t)TM, a.batch - row0);
    float* F = rr_slot<SS>(cx, 0);
    float* Y0 = rr_slot<SS>(cx, 1);
    float* FB = rr_slot<SS>(cx, 2);
    if (!cx.producer) {
      load_rows_t<RR_NCOMP>(Y0, a.y, row0, nv, TM, SD, cx.tid);
      if (a.fbase) load_rows_t<RR_NCOMP>(FB, a.fbase, row0, nv, TM, SD, cx.tid);
      if (CD) load_rows_t<RR_NCOMP>(cx.condb(), a.cond, row0, nv, TM, CD, cx.tid);
      rr_bar();
      rr_for_blocks(cx, [&](int d0) {
        float y0v[8], fb[8];
        rr_load8(cx, Y0, d0, y0v);
        rr_load8_if(cx, a.fbase != nullptr, FB, d0, fb);
#pragma unroll
        for (int u = 0; u < 8; ++u) y0v[u] = a.fbase ? __fadd_rn(y0v[u], __fmul_rn(a.h, fb[u])) : y0v[u];
        rr_store8(cx, cx.ycur(), d0, y0v);
      });
    }
    ENG::template eval<SS>(cx, f, a.ev.a, a.ev.c, a.ev.sigma, a.ev.sign, cx.beff(), 0);
    if (!cx.producer) {
      double v[6] = {0, 0, 0, 0, 0, 0};   // x_y, x_f, x_df, lp_f, lp_df, c_y
      if (a.norms && cx.row < nv) {
        rr_for_blocks(cx, [&](int d0) {
          float y0v[8], fv[8], fb[8];
          rr_load8(cx, Y0, d0, y0v);
          rr_load8(cx, F, d0, fv);
          rr_load8_if(cx, a.norms == 2, FB, d0, fb);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (d0 + u >= SD) continue;
            const float sc = __fadd_rn(a.atol, __fmul_rn(fabsf(y0v[u]), a.rtol));
            if (a.norms == 1) {
              const float q0 = __fdiv_rn(y0v[u], sc), q1 = __fdiv_rn(fv[u], sc);
              v[0] += (double)q0 * q0;
              v[1] += (double)q1 * q1;
            } else {
              const float q2 = __fdiv_rn(__fsub_rn(fv[u], fb[u]), sc);
              v[2] += (double)q2 * q2;
            }
          }
        });
      }
      rr_bar();                                            // slot 0 complete for every row
      if (a.f) store_rows_t<RR_NCOMP>(a.f, F, row0, nv, SD, cx.tid);
      if (a.norms) {
        if (a.cond_in_state && a.norms == 1) {
          const float* cs = a.cond_state ? a.cond_state : a.cond;
          for (int idx = cx.tid; idx < CD * nv; idx += RR_NCOMP) {
            const float c = cs[row0 * CD + idx];
            const float q = __fdiv_rn(c, __fadd_rn(a.atol, __fmul_rn(fabsf(c), a.rtol)));
            v[5] += (double)q * q;
          }
        }
        const int slot[6] = {P_X_Y, P_X_F, P_X_DF, P_LP_F, P_LP_DF, P_C_Y};
        rr_block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot);
      } else {
        rr_bar();
      }
    }
  }
  ENG::fini(cx);
}

// =============================================================================================
// k_dopri5_rr
// =============================================================================================
// DYN: the step (scalars, dt-scaled tableau, final flag) and the roles of the two state buffers come from the
// device-resident controller block a.ctl (ffb_control.cuh) instead of the launch arguments; a finished solve
// makes the kernel return at once.
template <bool SS, bool GEN, bool DYN>
__global__ void __launch_bounds__(ffb::RR_NTHR, 1) k_dopri5_rr(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_dopri5_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENG = EngineRR_<GEN>;
  if (DYN) { if (__ldg(&a.ctl->done) != 0) return; }
  // controller values are re-read (L1 hits) where they are used so that none stays in a register across the evaluations
#define FFB_STEP(x) (DYN ? __ldg(&a.ctl->x) : a.x)
#define FFB_SWAPPED() (DYN && __ldg(&a.ctl->cur) != 0)
  CtxR cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch), NSLOT, 6);
  const int SD = cx.SD, CD = cx.CD;
  const int bstride = f.n_calls * KMAX;
  if (!cx.producer) {
    for (int s = 0; s < 6; ++s) ENG::prep_beff(cx, f, DYN ? a.ctl->ev[s].tfeat : a.ev[s].tfeat, cx.beff() + s * bstride);
    rr_bar();
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    const int nv = (int)min((int64_t)TM, a.batch - row0);
    float* Y0 = rr_slot<SS>(cx, SLOT_Y0);
    double nonfinite = 0.0;
    if (!cx.producer) {
      const bool sw = FFB_SWAPPED();
      load_rows_t<RR_NCOMP>(Y0, sw ? a.y1 : a.y0, row0, nv, TM, SD, cx.tid);
      load_rows_t<RR_NCOMP>(rr_slot<SS>(cx, 0), sw ? a.f1 : a.f0, row0, nv, TM, SD, cx.tid);
      if (CD) load_rows_t<RR_NCOMP>(cx.condb(), a.cond, row0, nv, TM, CD, cx.tid);
      rr_bar();
      const float* K1 = rr_slot<SS>(cx, 0);
      const float c00 = FFB_STEP(cb[0][0]);
      rr_for_blocks(cx, [&](int d0) {
        float y0v[8], kv[8], y[8];
        rr_load8(cx, Y0, d0, y0v);
        rr_load8(cx, K1, d0, kv);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (d0 + u < SD && !is_finite_f(y0v[u])) nonfinite += 1.0;
          y[u] = __fadd_rn(y0v[u], __fmul_rn(kv[u], c00));
        }
        rr_store8(cx, cx.ycur(), d0, y);
      });
    }
    for (int i = 1; i <= 6; ++i) {
      if constexpr (DYN) ENG::template eval_ev<SS>(cx, f, EvCtl{a.ctl, i - 1}, cx.beff() + (i - 1) * bstride, i);
      else ENG::template eval<SS>(cx, f, a.ev[i - 1].a, a.ev[i - 1].c, a.ev[i - 1].sigma, a.ev[i - 1].sign, cx.beff() + (i - 1) * bstride, i);
      if (!cx.producer && i < 6) {
        // input of stage i+1: y0 + sum_j cb[i][j] k_j  (the 7th stage input is y1: FSAL)
        float cbi[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) cbi[j] = FFB_STEP(cb[i][j]);
        rr_for_blocks(cx, [&](int d0) {
          float y0v[8], kv[8], acc[8];
          rr_load8(cx, Y0, d0, y0v);
          rr_load8(cx, rr_slot<SS>(cx, 0), d0, kv);
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = __fmul_rn(kv[u], cbi[0]);
#pragma unroll
          for (int j = 1; j < 6; ++j) {
            rr_load8_if(cx, j <= i, rr_slot<SS>(cx, j), d0, kv);      // k_j = 0 for the stages not yet taken
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fmaf(kv[u], cbi[j], acc[u]);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = __fadd_rn(y0v[u], acc[u]);
          rr_store8(cx, cx.ycur(), d0, acc);
        });
      }
    }
    if (!cx.producer) {
      // cx.ycur() holds y1, slot 6 holds f1
      double v[2] = {0.0, nonfinite};
      const int final_ = FFB_STEP(final);
      float* OUT = rr_slot<SS>(cx, 1);     // K2 of an element is dead once its error / mid sums are formed
      if (cx.row < nv) {
        float ce[7], cm[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) { ce[j] = FFB_STEP(ce[j]); cm[j] = FFB_STEP(cm[j]); }
        const float dt_ = FFB_STEP(dt), xi_ = FFB_STEP(x_interp);
        rr_for_blocks(cx, [&](int d0) {
          float y0v[8], y1v[8], k0[8], kv[8], err[8], mid[8];
          rr_load8(cx, Y0, d0, y0v);
          rr_load8(cx, cx.ycur(), d0, y1v);
          rr_load8(cx, rr_slot<SS>(cx, 0), d0, k0);
#pragma unroll
          for (int u = 0; u < 8; ++u) { err[u] = __fmul_rn(k0[u], ce[0]); mid[u] = __fmul_rn(k0[u], cm[0]); }
#pragma unroll
          for (int j = 1; j < 7; ++j) {
            rr_load8(cx, rr_slot<SS>(cx, j), d0, kv);
#pragma unroll
            for (int u = 0; u < 8; ++u) { err[u] = fmaf(kv[u], ce[j], err[u]); mid[u] = fmaf(kv[u], cm[j], mid[u]); }
          }
          // kv now holds k7 = f1
          float out[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(y0v[u]), fabsf(y1v[u]))));
            const float q = __fdiv_rn(err[u], tol);
            if (d0 + u < SD) v[0] += (double)q * q;
            out[u] = final_ ? dense_output(y0v[u], y1v[u], __fadd_rn(y0v[u], mid[u]), k0[u], kv[u], dt_, xi_) : 0.0f;
          }
          if (final_) rr_store8(cx, OUT, d0, out);
        });
      }
      rr_bar();
      const bool sw = FFB_SWAPPED();
      store_rows_t<RR_NCOMP>(sw ? const_cast<float*>(a.y0) : a.y1, cx.ycur(), row0, nv, SD, cx.tid);
      store_rows_t<RR_NCOMP>(sw ? const_cast<float*>(a.f0) : a.f1, rr_slot<SS>(cx, 6), row0, nv, SD, cx.tid);
      if (final_) store_rows_t<RR_NCOMP>(a.y_out, OUT, row0, nv, SD, cx.tid);
      const int slot[2] = {P_X_ERR, P_NONFINITE};
      rr_block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot);
      if (cx.tid == 0) a.partials[tile * FFB_NPART + P_LP_ERR] = 0.0;
    }
  }
  ENG::fini(cx);
#undef FFB_STEP
#undef FFB_SWAPPED
}

// =============================================================================================
// k_fixed_rr
// =============================================================================================
template <bool SS, bool GEN>
__global__ void __launch_bounds__(ffb::RR_NTHR, 1) k_fixed_rr(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_fixed_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENG = EngineRR_<GEN>;
  CtxR cx;
  const int nslot = rr_fixed_slots(a.method);
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch), nslot, 8);
  const int SD = cx.SD, CD = cx.CD;
  const int nev = evals_per_step(a.method);
  const float third = (float)(1.0 / 3.0);
  const int bstride = f.n_calls * KMAX;
  const int first_e = (a.method == FFB_M_LEAPFROG) ? 1 : 0;   // first evaluation of every step but the first
  // per-quarter double buffer of the layer-0 bias: beff[(parity * 4 + q)][call][n]
  auto beff_buf = [&](uint32_t parity) { return cx.beff() + (size_t)((parity & 1u) * 4 + cx.q) * bstride; };
  auto prep_q = [&](const ffb_eval_scalars* evp, float* buf) {
    for (int i = cx.cg * 32 + cx.lane; i < bstride; i += 128) {
      const int c = i / KMAX, n = i - c * KMAX;
      float b = cx.sbias()[(c * NET_MAXL) * KMAX + n];
      const float* wt = cx.swt() + (c * cx.tdim) * KMAX + n;
      for (int j = 0; j < f.net[c].t_dim; ++j) b = fmaf(wt[j * KMAX], evp->tfeat[j], b);
      buf[i] = b;
    }
  };
  const int q_lo = f.out_off[0], q_hi = f.out_off[0] + f.net[0].N[f.net[0].n_layers - 1];
  const int p_lo = f.out_off[1], p_hi = f.out_off[1] + f.net[1].N[f.net[1].n_layers > 0 ? f.net[1].n_layers - 1 : 0];
  uint32_t nbuf = 0;                                          // evaluations done by this CTA (buffer parity)
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    const int nv = (int)min((int64_t)TM, a.batch - row0);
    float* Y0 = rr_slot<SS>(cx, nslot - 1);
    const float* K1 = rr_slot<SS>(cx, 0);
    const float* K2 = rr_slot<SS>(cx, nslot > 1 ? 1 : 0);
    const float* K3 = rr_slot<SS>(cx, nslot > 2 ? 2 : 0);
    const float* K4 = rr_slot<SS>(cx, nslot > 3 ? 3 : 0);
    int nan_step = 0x7fffffff;            // first Euler-Maruyama step of this thread's row that produced a NaN
    if (!cx.producer) {
      load_rows_t<RR_NCOMP>(cx.ycur(), a.x0, row0, nv, TM, SD, cx.tid);
      if (CD) load_rows_t<RR_NCOMP>(cx.condb(), a.cond, row0, nv, TM, CD, cx.tid);
      prep_q(a.ev_table, beff_buf(nbuf));                     // first evaluation of the trajectory
      rr_bar();
    }
    for (int step = 0; step < a.nsteps; ++step) {
      const float* st = a.step_table + (size_t)step * FFB_STEP_STRIDE;
      const ffb_eval_scalars* ev = a.ev_table + (size_t)step * nev;
      const float dt = st[0], g = st[1], sq = st[2], half = st[3];    // loaded before the evaluations that hide the latency
      for (int e = 0; e < nev; ++e) {
        unsigned mask = 3u;
        int dst = e;
        if (a.method == FFB_M_LEAPFROG) {         // e0: dp/dt(q, t0) [first step only], e1: dq/dt, e2: dp/dt
          if (e == 0 && step > 0) mask = 0u;
          else mask = (e == 1) ? 1u : 2u;
          dst = (e == 1) ? 0 : 1;
        }
        if (mask) {
          // the evaluation after this one (its layer-0 bias is prepared while this one's MMAs run)
          const ffb_eval_scalars* nxt = nullptr;
          if (e + 1 < nev) nxt = ev + e + 1;
          else if (step + 1 < a.nsteps) nxt = ev + nev + first_e;
          float ea = 0.f, ec = 0.f, es = 1.f, esg = 1.f;
          if (!cx.producer) { ea = ev[e].a; ec = ev[e].c; es = ev[e].sigma; esg = ev[e].sign; }
          ENG::template eval<SS>(cx, f, ea, ec, es, esg, beff_buf(nbuf), dst, mask,
                             [&]() { if (nxt) prep_q(nxt, beff_buf(nbuf + 1)); });
          ++nbuf;
        }
        if (cx.producer) continue;
        float* y = cx.ycur();
        switch (a.method) {
          case FFB_M_EULER:
            rr_for_blocks(cx, [&](int d0) {
              float yv[8], k1[8];
              rr_load8(cx, y, d0, yv); rr_load8(cx, K1, d0, k1);
#pragma unroll
              for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, k1[u]));
              rr_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_MIDPOINT:
            rr_for_blocks(cx, [&](int d0) {
              float yv[8], kv[8];
              if (e == 0) {
                rr_load8(cx, y, d0, yv); rr_load8(cx, K1, d0, kv);
                rr_store8(cx, Y0, d0, yv);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(kv[u], half));
              } else {
                rr_load8(cx, Y0, d0, yv); rr_load8(cx, K2, d0, kv);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, kv[u]));
              }
              rr_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_RK4:
            rr_for_blocks(cx, [&](int d0) {
              float y0v[8], k1[8], k2[8], k3[8], k4[8], yv[8];
              rr_load8(cx, K1, d0, k1);
              if (e == 0) {
                rr_load8(cx, y, d0, y0v);
                rr_store8(cx, Y0, d0, y0v);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(__fmul_rn(dt, k1[u]), third));
              } else if (e == 1) {
                rr_load8(cx, Y0, d0, y0v); rr_load8(cx, K2, d0, k2);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(dt, __fsub_rn(k2[u], __fmul_rn(k1[u], third))));
              } else if (e == 2) {
                rr_load8(cx, Y0, d0, y0v); rr_load8(cx, K2, d0, k2); rr_load8(cx, K3, d0, k3);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[u], k2[u]), k3[u])));
              } else {
                rr_load8(cx, Y0, d0, y0v); rr_load8(cx, K2, d0, k2); rr_load8(cx, K3, d0, k3); rr_load8(cx, K4, d0, k4);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const float sum = __fadd_rn(__fadd_rn(k1[u], __fmul_rn(3.0f, __fadd_rn(k2[u], k3[u]))), k4[u]);
                  yv[u] = __fadd_rn(y0v[u], __fmul_rn(__fmul_rn(sum, dt), 0.125f));
                }
              }
              rr_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_LEAPFROG:
            rr_for_blocks(cx, [&](int d0) {
              float yv[8], kv[8];
              rr_load8(cx, y, d0, yv); rr_load8(cx, (e == 1) ? K1 : K2, d0, kv);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int d = d0 + u;
                if (e == 1) { if (d >= q_lo && d < q_hi) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, kv[u])); }
                else if (d >= p_lo && d < p_hi) yv[u] = __fadd_rn(yv[u], __fmul_rn(half, kv[u]));
              }
              rr_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_EM: {
            // diffusion.py:552-559: f = drift - g^2 score (ev.c = g^2); x_mean = x + f dt; x = x_mean + g dw
            if (cx.row < nv) {
              const float* nz = a.noise ? a.noise + ((size_t)step * a.batch + row0 + cx.row) * SD : nullptr;
              rr_for_blocks(cx, [&](int d0) {
                float yv[8], k1[8], zz[8], xm[8];
                rr_load8(cx, y, d0, yv); rr_load8(cx, K1, d0, k1);
                if (nz) {
#pragma unroll
                  for (int u = 0; u < 8; ++u) zz[u] = nz[min(d0 + u, SD - 1)];
                } else {
                  const float4 za = philox_normal4(a.philox_seed, a.philox_offset, a.row_offset + row0 + cx.row, step, d0 >> 2);
                  zz[0] = za.x; zz[1] = za.y; zz[2] = za.z; zz[3] = za.w;
                  if (d0 + 4 < SD) {
                    const float4 zb = philox_normal4(a.philox_seed, a.philox_offset, a.row_offset + row0 + cx.row, step, (d0 >> 2) + 1);
                    zz[4] = zb.x; zz[5] = zb.y; zz[6] = zb.z; zz[7] = zb.w;
                  } else {
                    zz[4] = zz[5] = zz[6] = zz[7] = 0.0f;
                  }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  xm[u] = __fadd_rn(yv[u], __fmul_rn(k1[u], dt));
                  yv[u] = __fadd_rn(xm[u], __fmul_rn(g, __fmul_rn(zz[u], sq)));
                  if (d0 + u < SD && yv[u] != yv[u]) nan_step = min(nan_step, step);
                }
                rr_store8(cx, Y0, d0, xm);
                rr_store8(cx, y, d0, yv);
              });
            }
            break;
          }
          default: break;
        }
      }
    }
    if (!cx.producer) {
      rr_bar();
      store_rows_t<RR_NCOMP>(a.x_out, (a.method == FFB_M_EM) ? Y0 : cx.ycur(), row0, nv, SD, cx.tid);
      if (nan_step != 0x7fffffff) { atomicOr(a.status, FFB_ST_NAN_SAMPLE); atomicMin(a.status + 1, nan_step); }
      rr_bar();
    }
  }
  ENG::fini(cx);
}

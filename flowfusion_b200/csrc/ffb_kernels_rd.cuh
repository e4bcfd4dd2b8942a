// ffb_kernels_rd.cuh -- integrator kernels on the dual-tile tensor-core engine (ffb_engine_rd.cuh):
// fields without tangent rows (div_mode == FFB_DIV_NONE).  Compiled by ffb_rd.cu.
//
//   k_field_eval_rd  one evaluation (+ the norms of torchdiffeq's initial-step heuristic)
//   k_dopri5_rd      one attempted Dormand-Prince step (6 fused evaluations, FSAL, FP64 error partials,
//                    dense output at t_end)
//   k_fixed_rd       whole fixed-grid trajectory on-chip (euler, midpoint, rk4 3/8, Euler-Maruyama, leapfrog)
//
// Same contracts and the same FP32 statements as the single-tile kernels (ffb_kernels_rr.cuh); results are
// bit-identical to them.  Persistent: grid = min(ceil(#tiles / 2), #SMs), one 576-thread CTA per SM; in round r
// group g of CTA b integrates tile 2 (b + r gridDim) + g.  A group runs its tile from load to store without any
// CTA-wide barrier; the loader and MMA warps serve the groups that hold a tile in the round (`cx.active`).
#pragma once
#include "ffb_engine_rd.cuh"
#include "ffb_kernels_rr.cuh"     // load_rows_t / store_rows_t, rr_fixed_slots

namespace ffb {

template <int NV>
__device__ __forceinline__ void rd_block_reduce_store(CtxD& cx, double (&v)[NV], double* out, const int (&slot)[NV]) {
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
  }
  const int gw = cx.tid >> 5;
  if (cx.lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) cx.red()[gw * FFB_NPART + q] = v[q];
  }
  rd_gbar(cx);
  if (cx.tid == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double s = 0.0;
      for (int w = 0; w < RD_GWARPS; ++w) s += cx.red()[w * FFB_NPART + q];
      out[slot[q]] = s;
    }
  }
  rd_gbar(cx);
}

// tiles of the round that starts at tile `base`: which groups hold one, and this thread's tile
#define FFB_RD_ROUND(base)                                                        \
  cx.active = 1u | (((base) + 1 < ntiles) ? 2u : 0u);                             \
  const int64_t tile = (base) + cx.g;                                             \
  if (!cx.producer && tile >= ntiles) continue;

}  // namespace ffb

// =============================================================================================
// k_field_eval_rd
// =============================================================================================
template <int MEM, bool GEN>
__global__ void __launch_bounds__(ffb::RD_NTHR, 1) k_field_eval_rd(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_eval_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENG = EngineRD_<GEN, MEM>;
  CtxD cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch), 3, 1);
  const int SD = cx.SD, CD = cx.CD;
  if (!cx.producer) {
    ENG::prep_beff(cx, f, a.ev.tfeat, cx.beff(), threadIdx.x, RD_NCOMP);
    rd_allbar();
  }
  for (int64_t base = 2 * (int64_t)blockIdx.x; base < ntiles; base += 2 * (int64_t)gridDim.x) {
    FFB_RD_ROUND(base)
    const int64_t row0 = tile * TM;
    const int nv = (int)min((int64_t)TM, a.batch - row0);
    float* F = rd_slot<MEM>(cx, 0);
    float* Y0 = rd_slot<MEM>(cx, 1);
    float* FB = rd_slot<MEM>(cx, 2);
    float* YC = rd_ycur<MEM>(cx);
    if (!cx.producer) {
      load_rows_t<RD_GTHR>(Y0, a.y, row0, nv, TM, SD, cx.tid);
      if (a.fbase) load_rows_t<RD_GTHR>(FB, a.fbase, row0, nv, TM, SD, cx.tid);
      if (CD) load_rows_t<RD_GTHR>(rd_cond<MEM>(cx), a.cond, row0, nv, TM, CD, cx.tid);
      rd_gbar(cx);
      rd_for_blocks(cx, [&](int d0) {
        float y0v[8], fb[8];
        rd_load8(cx, Y0, d0, y0v);
        rd_load8_if(cx, a.fbase != nullptr, FB, d0, fb);
#pragma unroll
        for (int u = 0; u < 8; ++u) y0v[u] = a.fbase ? __fadd_rn(y0v[u], __fmul_rn(a.h, fb[u])) : y0v[u];
        rd_store8(cx, YC, d0, y0v);
      });
    }
    ENG::eval(cx, f, a.ev.a, a.ev.c, a.ev.sigma, a.ev.sign, cx.beff(), 0);
    if (!cx.producer) {
      double v[6] = {0, 0, 0, 0, 0, 0};   // x_y, x_f, x_df, lp_f, lp_df, c_y
      if (a.norms && cx.row < nv) {
        rd_for_blocks(cx, [&](int d0) {
          float y0v[8], fv[8], fb[8];
          rd_load8(cx, Y0, d0, y0v);
          rd_load8(cx, F, d0, fv);
          rd_load8_if(cx, a.norms == 2, FB, d0, fb);
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
      rd_gbar(cx);                                            // slot 0 complete for every row
      if (a.f) store_rows_t<RD_GTHR>(a.f, F, row0, nv, SD, cx.tid);
      if (a.norms) {
        if (a.cond_in_state && a.norms == 1) {
          const float* cs = a.cond_state ? a.cond_state : a.cond;
          for (int idx = cx.tid; idx < CD * nv; idx += RD_GTHR) {
            const float c = cs[row0 * CD + idx];
            const float q = __fdiv_rn(c, __fadd_rn(a.atol, __fmul_rn(fabsf(c), a.rtol)));
            v[5] += (double)q * q;
          }
        }
        const int slot[6] = {P_X_Y, P_X_F, P_X_DF, P_LP_F, P_LP_DF, P_C_Y};
        rd_block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot);
      } else {
        rd_gbar(cx);
      }
    }
  }
  ENG::fini(cx);
}

// =============================================================================================
// k_dopri5_rd
// =============================================================================================
// DYN: the step (scalars, dt-scaled tableau, final flag) and the roles of the two state buffers come from the
// device-resident controller block a.ctl (ffb_control.cuh) instead of the launch arguments; a finished solve
// makes the kernel return at once.
template <int MEM, bool GEN, bool DYN>
__global__ void __launch_bounds__(ffb::RD_NTHR, 1) k_dopri5_rd(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_dopri5_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENG = EngineRD_<GEN, MEM>;
  if (DYN) { if (__ldg(&a.ctl->done) != 0) return; }
  // controller values are re-read (L1 hits) where they are used so that none stays in a register across the evaluations
#define FFB_STEP(x) (DYN ? __ldg(&a.ctl->x) : a.x)
#define FFB_SWAPPED() (DYN && __ldg(&a.ctl->cur) != 0)
  CtxD cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch), NSLOT, 6);
  const int SD = cx.SD, CD = cx.CD;
  const int bstride = f.n_calls * KMAX;
  if (!cx.producer) {
    for (int s = 0; s < 6; ++s)
      ENG::prep_beff(cx, f, DYN ? a.ctl->ev[s].tfeat : a.ev[s].tfeat, cx.beff() + s * bstride, threadIdx.x, RD_NCOMP);
    rd_allbar();
  }
  for (int64_t base = 2 * (int64_t)blockIdx.x; base < ntiles; base += 2 * (int64_t)gridDim.x) {
    FFB_RD_ROUND(base)
    const int64_t row0 = tile * TM;
    const int nv = (int)min((int64_t)TM, a.batch - row0);
    float* Y0 = rd_slot<MEM>(cx, SLOT_Y0);
    float* YC = rd_ycur<MEM>(cx);
    double nonfinite = 0.0;
    if (!cx.producer) {
      const bool sw = FFB_SWAPPED();
      load_rows_t<RD_GTHR>(Y0, sw ? a.y1 : a.y0, row0, nv, TM, SD, cx.tid);
      load_rows_t<RD_GTHR>(rd_slot<MEM>(cx, 0), sw ? a.f1 : a.f0, row0, nv, TM, SD, cx.tid);
      if (CD) load_rows_t<RD_GTHR>(rd_cond<MEM>(cx), a.cond, row0, nv, TM, CD, cx.tid);
      rd_gbar(cx);
      const float* K1 = rd_slot<MEM>(cx, 0);
      const float c00 = FFB_STEP(cb[0][0]);
      rd_for_blocks(cx, [&](int d0) {
        float y0v[8], kv[8], y[8];
        rd_load8(cx, Y0, d0, y0v);
        rd_load8(cx, K1, d0, kv);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (d0 + u < SD && !is_finite_f(y0v[u])) nonfinite += 1.0;
          y[u] = __fadd_rn(y0v[u], __fmul_rn(kv[u], c00));
        }
        rd_store8(cx, YC, d0, y);
      });
    }
    for (int i = 1; i <= 6; ++i) {
      if constexpr (DYN) ENG::eval_ev(cx, f, EvCtl{a.ctl, i - 1}, cx.beff() + (i - 1) * bstride, i);
      else ENG::eval(cx, f, a.ev[i - 1].a, a.ev[i - 1].c, a.ev[i - 1].sigma, a.ev[i - 1].sign, cx.beff() + (i - 1) * bstride, i);
      if (!cx.producer && i < 6) {
        // input of stage i+1: y0 + sum_j cb[i][j] k_j  (the 7th stage input is y1: FSAL)
        float cbi[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) cbi[j] = FFB_STEP(cb[i][j]);
        rd_for_blocks(cx, [&](int d0) {
          float y0v[8], kv[8], acc[8];
          rd_load8(cx, Y0, d0, y0v);
          rd_load8(cx, rd_slot<MEM>(cx, 0), d0, kv);
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = __fmul_rn(kv[u], cbi[0]);
#pragma unroll
          for (int j = 1; j < 6; ++j) {
            rd_load8_if(cx, j <= i, rd_slot<MEM>(cx, j), d0, kv);      // k_j = 0 for the stages not yet taken
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fmaf(kv[u], cbi[j], acc[u]);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = __fadd_rn(y0v[u], acc[u]);
          rd_store8(cx, YC, d0, acc);
        });
      }
    }
    if (!cx.producer) {
      // the stage input holds y1, slot 6 holds f1
      double v[2] = {0.0, nonfinite};
      const int final_ = FFB_STEP(final);
      float* OUT = rd_slot<MEM>(cx, 1);     // K2 of an element is dead once its error / mid sums are formed
      if (cx.row < nv) {
        float ce[7], cm[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) { ce[j] = FFB_STEP(ce[j]); cm[j] = FFB_STEP(cm[j]); }
        const float dt_ = FFB_STEP(dt), xi_ = FFB_STEP(x_interp);
        rd_for_blocks(cx, [&](int d0) {
          float y0v[8], y1v[8], k0[8], kv[8], err[8], mid[8];
          rd_load8(cx, Y0, d0, y0v);
          rd_load8(cx, YC, d0, y1v);
          rd_load8(cx, rd_slot<MEM>(cx, 0), d0, k0);
#pragma unroll
          for (int u = 0; u < 8; ++u) { err[u] = __fmul_rn(k0[u], ce[0]); mid[u] = __fmul_rn(k0[u], cm[0]); }
#pragma unroll
          for (int j = 1; j < 7; ++j) {
            rd_load8(cx, rd_slot<MEM>(cx, j), d0, kv);
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
          if (final_) rd_store8(cx, OUT, d0, out);
        });
      }
      rd_gbar(cx);
      const bool sw = FFB_SWAPPED();
      store_rows_t<RD_GTHR>(sw ? const_cast<float*>(a.y0) : a.y1, YC, row0, nv, SD, cx.tid);
      store_rows_t<RD_GTHR>(sw ? const_cast<float*>(a.f0) : a.f1, rd_slot<MEM>(cx, 6), row0, nv, SD, cx.tid);
      if (final_) store_rows_t<RD_GTHR>(a.y_out, OUT, row0, nv, SD, cx.tid);
      const int slot[2] = {P_X_ERR, P_NONFINITE};
      rd_block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot);
      if (cx.tid == 0) a.partials[tile * FFB_NPART + P_LP_ERR] = 0.0;
    }
  }
  ENG::fini(cx);
#undef FFB_STEP
#undef FFB_SWAPPED
}

// =============================================================================================
// k_fixed_rd
// =============================================================================================
template <int MEM, bool GEN>
__global__ void __launch_bounds__(ffb::RD_NTHR, 1) k_fixed_rd(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_fixed_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENG = EngineRD_<GEN, MEM>;
  CtxD cx;
  const int nslot = rr_fixed_slots(a.method);
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch), nslot, 16);
  const int SD = cx.SD, CD = cx.CD;
  const int nev = evals_per_step(a.method);
  const float third = (float)(1.0 / 3.0);
  const int bstride = f.n_calls * KMAX;
  const int first_e = (a.method == FFB_M_LEAPFROG) ? 1 : 0;   // first evaluation of every step but the first
  // per-group, per-quarter double buffer of the layer-0 bias: beff[((g * 2 + parity) * 4 + q)][call][n]
  auto beff_buf = [&](uint32_t parity) { return cx.beff() + (size_t)((cx.g * 2 + (parity & 1u)) * 4 + cx.q) * bstride; };
  auto prep_q = [&](const ffb_eval_scalars* evp, float* buf) { ENG::prep_beff(cx, f, evp->tfeat, buf, cx.cg * 32 + cx.lane, 64); };
  const int q_lo = f.out_off[0], q_hi = f.out_off[0] + f.net[0].N[f.net[0].n_layers - 1];
  const int p_lo = f.out_off[1], p_hi = f.out_off[1] + f.net[1].N[f.net[1].n_layers > 0 ? f.net[1].n_layers - 1 : 0];
  uint32_t nbuf = 0;                                          // evaluations done by this group (buffer parity)
  for (int64_t base = 2 * (int64_t)blockIdx.x; base < ntiles; base += 2 * (int64_t)gridDim.x) {
    FFB_RD_ROUND(base)
    const int64_t row0 = tile * TM;
    const int nv = (int)min((int64_t)TM, a.batch - row0);
    float* Y0 = rd_slot<MEM>(cx, nslot - 1);
    float* y = rd_ycur<MEM>(cx);
    const float* K1 = rd_slot<MEM>(cx, 0);
    const float* K2 = rd_slot<MEM>(cx, nslot > 1 ? 1 : 0);
    const float* K3 = rd_slot<MEM>(cx, nslot > 2 ? 2 : 0);
    const float* K4 = rd_slot<MEM>(cx, nslot > 3 ? 3 : 0);
    bool saw_nan = false;
    if (!cx.producer) {
      load_rows_t<RD_GTHR>(y, a.x0, row0, nv, TM, SD, cx.tid);
      if (CD) load_rows_t<RD_GTHR>(rd_cond<MEM>(cx), a.cond, row0, nv, TM, CD, cx.tid);
      prep_q(a.ev_table, beff_buf(nbuf));                     // first evaluation of the trajectory
      rd_gbar(cx);
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
          ENG::eval(cx, f, ea, ec, es, esg, beff_buf(nbuf), dst, mask,
                    [&]() { if (nxt) prep_q(nxt, beff_buf(nbuf + 1)); });
          ++nbuf;
        }
        if (cx.producer) continue;
        switch (a.method) {
          case FFB_M_EULER:
            rd_for_blocks(cx, [&](int d0) {
              float yv[8], k1[8];
              rd_load8(cx, y, d0, yv); rd_load8(cx, K1, d0, k1);
#pragma unroll
              for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, k1[u]));
              rd_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_MIDPOINT:
            rd_for_blocks(cx, [&](int d0) {
              float yv[8], kv[8];
              if (e == 0) {
                rd_load8(cx, y, d0, yv); rd_load8(cx, K1, d0, kv);
                rd_store8(cx, Y0, d0, yv);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(kv[u], half));
              } else {
                rd_load8(cx, Y0, d0, yv); rd_load8(cx, K2, d0, kv);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, kv[u]));
              }
              rd_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_RK4:
            rd_for_blocks(cx, [&](int d0) {
              float y0v[8], k1[8], k2[8], k3[8], k4[8], yv[8];
              rd_load8(cx, K1, d0, k1);
              if (e == 0) {
                rd_load8(cx, y, d0, y0v);
                rd_store8(cx, Y0, d0, y0v);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(__fmul_rn(dt, k1[u]), third));
              } else if (e == 1) {
                rd_load8(cx, Y0, d0, y0v); rd_load8(cx, K2, d0, k2);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(dt, __fsub_rn(k2[u], __fmul_rn(k1[u], third))));
              } else if (e == 2) {
                rd_load8(cx, Y0, d0, y0v); rd_load8(cx, K2, d0, k2); rd_load8(cx, K3, d0, k3);
#pragma unroll
                for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[u], k2[u]), k3[u])));
              } else {
                rd_load8(cx, Y0, d0, y0v); rd_load8(cx, K2, d0, k2); rd_load8(cx, K3, d0, k3); rd_load8(cx, K4, d0, k4);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const float sum = __fadd_rn(__fadd_rn(k1[u], __fmul_rn(3.0f, __fadd_rn(k2[u], k3[u]))), k4[u]);
                  yv[u] = __fadd_rn(y0v[u], __fmul_rn(__fmul_rn(sum, dt), 0.125f));
                }
              }
              rd_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_LEAPFROG:
            rd_for_blocks(cx, [&](int d0) {
              float yv[8], kv[8];
              rd_load8(cx, y, d0, yv); rd_load8(cx, (e == 1) ? K1 : K2, d0, kv);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int d = d0 + u;
                if (e == 1) { if (d >= q_lo && d < q_hi) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, kv[u])); }
                else if (d >= p_lo && d < p_hi) yv[u] = __fadd_rn(yv[u], __fmul_rn(half, kv[u]));
              }
              rd_store8(cx, y, d0, yv);
            });
            break;
          case FFB_M_EM: {
            // diffusion.py:552-559: f = drift - g^2 score (ev.c = g^2); x_mean = x + f dt; x = x_mean + g dw
            if (cx.row < nv) {
              const float* nz = a.noise ? a.noise + ((size_t)step * a.batch + row0 + cx.row) * SD : nullptr;
              rd_for_blocks(cx, [&](int d0) {
                float yv[8], k1[8], zz[8], xm[8];
                rd_load8(cx, y, d0, yv); rd_load8(cx, K1, d0, k1);
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
                  if (d0 + u < SD) saw_nan |= (yv[u] != yv[u]);
                }
                rd_store8(cx, Y0, d0, xm);
                rd_store8(cx, y, d0, yv);
              });
            }
            break;
          }
          default: break;
        }
      }
    }
    if (!cx.producer) {
      rd_gbar(cx);
      store_rows_t<RD_GTHR>(a.x_out, (a.method == FFB_M_EM) ? Y0 : y, row0, nv, SD, cx.tid);
      if (saw_nan) atomicOr(a.status, FFB_ST_NAN_SAMPLE);
      rd_gbar(cx);
    }
  }
  ENG::fini(cx);
}

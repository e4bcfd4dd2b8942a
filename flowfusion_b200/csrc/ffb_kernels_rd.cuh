// ffb_kernels_rd.cuh -- the dopri5 attempt kernel on the dual-tile tensor-core engine (ffb_engine_rd.cuh):
// fields without tangent rows (div_mode == FFB_DIV_NONE).  Compiled by ffb_rd.cu.
//
//   k_dopri5_rd      one attempted Dormand-Prince step (6 fused evaluations, FSAL, FP64 error partials,
//                    dense output at t_end)
//
// Same contract and the same FP32 statements as the single-tile kernel k_dopri5_rr (ffb_kernels_rr.cuh); the MMAs are
// issued in the same order, so the results are bit-identical to it.  Persistent: one 608-thread CTA per SM (clusters of
// RD_CLUSTER CTAs share every weight chunk by multicast); in round r, group g of CTA b integrates tile
// 2 (b + r gridDim) + g.  A group runs its tile from load to store without any CTA-wide barrier.
//
// The fixed-grid and single-evaluation kernels stay on the single-tile engine: measured on B200 (profiles/
// r02_rd_ablations.txt) the dual-tile form of k_fixed was 7-9 % slower on cfg4 / cfg5, whose per-step algebra is short
// and whose wider state does not fit shared memory beside the two A_lo images.
#pragma once
#include "ffb_engine_rd.cuh"
#include "ffb_kernels_rr.cuh"     // load_rows_t / store_rows_t, rr_fixed_slots

namespace ffb {

template <int NV>
__device__ __forceinline__ void rd_block_reduce_store(CtxD& cx, double (&v)[NV], double* out, const int (&slot)[NV], bool write) {
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
  if (cx.tid == 0 && write) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double s = 0.0;
      for (int w = 0; w < RD_GWARPS; ++w) s += cx.red()[w * FFB_NPART + q];
      out[slot[q]] = s;
    }
  }
  rd_gbar(cx);
}

// Tile schedule.  The CTAs of a cluster must walk the same sequence of weight chunks (the chunks are multicast), so
// every CTA runs the same number of rounds and both groups of every CTA take part in every round: in round r, group g of
// CTA b integrates tile 2 (b + r gridDim) + g; a tile past the end of the batch is a DUMMY tile (no valid rows: nothing
// is read, nothing is stored, the arithmetic runs on zeros).
#define FFB_RD_ROUNDS(base) \
  for (int64_t base = 2 * (int64_t)(blockIdx.x - cx.crank); base < ntiles; base += 2 * (int64_t)gridDim.x)
#define FFB_RD_TILE(base)                                                         \
  cx.active = 3u;                                                                 \
  const int64_t tile = (base) + 2 * (int64_t)cx.crank + cx.g;                     \
  const bool real_tile = tile < ntiles;                                           \
  const int64_t row0 = tile * TM;                                                 \
  const int nv = real_tile ? (int)min((int64_t)TM, a.batch - row0) : 0;

#if RD_CLUSTER > 1
#define FFB_RD_CLUSTER_ATTR __cluster_dims__(RD_CLUSTER, 1, 1)
#else
#define FFB_RD_CLUSTER_ATTR
#endif

}  // namespace ffb

// =============================================================================================
// k_dopri5_rd
// =============================================================================================
// DYN: the step (scalars, dt-scaled tableau, final flag) and the roles of the two state buffers come from the
// device-resident controller block a.ctl (ffb_control.cuh) instead of the launch arguments; a finished solve
// makes the kernel return at once.
template <int MEM, bool GEN, bool DYN>
__global__ void FFB_RD_CLUSTER_ATTR __launch_bounds__(ffb::RD_NTHR, 1) k_dopri5_rd(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_dopri5_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENG = EngineRD_<GEN, MEM>;
  if (DYN) { if (__ldg(&a.ctl->done) != 0) return; }
  // controller values are re-read (L1 hits) where they are used so that none stays in a register across the evaluations
#define FFB_STEP(x) (DYN ? __ldg(&a.ctl->x) : a.x)
#define FFB_SWAPPED() (DYN && __ldg(&a.ctl->cur) != 0)
  CtxD cx;
  ENG::init(cx, f, reinterpret_cast<float*>(a.scratch), RD_SCR_SLOTS, 6);
  const int SD = cx.SD, CD = cx.CD;
  const int bstride = f.n_calls * KMAX;
  if (!cx.producer) {
    for (int s = 0; s < 6; ++s)
      ENG::prep_beff(cx, f, DYN ? a.ctl->ev[s].tfeat : a.ev[s].tfeat, cx.beff() + s * bstride, threadIdx.x, RD_NCOMP);
    rd_allbar();
  }
  FFB_RD_ROUNDS(base) {
    FFB_RD_TILE(base)
    RD_TRACE(cx, 900);                      // debug timeline: tile start in SM cycles ...
    RD_TRACE(cx, 950);                      // ... and in nanoseconds
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
    float* X0 = rd_slot<MEM>(cx, RD_XSLOT);       // partial sums, formed while the evaluation's MMAs run (see below)
    float* X1 = rd_slot<MEM>(cx, RD_XSLOT + 1);
    for (int i = 1; i <= 6; ++i) {
      // Everything the step algebra after evaluation i needs from k_1 .. k_i (slots 0 .. i-1, all complete) is summed
      // HERE, on the compute warps' idle time right after the layer-0 operand has been handed over: the chain
      // sum_j c_j k_j is evaluated in the same order (j = 0, 1, ...) as one fused pass would, so stopping before the
      // newest term and adding it after the evaluation gives the same bits.  What remains between the last-layer
      // epilogue and the next layer-0 operand is one FMA per element instead of up to 7 slot reads.
      auto partial = [&]() {
        if (i < 6) {
          float cbi[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) cbi[j] = FFB_STEP(cb[i][j]);
          rd_for_blocks(cx, [&](int d0) {
            float kv[8], acc[8];
            rd_load8(cx, rd_slot<MEM>(cx, 0), d0, kv);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = __fmul_rn(kv[u], cbi[0]);
#pragma unroll
            for (int j = 1; j < 5; ++j) {
              if (j < i) {
                rd_load8(cx, rd_slot<MEM>(cx, j), d0, kv);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = fmaf(kv[u], cbi[j], acc[u]);
              }
            }
            rd_store8(cx, X0, d0, acc);
          });
        } else {
          float ce[6], cm[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) { ce[j] = FFB_STEP(ce[j]); cm[j] = FFB_STEP(cm[j]); }
          rd_for_blocks(cx, [&](int d0) {
            float kv[8], err[8], mid[8];
            rd_load8(cx, rd_slot<MEM>(cx, 0), d0, kv);
#pragma unroll
            for (int u = 0; u < 8; ++u) { err[u] = __fmul_rn(kv[u], ce[0]); mid[u] = __fmul_rn(kv[u], cm[0]); }
#pragma unroll
            for (int j = 1; j < 6; ++j) {
              rd_load8(cx, rd_slot<MEM>(cx, j), d0, kv);
#pragma unroll
              for (int u = 0; u < 8; ++u) { err[u] = fmaf(kv[u], ce[j], err[u]); mid[u] = fmaf(kv[u], cm[j], mid[u]); }
            }
            rd_store8(cx, X0, d0, err);
            rd_store8(cx, X1, d0, mid);
          });
        }
      };
      if constexpr (DYN) ENG::eval_ev(cx, f, EvCtl{a.ctl, i - 1}, cx.beff() + (i - 1) * bstride, i, 3u, partial);
      else ENG::eval(cx, f, a.ev[i - 1].a, a.ev[i - 1].c, a.ev[i - 1].sigma, a.ev[i - 1].sign, cx.beff() + (i - 1) * bstride, i, 3u, partial);
      if (!cx.producer && i < 6) {
        // input of stage i+1: y0 + (sum_{j<i} cb[i][j] k_j + cb[i][i] k_i)  (the 7th stage input is y1: FSAL)
        const float cbii = FFB_STEP(cb[i][i]);
        rd_for_blocks(cx, [&](int d0) {
          float y0v[8], kv[8], acc[8];
          rd_load8(cx, Y0, d0, y0v);
          rd_load8(cx, rd_slot<MEM>(cx, i), d0, kv);
          rd_load8(cx, X0, d0, acc);
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = __fadd_rn(y0v[u], fmaf(kv[u], cbii, acc[u]));
          rd_store8(cx, YC, d0, acc);
        });
      }
    }
    if (!cx.producer) {
      // the stage input holds y1, slot 6 holds f1, X0 / X1 the error / mid-point sums over k_1 .. k_6
      double v[2] = {0.0, nonfinite};
      const int final_ = FFB_STEP(final);
      float* OUT = rd_slot<MEM>(cx, 1);     // K2 of an element is dead once its error / mid sums are formed
      if (cx.row < nv) {
        const float ce6 = FFB_STEP(ce[6]), cm6 = FFB_STEP(cm[6]);
        const float dt_ = FFB_STEP(dt), xi_ = FFB_STEP(x_interp);
        rd_for_blocks(cx, [&](int d0) {
          float y0v[8], y1v[8], k0[8], kv[8], err[8], mid[8];
          rd_load8(cx, Y0, d0, y0v);
          rd_load8(cx, YC, d0, y1v);
          rd_load8(cx, rd_slot<MEM>(cx, 0), d0, k0);
          rd_load8(cx, rd_slot<MEM>(cx, 6), d0, kv);           // k7 = f1
          rd_load8(cx, X0, d0, err);
          rd_load8(cx, X1, d0, mid);
          float out[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float e = fmaf(kv[u], ce6, err[u]), m = fmaf(kv[u], cm6, mid[u]);
            const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(y0v[u]), fabsf(y1v[u]))));
            const float q = __fdiv_rn(e, tol);
            if (d0 + u < SD) v[0] += (double)q * q;
            out[u] = final_ ? dense_output(y0v[u], y1v[u], __fadd_rn(y0v[u], m), k0[u], kv[u], dt_, xi_) : 0.0f;
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
      rd_block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot, real_tile);
      if (cx.tid == 0 && real_tile) a.partials[tile * FFB_NPART + P_LP_ERR] = 0.0;
    }
  }
  ENG::fini(cx);
#undef FFB_STEP
#undef FFB_SWAPPED
}


// ffb_engine_rrt.cuh -- the chunk-pipelined tensor-core engine (ffb_engine_rr.cuh) with forward-mode
// TANGENT rows: the divergence trace of the log-likelihood paths (flow.py:122-166, 598-652;
// diffusion.py:483-503 exact, :327-334 Hutchinson) as extra rows of the same 128-row tile.
//
// Row layout.  A trajectory ("sample") owns 1 primal row P and T tangent rows (T = D for the exact
// trace: tangent j carries d(activation)/dx_j; T = 1 for Hutchinson: the tangent along the probe).
// The hidden-layer epilogue of a tangent row needs silu'(z) of its primal row.  Rows are TMEM lanes and a
// warp can only read the 32 lanes of its quarter, so samples are packed back to back and whenever a
// sample's tangents continue into the next lane quarter a DUPLICATE primal row P' is inserted at the
// start of that quarter.  Every quarter is then self-contained: the primal's z reaches its tangents with
// one warp shuffle per column -- no shared-memory gate buffer and no barrier inside a layer.
//   D = 16: 7 samples use 122 of the 128 rows (2 duplicates); T = 1: 64 (P, T) pairs.
//
// All lanes evaluate sigmoid(z_primal) themselves (uniform code, the SFU cost of a warp instruction does
// not depend on the active lanes):  primal rows  a = z sg;  tangent rows  a = D * sg (1 + z (1 - sg)).
//
// State (y, k_1..k_7, log-det) is per SAMPLE: shared-memory buffers [d][ld], ld = samples per tile
// rounded up to 4.  The primal OWNER row of a sample does the stage algebra for the state columns its
// column group owns; two CTA barriers per evaluation order it against the operand build of the duplicates.
#pragma once
#include "ffb_kernels_rr.cuh"

namespace ffb {

enum { RT_IDLE = 0, RT_OWNER = 1, RT_DUP = 2, RT_TAN = 3 };
struct RowMap { int kind, smp, tj, plane; };

// Place samples of (1 + T) rows into 4 quarters of 32 lanes; returns the samples per tile and, for `row`,
// what that row is.  Deterministic and shared by host (tile count) and device (per-thread role).
// `cap` > 0 limits the samples per tile (used when the per-sample state would not fit shared memory).
__host__ __device__ inline int rrt_rowmap(int T, int row, RowMap* out, int cap = 0) {
  int r = 0, S = 0;
  if (out) { out->kind = RT_IDLE; out->smp = 0; out->tj = -1; out->plane = row & 31; }
  for (;;) {
    int rr = r + 1;
    for (int j = 0; j < T; ++j) { if ((rr & 31) == 0) ++rr; ++rr; }
    if (rr > TM || (cap > 0 && S >= cap)) break;
    int plane = r & 31;
    if (out && r == row) { out->kind = RT_OWNER; out->smp = S; out->tj = -1; out->plane = plane; }
    ++r;
    for (int j = 0; j < T; ++j) {
      if ((r & 31) == 0) {
        plane = 0;
        if (out && r == row) { out->kind = RT_DUP; out->smp = S; out->tj = -1; out->plane = 0; }
        ++r;
      }
      if (out && r == row) { out->kind = RT_TAN; out->smp = S; out->tj = j; out->plane = plane; }
      ++r;
    }
    ++S;
  }
  return S;
}
__host__ __device__ inline int rrt_ld(int S) { return ((S + 3) & ~3) < 4 ? 4 : ((S + 3) & ~3); }

struct TanCtx {
  int kind, smp, tj, plane;     // this thread's row
  int S, T, ld;                 // samples per tile, tangents per sample, sample stride of the state buffers
  bool exact;
  // transposed activation path (<= 4 primal rows per lane quarter): lane i evaluates sigmoid for primal row i / 8 of
  // the quarter, column i % 8 of the warp's 8 columns
  bool use_tr;
  int gsrc, my_u, pbase;        // lane of that primal row; i % 8; 8 * (index of THIS row's primal among the quarter's primal rows)
  uint32_t o_klp, o_diag, o_prb;
  float* jac;                   // k_field_eval_rrt only: this tile's (sample, tangent j, output n) Jacobian rows, or nullptr
  int jac_nv;                   // valid samples of the tile
  __device__ __forceinline__ float* klp() const { return reinterpret_cast<float*>(smem_base() + o_klp); }
  __device__ __forceinline__ float* diag() const { return reinterpret_cast<float*>(smem_base() + o_diag); }
  __device__ __forceinline__ float* prb() const { return reinterpret_cast<float*>(smem_base() + o_prb); }
  __device__ __forceinline__ bool primal() const { return kind == RT_OWNER || kind == RT_DUP; }
};

constexpr int RRT_KLP_ROWS = NSLOT + 3;     // d(logp)/dt per slot, lp0, spare
constexpr int RRT_DIAG_FLOATS = 1024;       // exact: S*T <= 128 diagonal entries; Hutchinson: 16 partials per sample

__host__ __device__ inline size_t smem_layout_rrt(int SD, int CD, int ld, int hutch, int nslot, int tdim, int nbeff,
                                                  size_t* off /*[12]*/) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  size_t v[12];
  v[0] = take(sizeof(float) * TC_NSTAGE * TC_STAGE_FLOATS);
  v[1] = take(sizeof(float) * SD * ld);                            // ycur [d][sample]
  v[2] = take(sizeof(float) * (CD > 0 ? CD : 1) * ld);             // cond
  v[3] = take(sizeof(float) * NET_MAXL * KMAX);              // biases
  v[4] = take(sizeof(float) * nbeff * KMAX);                       // layer-0 bias + time features per evaluation
  v[5] = take(sizeof(float) * (tdim > 0 ? tdim : 1) * KMAX);       // time-feature rows
  v[6] = take(sizeof(double) * (RR_NCOMP / 32) * FFB_NPART);
  v[7] = take(sizeof(uint64_t) * (2 * TC_NSTAGE + 4));
  v[8] = take(sizeof(float) * (size_t)nslot * SD * ld);            // state slots [slot][d][sample]
  v[9] = take(sizeof(float) * RRT_KLP_ROWS * ld);
  v[10] = take(sizeof(float) * RRT_DIAG_FLOATS);
  v[11] = take(hutch ? sizeof(float) * SD * ld : 0);               // probes [d][sample]
  if (off) for (int i = 0; i < 12; ++i) off[i] = v[i];
  return o;
}

template <bool GEN>
struct EngineRRT_ {
  static __device__ __forceinline__ void init(CtxR& cx, TanCtx& tc, const FieldDev& f, int nslot, int nbeff) {
    tc.exact = (f.div_mode == FFB_DIV_EXACT);
    tc.jac = nullptr; tc.jac_nv = 0;
    tc.T = tc.exact ? f.net[0].x_dim : 1;
    RowMap m;
    const int row = ((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31);
    tc.S = rrt_rowmap(tc.T, row, &m, f.rrt_cap);
    tc.kind = m.kind; tc.smp = m.smp; tc.tj = m.tj; tc.plane = m.plane;
    tc.ld = rrt_ld(tc.S);
    size_t off[12];
    smem_layout_rrt(f.state_dim, f.cond_dim, tc.ld, !tc.exact, nslot, field_tdim(f), nbeff, off);
    tc.o_klp = (uint32_t)off[9]; tc.o_diag = (uint32_t)off[10]; tc.o_prb = (uint32_t)off[11];
    EngineRR::init_at(cx, f, nullptr, off);
    // which rows are primal rows -> shared memory (borrowing the trace buffer), then every thread reads its quarter
    int* rk = reinterpret_cast<int*>(tc.diag());
    if (threadIdx.x < TM) {
      RowMap mm;
      rrt_rowmap(tc.T, threadIdx.x, &mm, f.rrt_cap);
      rk[threadIdx.x] = (mm.kind == RT_OWNER || mm.kind == RT_DUP) ? 1 : 0;
    }
    __syncthreads();
    const int q0 = ((threadIdx.x >> 5) & 3) * 32, lane = threadIdx.x & 31;
    int ppos = 0, found = -1, cnt = 0, worst = 0;
    for (int l = 0; l < 32; ++l) {
      if (rk[q0 + l]) {
        if (l < tc.plane) ++ppos;
        if (cnt == (lane >> 3)) found = l;
        ++cnt;
      }
    }
    for (int qq = 0; qq < 4; ++qq) {
      int c = 0;
      for (int l = 0; l < 32; ++l) c += rk[qq * 32 + l];
      worst = max(worst, c);
    }
    tc.use_tr = (worst <= 4);
    tc.gsrc = (found >= 0) ? found : lane;
    tc.my_u = lane & 7;
    tc.pbase = 8 * min(ppos, 3);
    __syncthreads();
  }
  static __device__ __forceinline__ float* slot(const CtxR& cx, const TanCtx& tc, int s) {
    return reinterpret_cast<float*>(smem_base() + cx.o_slots) + (size_t)s * cx.SD * tc.ld;
  }

  // layer-0 operand: primal rows x | cond, tangent rows e_j (exact) or the probe (Hutchinson)
  static __device__ __forceinline__ void build_A(CtxR& cx, const TanCtx& tc, const FieldDev& f) {
    const NetDev& net = f.net[0];
    const int K0 = net.K[0], xd = net.x_dim, cd = net.c_dim, ld = tc.ld, s = tc.smp;
    const bool prim = tc.primal(), tan = (tc.kind == RT_TAN);
    const float* yc = cx.ycur() + s;
    const float* cb = cx.condb() + s;
    const float* pb = tc.prb() + s;
    for (int k0 = 0; k0 < K0; k0 += KC) {
      const int k8 = k0 + 8 * cx.cg;
      if (k8 < K0) {
        uint32_t hi[8], lo[8];
        float vx[8], vc[8], vp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {                        // unconditional loads (clamped), selects below
          const int k = k8 + j;
          vx[j] = yc[min(k, xd - 1) * ld];
          vc[j] = (cd > 0) ? cb[min(max(k - xd, 0), cd - 1) * ld] : 0.0f;
          vp[j] = tc.exact ? 0.0f : pb[min(k, xd - 1) * ld];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k8 + j;
          const float pv = (k < xd) ? vx[j] : ((k < xd + cd) ? vc[j] : 0.0f);
          const float tv = (k < xd) ? (tc.exact ? (k == tc.tj ? 1.0f : 0.0f) : vp[j]) : 0.0f;
          tf32_split(prim ? pv : (tan ? tv : 0.0f), hi[j], lo[j]);
        }
        tc_st8(cx.lane_addr + RR_COL_AHI + k8, hi);
        tc_st8(cx.lane_addr + RR_COL_ALO + k8, lo);
      }
      EngineRR::signal_chunk(cx);
    }
  }

  static __device__ __forceinline__ void hidden(CtxR& cx, const TanCtx& tc, const NetDev& net, const float* beff) {
    if (GEN) { FFB_ACT_DISPATCH(net.act, hidden_act<ACT>(cx, tc, net, beff)); }
    else hidden_act<FFB_ACT_SILU>(cx, tc, net, beff);
  }
  template <int ACT>
  static __device__ __forceinline__ void hidden_act(CtxR& cx, const TanCtx& tc, const NetDev& net, const float* beff) {
    const bool prim = tc.primal(), tan = (tc.kind == RT_TAN);
    for (int l = 0; l + 1 < net.n_layers; ++l) {
      const int nc = net.Np[l] / KC;
      const float* bias = (l == 0) ? beff : cx.sbias() + l * KMAX;
      const uint32_t dcol = cx.lane_addr + cx.dbuf * 128u + 8u * cx.cg;
      cx.dbuf ^= 1u;
      EngineRR::wait_d_ready(cx, net.K[l]);
      RR_TRACE(cx, 200 + 10 * l);
      uint32_t m[2][8];
      tc_ld8(dcol, m[0]);
#pragma unroll
      for (int ci = 0; ci < RR_NCHUNK; ++ci) {
        if (ci < nc) {
          const int c0 = KC * ci + 8 * cx.cg;
          tc_wait_ld();
          if (ci + 1 < nc) tc_ld8(dcol + KC * (ci + 1), m[(ci + 1) & 1]);
          const float4 b0 = *reinterpret_cast<const float4*>(bias + c0);
          const float4 b1 = *reinterpret_cast<const float4*>(bias + c0 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint32_t hi[8], lo[8];
          if (tc.use_tr) {
            // gather z of (primal row lane/8, column lane%8) into this lane, one sigmoid per lane, scatter back
            float zz = 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float t = __shfl_sync(0xffffffffu, __uint_as_float(m[ci & 1][u]) + bb[u], tc.gsrc);
              zz = (tc.my_u == u) ? t : zz;
            }
            float aval, gate;
            act_fwd_grad<ACT>(zz, aval, gate);
#pragma unroll
            for (int u = 0; u < 8; u += 2) {                      // packed FP32 pairs for the gate product and the split
              const float g0 = __shfl_sync(0xffffffffu, gate, tc.pbase + u), g1 = __shfl_sync(0xffffffffu, gate, tc.pbase + u + 1);
              const float av0 = __shfl_sync(0xffffffffu, aval, tc.pbase + u), av1 = __shfl_sync(0xffffffffu, aval, tc.pbase + u + 1);
              const float2 t = __fmul2_rn(make_float2(__uint_as_float(m[ci & 1][u]), __uint_as_float(m[ci & 1][u + 1])), make_float2(g0, g1));
              const float2 a = make_float2(prim ? av0 : (tan ? t.x : 0.0f), prim ? av1 : (tan ? t.y : 0.0f));
              tf32_split2(a, hi[u], hi[u + 1], lo[u], lo[u + 1]);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float d = __uint_as_float(m[ci & 1][u]);
              const float z = __shfl_sync(0xffffffffu, d + bb[u], tc.plane);      // z of this row's primal
              float av, g;
              act_fwd_grad<ACT>(z, av, g);
              const float a = prim ? av : (tan ? d * g : 0.0f);
              tf32_split(a, hi[u], lo[u]);
            }
          }
          tc_st8(cx.lane_addr + RR_COL_AHI + c0, hi);
          tc_st8(cx.lane_addr + RR_COL_ALO + c0, lo);
          EngineRR::signal_chunk(cx);
          RR_TRACE(cx, 201 + 10 * l + ci);
        }
      }
    }
  }

  // one evaluation at cx.ycur(): derivative -> slot dst, divergence -> klp[dst]
  static __device__ __forceinline__ void eval(CtxR& cx, const TanCtx& tc, const FieldDev& f, float ev_a, float ev_c,
                                              float ev_sigma, float ev_sign, const float* beff, int dst) {
    eval_ev(cx, tc, f, EvVals{ev_a, ev_c, ev_sigma, ev_sign}, beff, dst);
  }
  // EV: source of the per-evaluation scalars, fetched where they are used (see EngineRR_::eval_ev)
  // DEFER: leave this evaluation's divergence to the caller (finish_trace); `prev` / `prev_dst` (>= 0): the deferred divergence
  // of the previous evaluation, finished here once the layer-0 operand is handed over (the owner threads would otherwise idle
  // until layer 0's accumulators arrive)
  template <class EV, bool DEFER = false>
  static __device__ __forceinline__ void eval_ev(CtxR& cx, const TanCtx& tc, const FieldDev& f, const EV& ev,
                                                 const float* beff, int dst, const EV* prev = nullptr, int prev_dst = -1) {
    const NetDev& net = f.net[0];
    if (cx.warp == RR_WLOAD) { EngineRR::load_net(cx, net); return; }
    if (cx.warp == RR_WMMA) { EngineRR::mma_net(cx, net); return; }
    if (cx.producer) return;                   // spare warps
    const int ld = tc.ld, s = tc.smp, xd = net.x_dim;
    rr_bar();                                  // cx.ycur() is final for every sample of the tile
    RR_TRACE(cx, 10);
    build_A(cx, tc, f);
    RR_TRACE(cx, 11);
    if (prev_dst >= 0) finish_trace(cx, tc, f, *prev, prev_dst);
    hidden(cx, tc, net, beff);
    // ---- last layer ---------------------------------------------------------------------------------
    const int nl = net.n_layers, Nreal = net.N[nl - 1];
    const float* bias = (nl == 1) ? beff : cx.sbias() + (nl - 1) * KMAX;
    const uint32_t dcol = cx.lane_addr + cx.dbuf * 128u + 8u * cx.cg;
    cx.dbuf ^= 1u;
    EngineRR::wait_d_ready(cx, net.K[nl - 1]);
    RR_TRACE(cx, 290);
    float* kd = slot(cx, tc, dst) + s;
    const float* yc = cx.ycur() + s;
    const bool score = (f.kind == FFB_FIELD_SCORE), use_sigma = f.use_sigma != 0, has_drift = f.has_drift != 0;
    for (int c0 = 8 * cx.cg; c0 < Nreal; c0 += KC) {
      uint32_t m[8];
      tc_ld8(dcol + (uint32_t)(c0 - 8 * cx.cg), m);
      tc_wait_ld();
      if (tc.kind == RT_OWNER) {
        const float ev_sign = ev.sign();
        const float ev_a = (score && has_drift) ? ev.a() : 0.0f, ev_c = score ? ev.c() : 0.0f;
        const float ev_sigma = (score && use_sigma) ? ev.sigma() : 1.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int n = c0 + u;
          if (n < Nreal) {
            const float o = __uint_as_float(m[u]) + bias[n];
            float xd_;
            if (score) {
              const float sc = use_sigma ? __fdiv_rn(o, ev_sigma) : o;
              const float lin = has_drift ? __fmul_rn(ev_a, yc[n * ld]) : 0.0f;
              xd_ = __fsub_rn(lin, __fmul_rn(ev_c, sc));
            } else {
              xd_ = o;
            }
            kd[n * ld] = xd_ * ev_sign;
          }
        }
      } else if (tc.kind == RT_TAN) {
        if (tc.exact) {
          float v = 0.0f;
          bool mine = false;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (c0 + u == tc.tj) { v = __uint_as_float(m[u]); mine = true; }
          if (mine) tc.diag()[s * tc.T + tc.tj] = v;                      // d f_j / d x_j
          if (tc.jac != nullptr && s < tc.jac_nv) {                       // full row: d net_n / d x_tj, n = c0..c0+7
            float* jr = tc.jac + ((size_t)s * tc.T + tc.tj) * tc.T + c0;
            if (c0 + 8 <= Nreal && (tc.T & 3) == 0) {
              *reinterpret_cast<float4*>(jr) = make_float4(__uint_as_float(m[0]), __uint_as_float(m[1]), __uint_as_float(m[2]), __uint_as_float(m[3]));
              *reinterpret_cast<float4*>(jr + 4) = make_float4(__uint_as_float(m[4]), __uint_as_float(m[5]), __uint_as_float(m[6]), __uint_as_float(m[7]));
            } else {
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (c0 + u < Nreal) jr[u] = __uint_as_float(m[u]);
            }
          }
        } else {
          float part = 0.0f;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (c0 + u < xd) part = fmaf(tc.prb()[(c0 + u) * ld + s], __uint_as_float(m[u]), part);
          tc.diag()[s * 16 + (c0 >> 3)] = part;                          // e^T J e, 8 columns at a time
        }
      }
    }
    tc_fence_before();
    RR_TRACE(cx, 12);
    rr_bar();                                  // slot dst and the trace pieces are complete
    RR_TRACE(cx, 13);
    if (!DEFER) finish_trace(cx, tc, f, ev, dst);
    RR_TRACE(cx, 28);
  }

  // divergence of the evaluation whose last layer left its trace pieces in tc.diag(): sum them (fixed order) and apply the
  // field transform -> klp[dst].  Run by the samples' owner threads; with DEFER the caller runs it for the PREVIOUS evaluation
  // while layer 0 of the next one is on the tensor pipe (nothing reads klp before the end of the attempt), or right away
  // after the last evaluation.
  template <class EV>
  static __device__ __forceinline__ void finish_trace(CtxR& cx, const TanCtx& tc, const FieldDev& f, const EV& ev, int dst) {
    if (cx.producer) return;
    const int ld = tc.ld, s = tc.smp, xd = f.net[0].x_dim;
    const bool score = (f.kind == FFB_FIELD_SCORE), use_sigma = f.use_sigma != 0, has_drift = f.has_drift != 0;
    if (tc.kind == RT_OWNER && cx.cg == 0) {
      float tr = 0.0f;
      if (tc.exact) {
        // same order of additions; the loads of a group of 8 are issued together instead of one dependent load per term
        const float* dg = tc.diag() + s * tc.T;
        for (int j0 = 0; j0 < tc.T; j0 += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = dg[min(j0 + u, tc.T - 1)];
#pragma unroll
          for (int u = 0; u < 8; ++u) if (j0 + u < tc.T) tr += v[u];
        }
      }
      else { for (int g = 0; 8 * g < xd; ++g) tr += tc.diag()[s * 16 + g]; }
      float dv;
      if (score) {
        const float trs = use_sigma ? __fdiv_rn(tr, ev.sigma()) : tr;
        const float lin = has_drift ? ev.a() * (float)xd : 0.0f;
        dv = lin - ev.c() * trs;
      } else {
        dv = tr;
      }
      tc.klp()[dst * ld + s] = dv * ev.sign();
    }
  }
};
using EngineRRT = EngineRRT_<false>;

// ---- per-sample state helpers (the owner row's thread, 8 state columns at a time) -----------------------
__device__ __forceinline__ void rt_load8(const CtxR& cx, const TanCtx& tc, const float* buf, int d0, float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = buf[min(d0 + u, cx.SD - 1) * tc.ld + tc.smp];
}
__device__ __forceinline__ void rt_load8_if(const CtxR& cx, const TanCtx& tc, bool on, const float* buf, int d0, float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = on ? buf[min(d0 + u, cx.SD - 1) * tc.ld + tc.smp] : 0.0f;
}
__device__ __forceinline__ void rt_store8(const CtxR& cx, const TanCtx& tc, float* buf, int d0, const float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (d0 + u < cx.SD) buf[(d0 + u) * tc.ld + tc.smp] = v[u];
}
// [d][sample] tile buffer <- row-major global rows [row0, row0 + nv); samples nv..S-1 are zero filled
__device__ __forceinline__ void rt_load_rows(float* dst, const float* __restrict__ src, int64_t row0, int nv, int S, int ld,
                                             int D, int tid) {
  const float* __restrict__ base = src + row0 * D;
  for (int idx = tid; idx < S * D; idx += RR_NCOMP) {
    const int r = fast_div(idx, D), d = idx - r * D;
    dst[d * ld + r] = (r < nv) ? base[idx] : 0.0f;
  }
}
__device__ __forceinline__ void rt_store_rows(float* __restrict__ dst, const float* src, int64_t row0, int nv, int ld, int D,
                                              int tid) {
  float* __restrict__ base = dst + row0 * D;
  for (int idx = tid; idx < nv * D; idx += RR_NCOMP) {
    const int r = fast_div(idx, D), d = idx - r * D;
    base[idx] = src[d * ld + r];
  }
}

}  // namespace ffb

// =============================================================================================
// k_field_eval_rrt: one evaluation with the divergence (+ torchdiffeq's initial-step norms)
// =============================================================================================
template <bool GEN>
__global__ void __launch_bounds__(ffb::RR_NTHR, 1) k_field_eval_rrt(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_eval_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENGT = EngineRRT_<GEN>;
  CtxR cx; TanCtx tc;
  ENGT::init(cx, tc, f, 3, 1);
  const int SD = cx.SD, CD = cx.CD, S = tc.S, ld = tc.ld;
  const bool owner = (tc.kind == RT_OWNER);
  if (!cx.producer) {
    EngineRR::prep_beff(cx, f, a.ev.tfeat, cx.beff());
    rr_bar();
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * S;
    const int nv = (int)min((int64_t)S, a.batch - row0);
    float* F = ENGT::slot(cx, tc, 0);
    float* Y0 = ENGT::slot(cx, tc, 1);
    float* FB = ENGT::slot(cx, tc, 2);
    if (!cx.producer) {
      rt_load_rows(Y0, a.y, row0, nv, S, ld, SD, cx.tid);
      if (a.fbase) rt_load_rows(FB, a.fbase, row0, nv, S, ld, SD, cx.tid);
      if (CD) rt_load_rows(cx.condb(), a.cond, row0, nv, S, ld, CD, cx.tid);
      if (!tc.exact) rt_load_rows(tc.prb(), a.probes, row0, nv, S, ld, SD, cx.tid);
      rr_bar();
      if (owner) {
        for (int d0 = 8 * cx.cg; d0 < SD; d0 += 32) {
          float y0v[8], fb[8];
          rt_load8(cx, tc, Y0, d0, y0v);
          rt_load8_if(cx, tc, a.fbase != nullptr, FB, d0, fb);
#pragma unroll
          for (int u = 0; u < 8; ++u) y0v[u] = a.fbase ? __fadd_rn(y0v[u], __fmul_rn(a.h, fb[u])) : y0v[u];
          rt_store8(cx, tc, cx.ycur(), d0, y0v);
        }
      }
    }
    tc.jac = a.jac ? a.jac + (size_t)row0 * tc.T * tc.T : nullptr;
    tc.jac_nv = nv;
    ENGT::eval(cx, tc, f, a.ev.a, a.ev.c, a.ev.sigma, a.ev.sign, cx.beff(), 0);
    if (!cx.producer) {
      rr_bar();                                           // klp[0] of every sample is written
      double v[6] = {0, 0, 0, 0, 0, 0};   // x_y, x_f, x_df, lp_f, lp_df, c_y
      if (a.norms && owner && tc.smp < nv) {
        for (int d0 = 8 * cx.cg; d0 < SD; d0 += 32) {
          float y0v[8], fv[8], fb[8];
          rt_load8(cx, tc, Y0, d0, y0v);
          rt_load8(cx, tc, F, d0, fv);
          rt_load8_if(cx, tc, a.norms == 2, FB, d0, fb);
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
        }
        if (cx.cg == 0) {
          const float dl = tc.klp()[tc.smp];
          if (a.norms == 1) {
            const float q = __fdiv_rn(dl, a.atol);
            v[3] += (double)q * q;
          } else {
            const float q = __fdiv_rn(__fsub_rn(dl, a.dlpbase[row0 + tc.smp]), a.atol);
            v[4] += (double)q * q;
          }
        }
      }
      if (a.f) rt_store_rows(a.f, F, row0, nv, ld, SD, cx.tid);
      if (a.dlp)
        for (int s = cx.tid; s < nv; s += RR_NCOMP) a.dlp[row0 + s] = tc.klp()[s];
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
  EngineRR::fini(cx);
}

// =============================================================================================
// k_dopri5_rrt: one attempted Dormand-Prince step of (x, log-det)
// =============================================================================================
template <bool GEN, bool DYN>      // DYN: step and buffer roles from the device-resident controller block (see k_dopri5_rr)
__global__ void __launch_bounds__(ffb::RR_NTHR, 1) k_dopri5_rrt(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_dopri5_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENGT = EngineRRT_<GEN>;
  if (DYN) { if (__ldg(&a.ctl->done) != 0) return; }
  // controller values are re-read (L1 hits) where they are used so that none stays in a register across the evaluations
#define FFB_STEP(x) (DYN ? __ldg(&a.ctl->x) : a.x)
#define FFB_SWAPPED() (DYN && __ldg(&a.ctl->cur) != 0)
  CtxR cx; TanCtx tc;
  ENGT::init(cx, tc, f, NSLOT, 6);
  const int SD = cx.SD, CD = cx.CD, S = tc.S, ld = tc.ld;
  const bool owner = (tc.kind == RT_OWNER);
  float* LP0 = tc.klp() + NSLOT * ld;
  // idx / S by multiplication: the 64-bit division behind the constant costs several hundred cycles, so it is done ONCE
  // per kernel, not once per stage (S >= 2 here would not even need the special case, S = 1 does: umulhi(idx, 2^32) overflows)
  const unsigned div_s = (S == 1) ? 0u : (unsigned)((0x100000000ull + (unsigned)S - 1u) / (unsigned)S);
  auto elem = [&](int idx) { const int d = (S == 1) ? idx : (int)__umulhi((unsigned)idx, div_s); return d * ld + (idx - d * S); };
  __shared__ float cbs[36];      // the stage coefficients cb[i][j]: read once (launch arguments or the controller block)
  if (!cx.producer) {
    for (int s = 0; s < 6; ++s) EngineRR::prep_beff(cx, f, DYN ? a.ctl->ev[s].tfeat : a.ev[s].tfeat, cx.beff() + s * KMAX);
    if (cx.tid < 36) cbs[cx.tid] = FFB_STEP(cb[cx.tid / 6][cx.tid % 6]);
    rr_bar();
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * S;
    const int nv = (int)min((int64_t)S, a.batch - row0);
    float* Y0 = ENGT::slot(cx, tc, SLOT_Y0);
    double nonfinite = 0.0;
    if (!cx.producer) {
      const bool sw = FFB_SWAPPED();
      const float* __restrict__ in_lp = sw ? a.lp1 : a.lp0;
      const float* __restrict__ in_dlp = sw ? a.dlp1 : a.dlp0;
      rt_load_rows(Y0, sw ? a.y1 : a.y0, row0, nv, S, ld, SD, cx.tid);
      rt_load_rows(ENGT::slot(cx, tc, 0), sw ? a.f1 : a.f0, row0, nv, S, ld, SD, cx.tid);
      if (CD) rt_load_rows(cx.condb(), a.cond, row0, nv, S, ld, CD, cx.tid);
      if (!tc.exact) rt_load_rows(tc.prb(), a.probes, row0, nv, S, ld, SD, cx.tid);
      for (int s = cx.tid; s < S; s += RR_NCOMP) {
        LP0[s] = (s < nv) ? in_lp[row0 + s] : 0.0f;
        tc.klp()[s] = (s < nv) ? in_dlp[row0 + s] : 0.0f;
        if (!is_finite_f(LP0[s])) nonfinite += 1.0;
      }
      rr_bar();
      {
        // Stage algebra, one (state column, sample) element per thread: a tile holds few samples (7 for 16 tangents), so
        // leaving it to the samples' owner rows (28 threads, 8 strided elements each) cost ~1.9 k cycles per evaluation
        // during which the tensor pipe idles; spread over the compute threads it is a few hundred.  Same statements, same bits.
        const float c00 = FFB_STEP(cb[0][0]);
        const float* K1 = ENGT::slot(cx, tc, 0);
        float* yc = cx.ycur();
        for (int idx = cx.tid; idx < S * SD; idx += RR_NCOMP) {
          const int e = elem(idx);
          const float y0v = Y0[e];
          if (!is_finite_f(y0v)) nonfinite += 1.0;
          yc[e] = __fadd_rn(y0v, __fmul_rn(K1[e], c00));
        }
      }
    }
    for (int i = 1; i <= 6; ++i) {
      // the divergence of evaluation i - 1 is finished inside evaluation i (after its layer-0 operand is handed over), the
      // last one right after the loop: klp is only read at the end of the attempt
      if constexpr (DYN) {
        const EvCtl cur{a.ctl, i - 1}, prv{a.ctl, i - 2};
        ENGT::template eval_ev<EvCtl, true>(cx, tc, f, cur, cx.beff() + (i - 1) * KMAX, i, &prv, i > 1 ? i - 1 : -1);
        if (i == 6) ENGT::finish_trace(cx, tc, f, cur, 6);
      } else {
        const EvVals cur{a.ev[i - 1].a, a.ev[i - 1].c, a.ev[i - 1].sigma, a.ev[i - 1].sign};
        const int ip = i > 1 ? i - 2 : 0;
        const EvVals prv{a.ev[ip].a, a.ev[ip].c, a.ev[ip].sigma, a.ev[ip].sign};
        ENGT::template eval_ev<EvVals, true>(cx, tc, f, cur, cx.beff() + (i - 1) * KMAX, i, &prv, i > 1 ? i - 1 : -1);
        if (i == 6) ENGT::finish_trace(cx, tc, f, cur, 6);
      }
      if (!cx.producer && i < 6) {
        RR_TRACE(cx, 26);
        float cbi[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) cbi[j] = cbs[i * 6 + j];
        float* yc = cx.ycur();
        for (int idx = cx.tid; idx < S * SD; idx += RR_NCOMP) {
          const int e = elem(idx);
          float acc = __fmul_rn(ENGT::slot(cx, tc, 0)[e], cbi[0]);
#pragma unroll
          for (int j = 1; j < 6; ++j) acc = fmaf((j <= i) ? ENGT::slot(cx, tc, j)[e] : 0.0f, cbi[j], acc);
          yc[e] = __fadd_rn(Y0[e], acc);
        }
        RR_TRACE(cx, 27);
      }
    }
    if (!cx.producer) {
      rr_bar();                                            // klp[6] of every sample is written
      double v[3] = {0.0, 0.0, nonfinite};
      const int final_ = FFB_STEP(final);
      const bool sw = FFB_SWAPPED();
      float* OUT = ENGT::slot(cx, tc, 1);              // K2 of an element is dead once its sums are formed
      float ce[7], cm[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) { ce[j] = FFB_STEP(ce[j]); cm[j] = FFB_STEP(cm[j]); }
      const float dt_ = FFB_STEP(dt), xi_ = FFB_STEP(x_interp);
      // error / mid-point sums and the dense output, one (state column, sample) element per thread
      for (int idx = cx.tid; idx < S * SD; idx += RR_NCOMP) {
        const int e = elem(idx), sm = e % ld;
        if (sm >= nv) continue;
        const float y0v = Y0[e], y1v = cx.ycur()[e], k0 = ENGT::slot(cx, tc, 0)[e];
        float err = __fmul_rn(k0, ce[0]), mid = __fmul_rn(k0, cm[0]), kv = k0;
#pragma unroll
        for (int j = 1; j < 7; ++j) {
          kv = ENGT::slot(cx, tc, j)[e];
          err = fmaf(kv, ce[j], err); mid = fmaf(kv, cm[j], mid);
        }
        const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(y0v), fabsf(y1v))));
        const float q = __fdiv_rn(err, tol);
        v[0] += (double)q * q;
        if (final_) OUT[e] = dense_output(y0v, y1v, __fadd_rn(y0v, mid), k0, kv, dt_, xi_);
      }
      if (owner && tc.smp < nv) {
        if (cx.cg == 0) {
          // the log-det column: same formulas on (lp0, d(logp)/dt of the 7 stages)
          const int s = tc.smp;
          const float* kl = tc.klp();
          const float l0 = LP0[s];
          float acc = __fmul_rn(kl[s], FFB_STEP(cb[5][0]));
          for (int j = 1; j < 6; ++j) acc = fmaf(kl[j * ld + s], FFB_STEP(cb[5][j]), acc);
          const float l1 = __fadd_rn(l0, acc);
          float err = __fmul_rn(kl[s], ce[0]);
          for (int j = 1; j < 7; ++j) err = fmaf(kl[j * ld + s], ce[j], err);
          const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(l0), fabsf(l1))));
          const float q = __fdiv_rn(err, tol);
          v[1] += (double)q * q;
          (sw ? const_cast<float*>(a.lp0) : a.lp1)[row0 + s] = l1;
          (sw ? const_cast<float*>(a.dlp0) : a.dlp1)[row0 + s] = kl[6 * ld + s];
          if (final_) {
            float mid = __fmul_rn(kl[s], cm[0]);
            for (int j = 1; j < 7; ++j) mid = fmaf(kl[j * ld + s], cm[j], mid);
            a.lp_out[row0 + s] = dense_output(l0, l1, __fadd_rn(l0, mid), kl[s], kl[6 * ld + s], dt_, xi_);
          }
        }
      }
      rr_bar();
      rt_store_rows(sw ? const_cast<float*>(a.y0) : a.y1, cx.ycur(), row0, nv, ld, SD, cx.tid);
      rt_store_rows(sw ? const_cast<float*>(a.f0) : a.f1, ENGT::slot(cx, tc, 6), row0, nv, ld, SD, cx.tid);
      if (final_) rt_store_rows(a.y_out, OUT, row0, nv, ld, SD, cx.tid);
      const int slot[3] = {P_X_ERR, P_LP_ERR, P_NONFINITE};
      rr_block_reduce_store(cx, v, a.partials + tile * FFB_NPART, slot);
    }
  }
  EngineRR::fini(cx);
#undef FFB_STEP
#undef FFB_SWAPPED
}

// =============================================================================================
// k_fixed_rrt: fixed-grid integration of (x, log-det): euler / midpoint / rk4 (3/8 rule), whole trajectory on-chip
// =============================================================================================
template <bool GEN>
__global__ void __launch_bounds__(ffb::RR_NTHR, 1) k_fixed_rrt(const __grid_constant__ ffb::FieldDev f,
        const __grid_constant__ ffb_fixed_args a, const int64_t ntiles) {
  using namespace ffb;
  using ENGT = EngineRRT_<GEN>;
  CtxR cx; TanCtx tc;
  const int nslot = rr_fixed_slots(a.method);
  ENGT::init(cx, tc, f, nslot, 1);
  const int SD = cx.SD, CD = cx.CD, S = tc.S, ld = tc.ld;
  const bool owner = (tc.kind == RT_OWNER);
  const int nev = evals_per_step(a.method);
  const float third = (float)(1.0 / 3.0);
  float* LPC = tc.klp() + (NSLOT + 1) * ld;          // running log-det
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * S;
    const int nv = (int)min((int64_t)S, a.batch - row0);
    float* Y0 = ENGT::slot(cx, tc, nslot - 1);
    const float* K1 = ENGT::slot(cx, tc, 0);
    const float* K2 = ENGT::slot(cx, tc, nslot > 1 ? 1 : 0);
    const float* K3 = ENGT::slot(cx, tc, nslot > 2 ? 2 : 0);
    const float* K4 = ENGT::slot(cx, tc, nslot > 3 ? 3 : 0);
    if (!cx.producer) {
      rt_load_rows(cx.ycur(), a.x0, row0, nv, S, ld, SD, cx.tid);
      if (CD) rt_load_rows(cx.condb(), a.cond, row0, nv, S, ld, CD, cx.tid);
      if (!tc.exact) rt_load_rows(tc.prb(), a.probes, row0, nv, S, ld, SD, cx.tid);
      for (int s = cx.tid; s < S; s += RR_NCOMP) LPC[s] = (s < nv && a.lp0) ? a.lp0[row0 + s] : 0.0f;
    }
    for (int step = 0; step < a.nsteps; ++step) {
      const float* st = a.step_table + (size_t)step * FFB_STEP_STRIDE;
      const ffb_eval_scalars* ev = a.ev_table + (size_t)step * nev;
      const float dt = st[0], half = st[3];
      for (int e = 0; e < nev; ++e) {
        float ea = 0.f, ec = 0.f, es = 1.f, esg = 1.f;
        if (!cx.producer) {
          ea = ev[e].a; ec = ev[e].c; es = ev[e].sigma; esg = ev[e].sign;
          EngineRR::prep_beff(cx, f, ev[e].tfeat, cx.beff());      // made visible by the barrier that opens the evaluation
        }
        ENGT::eval(cx, tc, f, ea, ec, es, esg, cx.beff(), e);
        if (cx.producer || !owner) continue;
        float* y = cx.ycur();
        for (int d0 = 8 * cx.cg; d0 < SD; d0 += 32) {
          float yv[8], y0v[8], k1[8], k2[8], k3[8], k4[8];
          rt_load8(cx, tc, K1, d0, k1);
          if (a.method == FFB_M_EULER) {
            rt_load8(cx, tc, y, d0, yv);
#pragma unroll
            for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, k1[u]));
          } else if (a.method == FFB_M_MIDPOINT) {
            if (e == 0) {
              rt_load8(cx, tc, y, d0, yv);
              rt_store8(cx, tc, Y0, d0, yv);
#pragma unroll
              for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(k1[u], half));
            } else {
              rt_load8(cx, tc, Y0, d0, yv); rt_load8(cx, tc, K2, d0, k2);
#pragma unroll
              for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(yv[u], __fmul_rn(dt, k2[u]));
            }
          } else {                                                   // rk4, 3/8 rule
            if (e == 0) {
              rt_load8(cx, tc, y, d0, y0v);
              rt_store8(cx, tc, Y0, d0, y0v);
#pragma unroll
              for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(__fmul_rn(dt, k1[u]), third));
            } else if (e == 1) {
              rt_load8(cx, tc, Y0, d0, y0v); rt_load8(cx, tc, K2, d0, k2);
#pragma unroll
              for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(dt, __fsub_rn(k2[u], __fmul_rn(k1[u], third))));
            } else if (e == 2) {
              rt_load8(cx, tc, Y0, d0, y0v); rt_load8(cx, tc, K2, d0, k2); rt_load8(cx, tc, K3, d0, k3);
#pragma unroll
              for (int u = 0; u < 8; ++u) yv[u] = __fadd_rn(y0v[u], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[u], k2[u]), k3[u])));
            } else {
              rt_load8(cx, tc, Y0, d0, y0v); rt_load8(cx, tc, K2, d0, k2); rt_load8(cx, tc, K3, d0, k3); rt_load8(cx, tc, K4, d0, k4);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float sum = __fadd_rn(__fadd_rn(k1[u], __fmul_rn(3.0f, __fadd_rn(k2[u], k3[u]))), k4[u]);
                yv[u] = __fadd_rn(y0v[u], __fmul_rn(__fmul_rn(sum, dt), 0.125f));
              }
            }
          }
          rt_store8(cx, tc, y, d0, yv);
        }
        if (cx.cg == 0) {                                          // the log-det column, same stage formulas
          const int s = tc.smp;
          const float* kl = tc.klp();
          if (a.method == FFB_M_EULER) LPC[s] = __fadd_rn(LPC[s], __fmul_rn(dt, kl[s]));
          else if (a.method == FFB_M_MIDPOINT && e == 1) LPC[s] = __fadd_rn(LPC[s], __fmul_rn(dt, kl[ld + s]));
          else if (a.method == FFB_M_RK4 && e == 3) {
            const float sum = __fadd_rn(__fadd_rn(kl[s], __fmul_rn(3.0f, __fadd_rn(kl[ld + s], kl[2 * ld + s]))), kl[3 * ld + s]);
            LPC[s] = __fadd_rn(LPC[s], __fmul_rn(__fmul_rn(sum, dt), 0.125f));
          }
        }
      }
    }
    if (!cx.producer) {
      rr_bar();
      rt_store_rows(a.x_out, cx.ycur(), row0, nv, ld, SD, cx.tid);
      if (a.lp_out)
        for (int s = cx.tid; s < nv; s += RR_NCOMP) a.lp_out[row0 + s] = LPC[s];
      rr_bar();
    }
  }
  EngineRR::fini(cx);
}

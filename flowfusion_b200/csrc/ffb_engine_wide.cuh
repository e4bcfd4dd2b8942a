// ffb_engine_wide.cuh -- tile engine for networks the tensor-core engines do not hold: layer widths above 128 (up to
// FFB_MAX_WIDTH = 512) and more than 8 Linear layers (up to FFB_MAX_LAYERS = 16).  The reference takes any `units=[...]`
// (diffusion.py:32-40, flow.py:37-44, symplectic.py:25); 256- and 512-wide score networks are ordinary.
//
// A 128-row tile of 512-wide FP32 activations is 256 KB, more than an SM's shared memory, and its 3xTF32 operand pair is
// twice that, more than tensor memory and shared memory together: a wide layer cannot be row-resident the way the
// 128-wide engines keep it.  This engine therefore walks the tile in PASSES of 32 rows on the FP32 pipe (packed FFMA2),
// with the same interface as the other engines (Ctx / init / eval / fini), so the engine-generic kernels
// (ffb_kernels_generic.cuh: k_field_eval, k_dopri5, k_fixed) run on it unchanged:
//   * 8 compute warps + 1 producer warp.  Warp w owns rows 4w..4w+3 of a pass from the layer-0 operand to the last layer:
//     it reads only its own rows of the k-major activation buffer and writes only its own rows of the other one
//     (ping-pong), so there is NO block barrier inside a pass -- only the weight ring's mbarriers.
//   * a thread is 4 rows x C columns (C = 4, 2, 1 for chunks of 128, 64, 32 output columns): per k one broadcast LDS.128
//     of the rows, one LDS.128 of the columns, 2C FFMA2.  Columns of a chunk are interleaved (lane tx owns n = j*32 + tx)
//     so the transposed float4 stores of the outputs are conflict-free.
//   * weights stream from L2 through a 4 x 8 KB ring (16 k-rows of one 128-column chunk per stage, cp.async.bulk), once
//     per pass; a layer wider than 128 is walked chunk by chunk against the same input rows.
//   * forward-mode tangent rows (divergence trace) sit in the SAME thread as their primal row -- the 4 rows of a warp
//     are 4 samples (no divergence), 2 x (sample, Hutchinson tangent) or (sample, 3 of its tangents) -- so the gate
//     act'(z) reaches the tangents in registers: no gate buffer, no barrier.  A sample with T tangents takes ceil(T/3)
//     row groups, its primal row recomputed in each (75 % of the rows carry tangents).
//   * every activation of ffb200.h (SiLU, Tanh, ReLU, Softplus, GELU) through one compile-time dispatch per pass.
// Arithmetic: plain FP32 FMA chains in k order -- the same numbers as the FFMA2 debug engine (ffb_engine.cuh).
#pragma once
#include "ffb_engine.cuh"

namespace ffb {

constexpr int WD_R = 32;                          // rows of a pass
constexpr int WD_RS = WD_R + 4;                   // row stride (floats) of the k-major activation buffers
constexpr int WD_CW = 128;                        // output columns of a weight chunk
constexpr int WD_KC = 16;                         // k rows per ring stage
constexpr int WD_NSTAGE = 4;
constexpr int WD_STAGE_FLOATS = WD_KC * WD_CW;    // 8 KB

struct CtxW {
  uint32_t o_act, o_ring, o_ycur, o_cond, o_prb, o_tan, o_beff, o_klp, o_red, o_bar, o_net, o_slots;
  int slots_smem;
  int maxk;
  __device__ __forceinline__ float* act() const { return reinterpret_cast<float*>(smem_base() + o_act); }
  __device__ __forceinline__ float* stage_buf() const { return act(); }    // free between evaluations
  __device__ __forceinline__ float* ring() const { return reinterpret_cast<float*>(smem_base() + o_ring); }
  __device__ __forceinline__ float* ycur() const { return reinterpret_cast<float*>(smem_base() + o_ycur); }
  __device__ __forceinline__ float* condb() const { return reinterpret_cast<float*>(smem_base() + o_cond); }
  __device__ __forceinline__ float* prb() const { return reinterpret_cast<float*>(smem_base() + o_prb); }
  __device__ __forceinline__ float* tan() const { return reinterpret_cast<float*>(smem_base() + o_tan); }
  __device__ __forceinline__ float* beff() const { return reinterpret_cast<float*>(smem_base() + o_beff); }
  __device__ __forceinline__ float* klp() const { return reinterpret_cast<float*>(smem_base() + o_klp); }
  __device__ __forceinline__ double* red() const { return reinterpret_cast<double*>(smem_base() + o_red); }
  __device__ __forceinline__ uint64_t* full() const { return reinterpret_cast<uint64_t*>(smem_base() + o_bar); }
  __device__ __forceinline__ uint64_t* empty() const { return full() + WD_NSTAGE; }
  __device__ __forceinline__ const WideNet& net(int c) const { return reinterpret_cast<const WideNet*>(smem_base() + o_net)[c]; }
  float* scr;
  int S, T, SD, CD;
  int tid, lane, warp;
  bool producer;
  int stage;
  uint32_t phase;
};

template <bool SS>
__device__ __forceinline__ float* slot_ptr_t(const CtxW& cx, int slot) {
  if (SS) return reinterpret_cast<float*>(smem_base() + cx.o_slots) + (size_t)slot * cx.SD * LDA;
  return cx.scr + (size_t)slot * cx.SD * LDA;
}

// H = 32-row halves of a pass (1 or 2): with two halves a thread owns 8 rows x C columns, which halves the weight-operand
// LDS traffic and the operand-duplication MOVs per FFMA2 and doubles the independent work per warp; the activation buffers
// double with it, so the launchers take H = 2 only when the field still fits shared memory (widths up to 256).
__host__ __device__ inline size_t smem_layout_wide(int SD, int CD, int T, int hutch, int slots_smem, int maxk, size_t* off /*[12]*/,
                                                   int H = 1) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  size_t v[12];
  const size_t act_floats = (size_t)2 * maxk * (32 * H + 4), stage_floats = (size_t)SD * LDA;
  v[0] = take(sizeof(float) * (act_floats > stage_floats ? act_floats : stage_floats));   // activations (ping-pong) / staging
  v[1] = take(sizeof(float) * WD_NSTAGE * WD_STAGE_FLOATS);       // weight ring
  v[2] = take(sizeof(float) * SD * LDA);                          // ycur
  v[3] = take(sizeof(float) * (CD > 0 ? CD : 1) * LDA);           // cond
  v[4] = take(hutch ? sizeof(float) * SD * LDA : 0);              // probes
  v[5] = take(T > 0 ? sizeof(float) * SD * LDA : 0);              // tangent outputs: diagonal [j][s] (exact) or J e [d][s] (Hutchinson)
  v[6] = take(sizeof(float) * FFB_MAX_WIDTH);                     // layer-0 bias + time features
  v[7] = take(sizeof(float) * (NSLOT + 2) * TM);                  // klp
  v[8] = take(sizeof(double) * 8 * FFB_NPART);                    // block-reduction scratch
  v[9] = take(sizeof(uint64_t) * 2 * WD_NSTAGE);                  // barriers
  v[10] = take(2 * sizeof(WideNet));                              // network descriptors
  v[11] = take(slots_smem ? sizeof(float) * NSLOT * SD * LDA : 0);
  if (off) for (int i = 0; i < 12; ++i) off[i] = v[i];
  return o;
}

// GEN = true: every activation of ffb200.h through a compile-time dispatch per pass; false: SiLU only (the two-half variant is
// instantiated for SiLU alone to keep the build time of this translation unit in check)
template <int H, bool GEN = true>
struct EngineWideT {
  using Ctx = CtxW;
  static constexpr int NTHR = ffb::NTHR;      // 8 compute warps + the producer warp
  static constexpr int RS = 32 * H + 4;       // row stride (floats) of the k-major activation buffers

  static __device__ __forceinline__ void init(CtxW& cx, const FieldDev& f, float* scratch) {
    const int T = (f.div_mode == FFB_DIV_EXACT) ? f.net[0].x_dim : (f.div_mode == FFB_DIV_HUTCH ? 1 : 0);
    size_t off[12];
    smem_layout_wide(f.state_dim, f.cond_dim, T, f.div_mode == FFB_DIV_HUTCH, f.slots_smem, f.wide_maxk, off, H);
    cx.o_act = (uint32_t)off[0]; cx.o_ring = (uint32_t)off[1]; cx.o_ycur = (uint32_t)off[2]; cx.o_cond = (uint32_t)off[3];
    cx.o_prb = (uint32_t)off[4]; cx.o_tan = (uint32_t)off[5]; cx.o_beff = (uint32_t)off[6]; cx.o_klp = (uint32_t)off[7];
    cx.o_red = (uint32_t)off[8]; cx.o_bar = (uint32_t)off[9]; cx.o_net = (uint32_t)off[10]; cx.o_slots = (uint32_t)off[11];
    cx.slots_smem = f.slots_smem;
    cx.maxk = f.wide_maxk;
    cx.T = T; cx.S = TM / (1 + T); cx.SD = f.state_dim; cx.CD = f.cond_dim;
    cx.tid = threadIdx.x; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
    cx.producer = (cx.warp == NCOMP / 32);
    cx.scr = scratch + (size_t)blockIdx.x * NSLOT * f.state_dim * LDA;
    cx.stage = 0;
    cx.phase = cx.producer ? 1u : 0u;
    if (threadIdx.x == 0) {
      for (int s = 0; s < WD_NSTAGE; ++s) { mbar_init(&cx.full()[s], 1); mbar_init(&cx.empty()[s], NCOMP / 32); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = 0; c < f.n_calls; ++c) {
      const int* src = reinterpret_cast<const int*>(f.wide[c]);
      int* dstp = reinterpret_cast<int*>(smem_base() + cx.o_net) + c * (int)(sizeof(WideNet) / sizeof(int));
      for (int i = threadIdx.x; i < (int)(sizeof(WideNet) / sizeof(int)); i += NTHR) dstp[i] = src[i];
    }
    __syncthreads();
  }
  static __device__ __forceinline__ void fini(CtxW&) {}

  static __device__ __forceinline__ void advance(CtxW& cx) {
    if (++cx.stage == WD_NSTAGE) { cx.stage = 0; cx.phase ^= 1u; }
  }

  // ---- producer warp: the whole network once per pass, chunk by chunk ---------------------------------------
  static __device__ __forceinline__ void produce(CtxW& cx, const WideNet& net, int npass) {
    for (int p = 0; p < npass; ++p)
      for (int l = 0; l < net.n_layers; ++l) {
        const int K = net.K[l], Np = net.Np[l], CW = min(Np, WD_CW);
        for (int nc = 0; nc < Np / CW; ++nc) {
          const float* base = net.W[l] + (size_t)nc * K * CW;
          for (int k0 = 0; k0 < K; k0 += WD_KC) {
            const int rows = min(WD_KC, K - k0);
            if (cx.lane == 0) {
              mbar_wait(&cx.empty()[cx.stage], cx.phase);
              const uint32_t bytes = (uint32_t)(rows * CW) * sizeof(float);
              mbar_expect_tx(&cx.full()[cx.stage], bytes);
              bulk_g2s(cx.ring() + cx.stage * WD_STAGE_FLOATS, base + (size_t)k0 * CW, bytes, &cx.full()[cx.stage]);
            }
            advance(cx);
          }
        }
      }
  }

  // ---- this thread's 4 H rows x C columns of one chunk: acc[2 h + i][j] = (rows 2i, 2i+1 of half h) x column j -----------
  template <int C>
  static __device__ __forceinline__ void gemm_chunk(CtxW& cx, const float* actin, int K, float2 (&acc)[2 * H][C]) {
#pragma unroll
    for (int i = 0; i < 2 * H; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) acc[i][j] = make_float2(0.f, 0.f);
    constexpr int CW = 32 * C;
    for (int k0 = 0; k0 < K; k0 += WD_KC) {
      const int rows = min(WD_KC, K - k0);
      mbar_wait(&cx.full()[cx.stage], cx.phase);
      const float* ap = actin + (size_t)k0 * RS + cx.warp * 4;
      const float* bp = cx.ring() + cx.stage * WD_STAGE_FLOATS + cx.lane * C;
#pragma unroll 4
      for (int kk = 0; kk < rows; ++kk) {
        float b[C];
        load_b<C>(b, bp + kk * CW);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float4 a = *reinterpret_cast<const float4*>(ap + kk * RS + 32 * h);
          const float2 a01 = make_float2(a.x, a.y), a23 = make_float2(a.z, a.w);
#pragma unroll
          for (int j = 0; j < C; ++j) {
            const float2 bb = make_float2(b[j], b[j]);
            acc[2 * h][j] = __ffma2_rn(a01, bb, acc[2 * h][j]);
            acc[2 * h + 1][j] = __ffma2_rn(a23, bb, acc[2 * h + 1][j]);
          }
        }
      }
      __syncwarp();
      if (cx.lane == 0) mbar_arrive(&cx.empty()[cx.stage]);
      advance(cx);
    }
  }

  // what the 4 rows of row group g are: sample (or -1 = dead row) and tangent index (-1 = primal row)
  static __device__ __forceinline__ void rowmap(int T, int S, int ngt, int g, int (&s)[4], int (&jt)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int smp, j;
      if (T == 0) { smp = g * 4 + i; j = -1; }
      else if (T == 1) { smp = g * 2 + (i >> 1); j = (i & 1) ? 0 : -1; }
      else {
        smp = g / ngt;
        j = (i == 0) ? -1 : (g - smp * ngt) * 3 + i - 1;
        if (j >= T) smp = -1;
      }
      s[i] = (smp >= 0 && smp < S) ? smp : -1;
      jt[i] = j;
    }
  }

  // one chunk of one layer for this warp's rows: contraction, then the hidden-layer epilogue (activation; tangent rows
  // gated by act'(z) of the primal row in the same thread) or the last layer's (field transform / trace entries)
  template <int C, int ACT, bool SS>
  static __device__ __forceinline__ void chunk(CtxW& cx, const FieldDev& f, const WideNet& net, const ffb_eval_scalars& ev,
                                               int c, int dst, const float* actin, float* actout, int l, int nc,
                                               const int (&s)[H][4], const int (&jt)[H][4]) {
    float2 acc[2 * H][C];
    gemm_chunk<C>(cx, actin, net.K[l], acc);
    const bool last = (l == net.n_layers - 1);
    const float* bias = (l == 0) ? cx.beff() : net.b[l];
    const int T = cx.T;
    const int Dout = net.N[l];
    float* kd = slot_ptr_t<SS>(cx, dst) + f.out_off[c] * LDA;
    const float* yc = cx.ycur() + f.out_off[c] * LDA;
    const float sgn = ev.sign * f.out_sign[c];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float v[4][C];
#pragma unroll
      for (int j = 0; j < C; ++j) {
        v[0][j] = acc[2 * h][j].x; v[1][j] = acc[2 * h][j].y; v[2][j] = acc[2 * h + 1][j].x; v[3][j] = acc[2 * h + 1][j].y;
      }
      if (!last) {
#pragma unroll
        for (int j = 0; j < C; ++j) {
          const int n = nc * (32 * C) + j * 32 + cx.lane;
          const float bj = bias[n];
          if (T == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i][j] = act_fwd<ACT>(v[i][j] + bj);
          } else if (T == 1) {
            float a, g;
            act_fwd_grad<ACT>(v[0][j] + bj, a, g); v[0][j] = a; v[1][j] *= g;
            act_fwd_grad<ACT>(v[2][j] + bj, a, g); v[2][j] = a; v[3][j] *= g;
          } else {
            float a, g;
            act_fwd_grad<ACT>(v[0][j] + bj, a, g);
            v[0][j] = a; v[1][j] *= g; v[2][j] *= g; v[3][j] *= g;
          }
          *reinterpret_cast<float4*>(actout + (size_t)n * RS + 32 * h + cx.warp * 4) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
        }
        continue;
      }
      // last layer (one chunk: its width is at most the state's 128 columns)
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const int n = j * 32 + cx.lane;
        if (n >= Dout) continue;
        const float bj = bias[n];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (s[h][i] < 0) continue;
          if (jt[h][i] < 0) {
            const float o = v[i][j] + bj;
            float xd_;
            if (f.kind == FFB_FIELD_SCORE) {
              const float sc = f.use_sigma ? __fdiv_rn(o, ev.sigma) : o;
              const float lin = f.has_drift ? __fmul_rn(ev.a, yc[n * LDA + s[h][i]]) : 0.0f;
              xd_ = __fsub_rn(lin, __fmul_rn(ev.c, sc));
            } else {
              xd_ = o;
            }
            kd[n * LDA + s[h][i]] = xd_ * sgn;
          } else if (f.div_mode == FFB_DIV_HUTCH) {
            cx.tan()[n * LDA + s[h][i]] = v[i][j];
          } else if (n == jt[h][i]) {
            cx.tan()[n * LDA + s[h][i]] = v[i][j];
          }
        }
      }
    }
  }

  template <int ACT, bool SS>
  static __device__ __forceinline__ void passes(CtxW& cx, const FieldDev& f, const WideNet& net, const ffb_eval_scalars& ev,
                                                int c, int dst, int npass, int ngt) {
    const int T = cx.T, S = cx.S;
    const int K0 = net.K[0], xd = net.x_dim, cd = net.c_dim;
    const float* ycin = cx.ycur() + f.in_off[c] * LDA;
    for (int p = 0; p < npass; ++p) {
      int s[H][4], jt[H][4];
#pragma unroll
      for (int h = 0; h < H; ++h) rowmap(T, S, ngt, (p * H + h) * (NCOMP / 32) + cx.warp, s[h], jt[h]);
      float* A = cx.act();
      float* B = cx.act() + (size_t)cx.maxk * RS;
      // layer-0 operand of this warp's rows: [x | cond | zero pad]; tangent rows: one-hot (exact) or the probe (Hutchinson)
      for (int k = cx.lane; k < K0; k += 32) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float val[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x = 0.0f;
            if (s[h][i] >= 0) {
              if (jt[h][i] < 0) {
                if (k < xd) x = ycin[k * LDA + s[h][i]];
                else if (k < xd + cd) x = cx.condb()[(k - xd) * LDA + s[h][i]];
              } else if (k < xd) {
                x = (f.div_mode == FFB_DIV_EXACT) ? (k == jt[h][i] ? 1.0f : 0.0f) : cx.prb()[k * LDA + s[h][i]];
              }
            }
            val[i] = x;
          }
          *reinterpret_cast<float4*>(A + (size_t)k * RS + 32 * h + cx.warp * 4) = make_float4(val[0], val[1], val[2], val[3]);
        }
      }
      __syncwarp();
      for (int l = 0; l < net.n_layers; ++l) {
        const int Np = net.Np[l], CW = min(Np, WD_CW);
        for (int nc = 0; nc < Np / CW; ++nc) {
          if (CW == 128) chunk<4, ACT, SS>(cx, f, net, ev, c, dst, A, B, l, nc, s, jt);
          else if (CW == 64) chunk<2, ACT, SS>(cx, f, net, ev, c, dst, A, B, l, nc, s, jt);
          else chunk<1, ACT, SS>(cx, f, net, ev, c, dst, A, B, l, nc, s, jt);
        }
        __syncwarp();
        float* t = A; A = B; B = t;
      }
    }
  }

  // ---- evaluate the vector field at cx.ycur(); derivative -> slot dst, divergence -> klp[dst] ------------------
  template <bool SS>
  static __device__ __forceinline__ void eval(CtxW& cx, const FieldDev& f, const ffb_eval_scalars& ev, int dst,
                                              unsigned call_mask = 3u) {
    for (int c = 0; c < f.n_calls; ++c) {
      if (!((call_mask >> c) & 1u)) continue;
      const WideNet& net = cx.net(c);
      const int T = cx.T, S = cx.S;
      const int ngt = (T + 2) / 3;
      const int ngroups = (T == 0) ? (S + 3) / 4 : (T == 1 ? (S + 1) / 2 : S * ngt);
      const int npass = (ngroups + H * (NCOMP / 32) - 1) / (H * (NCOMP / 32));
      if (cx.producer) { produce(cx, net, npass); continue; }
      const int Np0 = net.Np[0];
      for (int n = cx.tid; n < Np0; n += NCOMP) {
        float b = net.b[0][n];
        for (int j = 0; j < net.t_dim; ++j) b = fmaf(net.Wt[(size_t)j * Np0 + n], ev.tfeat[j], b);
        cx.beff()[n] = b;
      }
      bar_compute();
      if constexpr (GEN) {
        FFB_ACT_DISPATCH(net.act, (passes<ACT, SS>(cx, f, net, ev, c, dst, npass, ngt)));
      } else {
        passes<FFB_ACT_SILU, SS>(cx, f, net, ev, c, dst, npass, ngt);
      }
      bar_compute();                        // derivative slot and tangent outputs complete
      if (T > 0) {
        const int xd = net.x_dim;
        for (int s = cx.tid; s < S; s += NCOMP) {
          float tr = 0.0f;
          if (f.div_mode == FFB_DIV_EXACT) {
            for (int j = 0; j < T; ++j) tr += cx.tan()[j * LDA + s];
          } else {
            for (int d = 0; d < xd; ++d) tr = fmaf(cx.prb()[d * LDA + s], cx.tan()[d * LDA + s], tr);
          }
          float dv;
          if (f.kind == FFB_FIELD_SCORE) {
            const float trs = f.use_sigma ? __fdiv_rn(tr, ev.sigma) : tr;
            const float lin = f.has_drift ? ev.a * (float)xd : 0.0f;
            dv = lin - ev.c * trs;
          } else {
            dv = tr;
          }
          cx.klp()[dst * TM + s] = dv * ev.sign;
        }
        bar_compute();
      }
    }
  }
};

using EngineWide = EngineWideT<1, true>;      // the training / Hamiltonian kernels (csrc/ffb_train.cu) use its contraction and ring

}  // namespace ffb

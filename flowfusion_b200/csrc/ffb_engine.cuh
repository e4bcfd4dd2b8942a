// ffb_engine.cuh -- the per-CTA "tile engine": a fused MLP forward (+ forward-mode tangents)
// over a tile of 128 rows, shared by every integrator kernel in ffb_kernels.cu.
//
// Design (B200 / sm_100a):
//   * one persistent CTA per SM: 8 compute warps (256 threads) + 1 producer warp;
//   * activations live in shared memory k-major: act[k][row] (stride LDA = 132 floats), so a
//     thread reads 4 consecutive rows with one LDS.128 and the layer output is written back
//     in place, transposed, without bank conflicts;
//   * weights are packed once per network (ffb_net_create) into k-major [K][Np] images with
//     a column permutation that makes each thread's columns contiguous; the producer warp
//     streams them from L2 with cp.async.bulk (UBLKCP) into a 3-stage ring guarded by
//     full/empty mbarriers -- the 4x128 networks (221-471 KB) do not fit beside the
//     activations, and L2 -> SMEM traffic is < 5 B/clk/SM;
//   * the contraction runs on the FP32 pipe with packed FFMA2 (fma.rn.f32x2, the only way to
//     reach the 128 FMA/clk/SM FP32 rate on sm_100): 8x8 register tiles, 32 FFMA2 + 4 LDS.128
//     per k per thread;
//   * a row is either a trajectory ("primal") or one forward-mode tangent of a trajectory;
//     tangent rows skip the bias and are gated by silu'(z) of their primal row, which makes the
//     divergence trace (flow.py:157-161, diffusion.py:483-503) one more block of GEMM rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ffb200.h"

namespace ffb {

constexpr int TM = FFB_TILE_ROWS;        // rows per tile
constexpr int LDA = TM + 4;              // smem row stride (floats) of every k-major buffer
constexpr int KMAX = 128;                // widest layer of the tensor-core / FFMA tile engines (wider: ffb_engine_wide.cuh)
constexpr int NET_MAXL = 8;              // Linear layers per network on those engines (deeper: ffb_engine_wide.cuh)
constexpr int NCOMP = 256;               // compute threads
constexpr int NTHR = NCOMP + 32;         // + producer warp
constexpr int KC = 32;                   // weight rows per ring stage
constexpr int NSTAGE = 3;
constexpr int CHUNK_FLOATS = KC * KMAX;  // 16 KB per stage
constexpr int NSLOT = 8;                 // per-CTA state slots in global scratch: Y0, K1..K7
constexpr int SLOT_Y0 = 7;

struct NetDev {
  int n_layers;
  int t_dim, x_dim, c_dim;
  int K[NET_MAXL];       // rows of the packed image (layer 0: x_dim + c_dim rounded up to 4)
  int N[NET_MAXL];       // real output width
  int Np[NET_MAXL];      // padded output width: 16, 32, 64 or 128
  const float* W[NET_MAXL];   // packed [K][Np]
  const float* b[NET_MAXL];   // packed [Np]
  const float* Wt;                  // packed time rows [t_dim][Np[0]]
  int act;                          // FFB_ACT_* of the hidden layers
};

// One network as the wide engine (ffb_engine_wide.cuh) reads it: any number of layers up to FFB_MAX_LAYERS, widths up to
// FFB_MAX_WIDTH.  Lives in device memory (FieldDev carries a pointer), copied to shared memory by the engine.
constexpr int WD_MAXL = FFB_MAX_LAYERS;
struct WideNet {
  int n_layers, t_dim, x_dim, c_dim, act, max_k;   // max_k: widest operand (rows of a packed image) of this network
  int K[WD_MAXL];            // rows of the packed image (layer 0: x_dim + c_dim rounded up to 4; else Np of the layer before)
  int N[WD_MAXL];            // real output width
  int Np[WD_MAXL];           // padded output width: 32, 64, 128 or a multiple of 128
  const float* W[WD_MAXL];   // [Np / CW][K][CW] with CW = min(Np, 128), columns of a chunk interleaved (p = tx * C + j <-> n = j * 32 + tx)
  const float* b[WD_MAXL];   // [Np], real column order, zero padded
  const float* Wt;           // [t_dim][Np[0]] time rows of layer 0, real column order
};

struct FieldDev {
  int n_calls;
  NetDev net[2];             // networks that only fit the wide engine: a one-layer summary (dims, activation), no images
  const WideNet* wide[2];    // device pointers
  int wide_maxk;             // widest operand over the field's networks (sizes the wide engine's activation buffers)
  int in_off[2], out_off[2];
  float out_sign[2];
  int state_dim, cond_dim, kind, use_sigma, has_drift, div_mode;
  int slots_smem;   // 1: the Y0 / K1..K7 state slots live in shared memory, 0: in the global scratch
  int rrt_cap;      // tangent-row engine: samples per tile (0 = as many as fit the 128 rows)
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk async copy, named barrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP); completes `bytes` on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// barrier over the 256 compute threads only (the producer warp never joins)
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// activation
// ---------------------------------------------------------------------------------------------
// sigmoid(z) = 1 / (1 + 2^(-z log2 e)) on the SFU: MUFU.EX2 + MUFU.RCP (each <= 2 ulp), 3 FP32 ops.
// z -> -inf gives rcp(+inf) = 0, z -> +inf gives rcp(1) = 1: no range fix-ups needed.
__device__ __forceinline__ float sigmoidf_fast(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// hidden-layer activation a(z) and its derivative g(z) (the gate of the tangent rows), selected at compile time
template <int ACT>
__device__ __forceinline__ void act_fwd_grad(float z, float& a, float& g) {
  if (ACT == FFB_ACT_SILU) {
    const float sg = sigmoidf_fast(z);
    a = z * sg; g = sg * (1.0f + z * (1.0f - sg));
  } else if (ACT == FFB_ACT_TANH) {
    const float t = fmaf(2.0f, sigmoidf_fast(2.0f * z), -1.0f);       // tanh z = 2 sigmoid(2 z) - 1
    a = t; g = 1.0f - t * t;
  } else if (ACT == FFB_ACT_RELU) {
    a = fmaxf(z, 0.0f); g = z > 0.0f ? 1.0f : 0.0f;
  } else if (ACT == FFB_ACT_SOFTPLUS) {
    a = z > 20.0f ? z : log1pf(expf(z));                               // torch: beta = 1, threshold = 20
    g = z > 20.0f ? 1.0f : sigmoidf_fast(z);
  } else {                                                             // GELU, erf form
    const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
    a = z * cdf; g = cdf + z * (0.39894228040143268f * expf(-0.5f * z * z));
  }
}
template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
  if (ACT == FFB_ACT_SILU) return z * sigmoidf_fast(z);
  float a, g;
  act_fwd_grad<ACT>(z, a, g);
  return a;
}
// run fn(std::integral_constant<int, ACT>) for the run-time activation code (warp-uniform)
#define FFB_ACT_DISPATCH(code, CALL)                                    \
  switch (code) {                                                       \
    case FFB_ACT_TANH: { constexpr int ACT = FFB_ACT_TANH; CALL; break; }         \
    case FFB_ACT_RELU: { constexpr int ACT = FFB_ACT_RELU; CALL; break; }         \
    case FFB_ACT_SOFTPLUS: { constexpr int ACT = FFB_ACT_SOFTPLUS; CALL; break; } \
    case FFB_ACT_GELU: { constexpr int ACT = FFB_ACT_GELU; CALL; break; }         \
    default: { constexpr int ACT = FFB_ACT_SILU; CALL; break; }                   \
  }

// ---------------------------------------------------------------------------------------------
// per-CTA context
// ---------------------------------------------------------------------------------------------
// Shared-memory buffers are kept as byte offsets into the dynamic shared block and turned into
// pointers at the point of use, so the compiler keeps the .shared state space (LDS/STS, not
// generic LD/ST) even though Ctx itself lives in local memory across non-inlined calls.
__device__ __forceinline__ unsigned char* smem_base() {
  extern __shared__ __align__(128) unsigned char ffb_smem_raw[];
  return ffb_smem_raw;
}
struct Ctx {
  // shared memory (byte offsets)
  uint32_t o_act;   // [KMAX][LDA]   layer activations, k-major
  uint32_t o_ring;  // [NSTAGE][CHUNK_FLOATS]
  uint32_t o_ycur;  // [SD][LDA]     current stage input (primal rows)
  uint32_t o_cond;  // [CD][LDA]
  uint32_t o_prb;   // [SD][LDA]     Hutchinson probes (or unused)
  uint32_t o_gate;  // [KMAX][GS]    silu'(z) of primal rows
  uint32_t o_beff;  // [KMAX]        layer-0 bias + time-feature contribution (packed order)
  uint32_t o_klp;   // [NSLOT+2][TM] d(logp)/dt per slot; row NSLOT = lp0, NSLOT+1 = lp_cur
  uint32_t o_red;   // [8][FFB_NPART] block-reduction scratch (double)
  uint32_t o_bar;   // full[NSTAGE], empty[NSTAGE]
  uint32_t o_slots; // [NSLOT][SD][LDA] state slots (when they fit)
  int slots_smem;
  __device__ __forceinline__ float* act() const { return reinterpret_cast<float*>(smem_base() + o_act); }
  __device__ __forceinline__ float* stage_buf() const { return act(); }   // free between evaluations
  __device__ __forceinline__ float* ring() const { return reinterpret_cast<float*>(smem_base() + o_ring); }
  __device__ __forceinline__ float* ycur() const { return reinterpret_cast<float*>(smem_base() + o_ycur); }
  __device__ __forceinline__ float* condb() const { return reinterpret_cast<float*>(smem_base() + o_cond); }
  __device__ __forceinline__ float* prb() const { return reinterpret_cast<float*>(smem_base() + o_prb); }
  __device__ __forceinline__ float* gate() const { return reinterpret_cast<float*>(smem_base() + o_gate); }
  __device__ __forceinline__ float* beff() const { return reinterpret_cast<float*>(smem_base() + o_beff); }
  __device__ __forceinline__ float* klp() const { return reinterpret_cast<float*>(smem_base() + o_klp); }
  __device__ __forceinline__ double* red() const { return reinterpret_cast<double*>(smem_base() + o_red); }
  __device__ __forceinline__ uint64_t* full() const { return reinterpret_cast<uint64_t*>(smem_base() + o_bar); }
  __device__ __forceinline__ uint64_t* empty() const { return full() + NSTAGE; }
  // global scratch of this CTA: [NSLOT][SD][LDA]
  float* scr;
  // tile geometry
  int S;           // trajectories per tile
  int T;           // tangents per trajectory
  int GS;          // gate row stride
  int SD, CD;
  // thread coordinates
  int tid, lane, warp, tx, ty;
  bool producer;
  // pipeline state
  int stage;
  uint32_t phase;
};

__device__ __forceinline__ float* slot_ptr(const Ctx& cx, int slot) {
  float* base = cx.slots_smem ? reinterpret_cast<float*>(smem_base() + cx.o_slots) : cx.scr;
  return base + (size_t)slot * cx.SD * LDA;
}

template <bool SS>
__device__ __forceinline__ float* slot_ptr_t(const Ctx& cx, int slot) {
  if (SS) return reinterpret_cast<float*>(smem_base() + cx.o_slots) + (size_t)slot * cx.SD * LDA;
  return cx.scr + (size_t)slot * cx.SD * LDA;
}

__device__ __forceinline__ void pipe_advance(Ctx& cx) {
  if (++cx.stage == NSTAGE) { cx.stage = 0; cx.phase ^= 1u; }
}

// size of the dynamic shared memory block for a field (host + device agree through this)
__host__ __device__ inline size_t smem_layout(int SD, int CD, int T, int hutch, int slots_smem, size_t* off /*[12]*/) {
  int S = TM / (1 + T);
  int GS = (T > 0) ? (S | 1) : 0;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  size_t v[12];
  v[0] = take(sizeof(float) * KMAX * LDA);                 // act
  v[1] = take(sizeof(float) * NSTAGE * CHUNK_FLOATS);      // ring
  v[2] = take(sizeof(float) * SD * LDA);                   // ycur
  v[3] = take(sizeof(float) * (CD > 0 ? CD : 1) * LDA);    // cond
  v[4] = take(hutch ? sizeof(float) * SD * LDA : 0);       // probes
  v[5] = take(sizeof(float) * KMAX * (GS > 0 ? GS : 1));   // gate
  v[6] = take(sizeof(float) * KMAX);                       // beff
  v[7] = take(sizeof(float) * (NSLOT + 2) * TM);           // klp
  v[8] = take(sizeof(double) * 8 * FFB_NPART);             // red
  v[9] = take(sizeof(uint64_t) * 2 * NSTAGE);              // barriers
  v[10] = take(slots_smem ? sizeof(float) * NSLOT * SD * LDA : 0);   // state slots
  if (off) for (int i = 0; i < 11; ++i) off[i] = v[i];
  return o;
}

__device__ __forceinline__ void ctx_init(Ctx& cx, const FieldDev& f, float* scratch) {
  int T = (f.div_mode == FFB_DIV_EXACT) ? f.net[0].x_dim : (f.div_mode == FFB_DIV_HUTCH ? 1 : 0);
  size_t off[12];
  smem_layout(f.state_dim, f.cond_dim, T, f.div_mode == FFB_DIV_HUTCH, f.slots_smem, off);
  cx.o_act = (uint32_t)off[0];
  cx.o_ring = (uint32_t)off[1];
  cx.o_ycur = (uint32_t)off[2];
  cx.o_cond = (uint32_t)off[3];
  cx.o_prb = (uint32_t)off[4];
  cx.o_gate = (uint32_t)off[5];
  cx.o_beff = (uint32_t)off[6];
  cx.o_klp = (uint32_t)off[7];
  cx.o_red = (uint32_t)off[8];
  cx.o_bar = (uint32_t)off[9];
  cx.o_slots = (uint32_t)off[10];
  cx.slots_smem = f.slots_smem;
  cx.T = T;
  cx.S = TM / (1 + T);
  cx.GS = (T > 0) ? (cx.S | 1) : 0;
  cx.SD = f.state_dim;
  cx.CD = f.cond_dim;
  cx.tid = threadIdx.x;
  cx.lane = threadIdx.x & 31;
  cx.warp = threadIdx.x >> 5;
  cx.producer = (cx.warp == NCOMP / 32);
  // warp w covers 4 row-groups x 8 col-groups of the 16 x 16 thread grid
  cx.ty = ((cx.warp >> 1) << 2) + (cx.lane >> 3);
  cx.tx = ((cx.warp & 1) << 3) + (cx.lane & 7);
  cx.scr = scratch + (size_t)blockIdx.x * NSLOT * f.state_dim * LDA;
  cx.stage = 0;
  cx.phase = cx.producer ? 1u : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&cx.full()[s], 1); mbar_init(&cx.empty()[s], NCOMP / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// producer: stream one network's packed weights through the ring
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void produce_net(Ctx& cx, const NetDev& net) {
  for (int l = 0; l < net.n_layers; ++l) {
    const int K = net.K[l], Np = net.Np[l];
    for (int k0 = 0; k0 < K; k0 += KC) {
      const int rows = min(KC, K - k0);
      if (cx.lane == 0) {
        mbar_wait(&cx.empty()[cx.stage], cx.phase);
        const uint32_t bytes = (uint32_t)(rows * Np) * sizeof(float);
        mbar_expect_tx(&cx.full()[cx.stage], bytes);
        bulk_g2s(cx.ring() + cx.stage * CHUNK_FLOATS, net.W[l] + (size_t)k0 * Np, bytes, &cx.full()[cx.stage]);
      }
      pipe_advance(cx);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// consumer: one layer's contraction for this thread's 8 rows x C columns
// ---------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void load_b(float (&b)[C], const float* p) {
  if constexpr (C == 8) {
    float4 b0 = *reinterpret_cast<const float4*>(p);
    float4 b1 = *reinterpret_cast<const float4*>(p + 64);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  } else if constexpr (C == 4) {
    float4 b0 = *reinterpret_cast<const float4*>(p);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
  } else if constexpr (C == 2) {
    float2 b0 = *reinterpret_cast<const float2*>(p);
    b[0] = b0.x; b[1] = b0.y;
  } else {
    b[0] = *p;
  }
}
// packed column offset of this thread, and the real column of its j-th packed column
template <int C>
__device__ __forceinline__ int col_base(int tx) { return (C == 8) ? tx * 4 : tx * C; }
template <int C>
__device__ __forceinline__ int col_real(int tx, int j) { return j * 16 + tx; }   // same formula for every C
// thread row i (0..7) -> tile row
__device__ __forceinline__ int row_of(int ty, int i) { return (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)); }

template <int C>
__device__ __forceinline__ void gemm_layer(Ctx& cx, const NetDev& net, int l, float2 (&acc)[4][C]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) acc[i][j] = make_float2(0.f, 0.f);
  const int K = net.K[l], Np = net.Np[l];
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int rows = min(KC, K - k0);
    mbar_wait(&cx.full()[cx.stage], cx.phase);
    const float* ap = cx.act() + (size_t)k0 * LDA + cx.ty * 4;
    const float* bp = cx.ring() + cx.stage * CHUNK_FLOATS + col_base<C>(cx.tx);
#pragma unroll 4
    for (int kk = 0; kk < rows; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(ap + kk * LDA);
      const float4 a1 = *reinterpret_cast<const float4*>(ap + kk * LDA + 64);
      float b[C];
      load_b<C>(b, bp + kk * Np);
      const float2 a[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y),
                           make_float2(a1.z, a1.w)};
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float2 bb = make_float2(b[j], b[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = __ffma2_rn(a[i], bb, acc[i][j]);
      }
    }
    __syncwarp();
    if (cx.lane == 0) mbar_arrive(&cx.empty()[cx.stage]);
    pipe_advance(cx);
  }
}

// write this thread's outputs for column (real index n) back into act, transposed
__device__ __forceinline__ void store_col(Ctx& cx, int n, float v0, float v1, float v2, float v3, float v4, float v5,
                                          float v6, float v7) {
  float* p = cx.act() + (size_t)n * LDA + cx.ty * 4;
  *reinterpret_cast<float4*>(p) = make_float4(v0, v1, v2, v3);
  *reinterpret_cast<float4*>(p + 64) = make_float4(v4, v5, v6, v7);
}

// One layer: contraction + epilogue.  `bias` is in packed column order.
//   hidden layers: primal rows  a = silu(z + b),  gate = silu'(z + b)
//                  tangent rows a = z * gate[primal row of the same trajectory]
//   last layer:    primal rows  z + b, tangent rows z   (raw, no activation)
template <int C>
__device__ __forceinline__ void run_layer(Ctx& cx, const NetDev& net, int l, const float* bias, bool last) {
  float2 acc[4][C];
  gemm_layer<C>(cx, net, l, acc);
  float v[8][C];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) { v[2 * i][j] = acc[i][j].x; v[2 * i + 1][j] = acc[i][j].y; }
  float bj[C];
  load_b<C>(bj, bias + col_base<C>(cx.tx));
  const int S = cx.S;
  const int live = S * (1 + cx.T);
  if (cx.T == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) {
        float z = v[i][j] + bj[j];
        v[i][j] = last ? z : z * sigmoidf_fast(z);
      }
    bar_compute();   // every warp has finished reading act
  } else {
    int rs[8];       // trajectory index of tangent rows, -1 for primal rows, -2 for dead rows
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = row_of(cx.ty, i);
      rs[i] = (r < S) ? -1 : (r < live ? (r % S) : -2);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (rs[i] != -1) continue;
      const int r = row_of(cx.ty, i);
#pragma unroll
      for (int j = 0; j < C; ++j) {
        float z = v[i][j] + bj[j];
        if (last) {
          v[i][j] = z;
        } else {
          float sg = sigmoidf_fast(z);
          v[i][j] = z * sg;
          cx.gate()[col_real<C>(cx.tx, j) * cx.GS + r] = sg * (1.0f + z * (1.0f - sg));
        }
      }
    }
    bar_compute();   // act fully read, gates visible
    if (!last) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (rs[i] == -1) continue;
#pragma unroll
        for (int j = 0; j < C; ++j)
          v[i][j] = (rs[i] >= 0) ? v[i][j] * cx.gate()[col_real<C>(cx.tx, j) * cx.GS + rs[i]] : 0.0f;
      }
    }
  }
  const int Nreal = net.N[l];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int n = col_real<C>(cx.tx, j);
    if (!last || n < Nreal)
      store_col(cx, n, v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
  }
  bar_compute();     // outputs visible before the next layer reads them
}

// ---------------------------------------------------------------------------------------------
// one network forward on the tile (consumer side).  Inputs must already be in cx.act().
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void consume_net(Ctx& cx, const NetDev& net) {
  for (int l = 0; l < net.n_layers; ++l) {
    const bool last = (l == net.n_layers - 1);
    const float* bias = (l == 0) ? cx.beff() : net.b[l];
    switch (net.Np[l]) {
      case 128: run_layer<8>(cx, net, l, bias, last); break;
      case 64:  run_layer<4>(cx, net, l, bias, last); break;
      case 32:  run_layer<2>(cx, net, l, bias, last); break;
      default:  run_layer<1>(cx, net, l, bias, last); break;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// evaluate the vector field at the state held in cx.ycur(); derivative -> scratch slot `dst`,
// divergence (if any) -> klp[dst].  Called by ALL threads (the producer warp streams weights).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void eval_field(Ctx& cx, const FieldDev& f, const ffb_eval_scalars& ev, int dst,
                                           unsigned call_mask = 3u) {
  for (int c = 0; c < f.n_calls; ++c) {
    if (!((call_mask >> c) & 1u)) continue;
    const NetDev& net = f.net[c];
    if (cx.producer) { produce_net(cx, net); continue; }
    const int S = cx.S, T = cx.T, live = S * (1 + T);
    const int Np0 = net.Np[0], xd = net.x_dim, cd = net.c_dim, K0 = net.K[0];
    // layer-0 bias with the (row-uniform) time features folded in
    for (int p = cx.tid; p < Np0; p += NCOMP) {
      float b = net.b[0][p];
      for (int j = 0; j < net.t_dim; ++j) b = fmaf(net.Wt[(size_t)j * Np0 + p], ev.tfeat[j], b);
      cx.beff()[p] = b;
    }
    // layer-0 input rows: [x | cond | zero pad]
    for (int idx = cx.tid; idx < K0 * TM; idx += NCOMP) {
      const int k = idx / TM, r = idx - k * TM;
      float val = 0.0f;
      if (r < S) {
        if (k < xd) val = cx.ycur()[(f.in_off[c] + k) * LDA + r];
        else if (k < xd + cd) val = cx.condb()[(k - xd) * LDA + r];
      } else if (r < live && k < xd) {
        const int j = r / S - 1, s = r - (j + 1) * S;
        val = (f.div_mode == FFB_DIV_EXACT) ? (k == j ? 1.0f : 0.0f) : cx.prb()[k * LDA + s];
      }
      cx.act()[k * LDA + r] = val;
    }
    bar_compute();
    consume_net(cx, net);
    // field transform on the primal rows; act[n][r] holds the raw network output
    const int Dout = net.N[net.n_layers - 1];
    float* kd = slot_ptr(cx, dst);
    for (int idx = cx.tid; idx < Dout * S; idx += NCOMP) {
      const int d = idx / S, r = idx - d * S;
      const float o = cx.act()[d * LDA + r];
      float xd_;
      if (f.kind == FFB_FIELD_SCORE) {
        const float sc = f.use_sigma ? __fdiv_rn(o, ev.sigma) : o;
        const float lin = f.has_drift ? __fmul_rn(ev.a, cx.ycur()[(f.out_off[c] + d) * LDA + r]) : 0.0f;
        xd_ = __fsub_rn(lin, __fmul_rn(ev.c, sc));
      } else {
        xd_ = o;
      }
      kd[(f.out_off[c] + d) * LDA + r] = xd_ * (ev.sign * f.out_sign[c]);
    }
    if (T > 0) {
      for (int s = cx.tid; s < S; s += NCOMP) {
        float tr = 0.0f;
        if (f.div_mode == FFB_DIV_EXACT) {
          for (int j = 0; j < T; ++j) tr += cx.act()[j * LDA + (j + 1) * S + s];
        } else {
          for (int d = 0; d < xd; ++d) tr = fmaf(cx.prb()[d * LDA + s], cx.act()[d * LDA + S + s], tr);
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
    }
    bar_compute();
  }
}

struct EngineFFMA {
  using Ctx = ffb::Ctx;
  static constexpr int NTHR = ffb::NTHR;
  static __device__ __forceinline__ void init(Ctx& cx, const FieldDev& f, float* scratch) { ctx_init(cx, f, scratch); }
  static __device__ __forceinline__ void fini(Ctx&) {}
  template <bool SS>
  static __device__ __forceinline__ void eval(Ctx& cx, const FieldDev& f, const ffb_eval_scalars& ev, int dst,
                                              unsigned call_mask = 3u) {
    eval_field(cx, f, ev, dst, call_mask);
  }
};

}  // namespace ffb

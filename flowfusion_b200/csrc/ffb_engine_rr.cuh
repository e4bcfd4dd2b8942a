// ffb_engine_rr.cuh -- "row-resident" tensor-core tile engine for fields WITHOUT tangent rows
// (sampling paths: PF-ODE, reverse SDE, flow sampling, symplectic).  Same arithmetic as
// ffb_engine_tc.cuh (tcgen05 kind::tf32, 3 products per k-step, FP32 accumulate in tensor memory),
// different schedule:
//
//   * 16 epilogue warps: warp w owns TMEM lane quarter q = w & 3 (rows 32q .. 32q+31) and column
//     group cg = w >> 2: columns [32c + 8cg, 32c + 8cg + 8) of every 32-column chunk c.  A thread is
//     one row x 8 columns of a chunk, so every chunk of a layer's output is finished by all 16 warps
//     together and chunks complete one after the other (4 warps per scheduler hide the SFU / TMEM
//     latencies of each other);
//   * chunk hand-off: as soon as chunk c of the next A operand (A_hi | A_lo, written with tcgen05.st)
//     is in tensor memory the MMA warp issues that chunk's 12 MMAs of the NEXT layer into the other
//     accumulator buffer (D0 | D1 | A_hi | A_lo = 4 x 128 TMEM columns), so the tensor core trails
//     the epilogue by one chunk instead of waiting for the whole layer;
//   * everything between two network evaluations is row-local: the thread that reads column n of
//     the last layer's accumulator applies the field transform, writes the stage derivative into
//     its own element of the state slot, runs the integrator's stage algebra for the state columns
//     it owns ((d >> 3) & 3 == cg) and splits the next layer-0 operand straight back into tensor
//     memory.  The only synchronisation is a 128-thread named barrier between the 4 warps that
//     share a lane quarter -- no CTA-wide barrier inside a trajectory;
//   * the layer-0 bias with the time features folded in (one vector per evaluation) is prepared
//     ahead of the evaluation that uses it.
//
// Accumulation order: per 32-row K chunk the two cross products (A_hi W_lo, A_lo W_hi) are issued
// before the main product (A_hi W_hi).  csrc/tc_probe.cu measures the orders: cross-first over the
// whole K 3.8e-7 rms, per-k-step interleave 1.1e-6, FP32 FMA chain 2.3e-7 (relative to rms |out|).
#pragma once
#include "ffb_engine_tc.cuh"

namespace ffb {

constexpr int RR_NCOMP = 512;
constexpr int RR_CW0 = 0;                      // first compute warp
constexpr int RR_NTHR = RR_NCOMP + 64;
constexpr int RR_WLOAD = RR_NCOMP / 32;
constexpr int RR_WMMA = RR_NCOMP / 32 + 1;
constexpr int RR_NCHUNK = KMAX / KC;
constexpr uint32_t RR_COL_AHI = 256, RR_COL_ALO = 384;     // D0 = 0, D1 = 128
// Synchronisation of one ring stage (= one 32-row K chunk of one layer): ONE mbarrier `full[s]` collects
// both conditions the MMA warp needs -- the weight chunk has landed (loader: arrive.expect_tx + the copy's
// complete_tx) and the matching 32 columns of the A operand are in tensor memory (one arrive per compute
// warp) -- because every mbarrier poll costs the MMA warp ~130 cycles while the epilogue warps keep the
// shared-memory pipe busy (measured with the -DFFB_TRACE timeline), and the tensor pipe's queue only
// hides ~250 cycles.  Compute warps and MMA warp walk the ring in the same order (layer by layer, chunk
// by chunk), so both sides know the stage of every chunk without exchanging it.

__device__ __forceinline__ void tc_ld8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr));
}

// Ring-stage release: the stages of a layer are released by compute thread 0 when it sees the layer's d_ready (a
// tcgen05.commit per stage was measured slower).
// TF32 hi/lo split of a packed pair on the FMA pipe (Veltkamp): c = 8193 a, hi = c - 8192 a (one FMA: exact, a multiple of
// 2^13 ulp(a), i.e. a rounded to TF32's 11 significant bits), lo = a - hi (exact): 3 packed instructions per pair.
// |lo| <= 2^-11 |a|; differs from the integer round-half-away of tf32_split on ties only.  |a| must stay below
// FLT_MAX / 8193 (4e34).
__device__ __forceinline__ void tf32_split2(const float2 a, uint32_t& hi0, uint32_t& hi1, uint32_t& lo0, uint32_t& lo1) {
  const float2 c = __fmul2_rn(a, make_float2(8193.0f, 8193.0f));
  const float2 h = __ffma2_rn(a, make_float2(-8192.0f, -8192.0f), c);
  const float2 l = __ffma2_rn(h, make_float2(-1.0f, -1.0f), a);
  hi0 = __float_as_uint(h.x); hi1 = __float_as_uint(h.y);
  lo0 = __float_as_uint(l.x); lo1 = __float_as_uint(l.y);
}

// Debug timeline (compiled in with -DFFB_TRACE only): CTA 0 records clock64() at hand-off points, one
// private region per role (0: compute warp 0, 1: MMA warp, 2: compute warp 15), no atomics.
#ifdef FFB_TRACE
#define RR_TRACE(cx, tag) rr_trace(cx, tag)
#else
#define RR_TRACE(cx, tag) do {} while (0)
#endif
constexpr int RR_TRACE_CAP = 2048;

// per-evaluation scalars by value (launch arguments) or read from the device-resident controller block when used
struct EvVals {
  float a_, c_, sigma_, sign_;
  __device__ __forceinline__ float a() const { return a_; }
  __device__ __forceinline__ float c() const { return c_; }
  __device__ __forceinline__ float sigma() const { return sigma_; }
  __device__ __forceinline__ float sign() const { return sign_; }
};
struct EvCtl {
  const ffb_dopri5_ctl* ctl; int i;
  __device__ __forceinline__ float a() const { return __ldg(&ctl->ev[i].a); }
  __device__ __forceinline__ float c() const { return __ldg(&ctl->ev[i].c); }
  __device__ __forceinline__ float sigma() const { return __ldg(&ctl->ev[i].sigma); }
  __device__ __forceinline__ float sign() const { return __ldg(&ctl->ev[i].sign); }
};

struct CtxR {
  int tr_role, tr_n;
  uint32_t o_ring, o_ycur, o_cond, o_sbias, o_beff, o_wt, o_red, o_bar, o_slots;
  __device__ __forceinline__ float* ring() const { return reinterpret_cast<float*>(smem_base() + o_ring); }
  __device__ __forceinline__ float* ycur() const { return reinterpret_cast<float*>(smem_base() + o_ycur); }
  __device__ __forceinline__ float* condb() const { return reinterpret_cast<float*>(smem_base() + o_cond); }
  __device__ __forceinline__ float* sbias() const { return reinterpret_cast<float*>(smem_base() + o_sbias); }
  __device__ __forceinline__ float* beff() const { return reinterpret_cast<float*>(smem_base() + o_beff); }
  __device__ __forceinline__ float* swt() const { return reinterpret_cast<float*>(smem_base() + o_wt); }
  __device__ __forceinline__ double* red() const { return reinterpret_cast<double*>(smem_base() + o_red); }
  __device__ __forceinline__ uint64_t* full() const { return reinterpret_cast<uint64_t*>(smem_base() + o_bar); }
  __device__ __forceinline__ uint64_t* empty() const { return full() + TC_NSTAGE; }
  __device__ __forceinline__ uint64_t* d_ready() const { return full() + 2 * TC_NSTAGE; }
  __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(full() + 2 * TC_NSTAGE + 1); }
  float* scr;
  int SD, CD, tdim, ncalls;
  int tid, lane, warp, cwarp;
  int q, cg, row;       // lane quarter, column group, tile row (= TMEM lane) of a compute thread
  bool producer;
  uint32_t tmem, lane_addr;
  int stage; uint32_t phase;       // ring position: loader / MMA warp (with phase); compute warps (stage of the next chunk they hand over)
  uint32_t ph_d, dbuf;             // d_ready parity (compute warps), accumulator buffer of the current layer
  int rstage;                      // compute warps: ring stage of the first chunk of the layer whose d_ready is awaited next
};

__device__ __forceinline__ void rr_trace(CtxR& cx, int tag) {
  if (cx.tr_role >= 0 && g_trace && cx.tr_n < RR_TRACE_CAP) {
    long long* p = g_trace + 2 * (cx.tr_role * RR_TRACE_CAP + cx.tr_n);
    p[0] = clock64(); p[1] = tag;
    ++cx.tr_n;
  }
}

template <bool SS>
__device__ __forceinline__ float* rr_slot(const CtxR& cx, int slot) {
  if (SS) return reinterpret_cast<float*>(smem_base() + cx.o_slots) + (size_t)slot * cx.SD * LDA;
  return cx.scr + (size_t)slot * cx.SD * LDA;
}

// the 512 compute threads / the 4 warps of one lane quarter
__device__ __forceinline__ void rr_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void rr_qbar(const CtxR& cx) { asm volatile("bar.sync %0, 128;" ::"r"(2 + cx.q) : "memory"); }

__host__ __device__ inline size_t smem_layout_rr(int SD, int CD, int nslot_smem, int ncalls, int tdim, int nbeff,
                                                 size_t* off /*[12]*/) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  size_t v[12];
  v[0] = take(sizeof(float) * TC_NSTAGE * TC_STAGE_FLOATS);                 // weight ring
  v[1] = take(sizeof(float) * SD * LDA);                                    // ycur
  v[2] = take(sizeof(float) * (CD > 0 ? CD : 1) * LDA);                     // cond
  v[3] = take(sizeof(float) * ncalls * NET_MAXL * KMAX);              // biases
  v[4] = take(sizeof(float) * nbeff * ncalls * KMAX);                       // layer-0 bias + time features, per evaluation
  v[5] = take(sizeof(float) * ncalls * (tdim > 0 ? tdim : 1) * KMAX);       // layer-0 time-feature rows
  v[6] = take(sizeof(double) * (RR_NCOMP / 32) * FFB_NPART);                // block-reduction scratch
  v[7] = take(sizeof(uint64_t) * (2 * TC_NSTAGE + RR_NCHUNK + 4));          // mbarriers + TMEM base slot
  v[8] = take(sizeof(float) * (size_t)nslot_smem * SD * LDA);               // state slots (0 when they live in global scratch)
  if (off) for (int i = 0; i < 9; ++i) off[i] = v[i];
  return o;
}

// GEN = false: SiLU networks (the hot instantiation, no activation dispatch in the epilogue);
// GEN = true: the activation code of the network is dispatched at run time (FFB_ACT_*).
template <bool GEN>
struct EngineRR_ {
  static constexpr int NTHR = RR_NTHR;

  static __device__ __forceinline__ void init(CtxR& cx, const FieldDev& f, float* scratch, int nslot, int nbeff) {
    size_t off[12];
    smem_layout_rr(f.state_dim, f.cond_dim, f.slots_smem ? nslot : 0, f.n_calls, field_tdim(f), nbeff, off);
    init_at(cx, f, scratch, off);
  }
  // off[0..8]: ring, ycur, cond, biases, beff, time rows, reduction scratch, barriers, slots
  static __device__ __forceinline__ void init_at(CtxR& cx, const FieldDev& f, float* scratch, const size_t* off) {
    cx.tdim = field_tdim(f);
    cx.o_ring = (uint32_t)off[0]; cx.o_ycur = (uint32_t)off[1]; cx.o_cond = (uint32_t)off[2];
    cx.o_sbias = (uint32_t)off[3]; cx.o_beff = (uint32_t)off[4]; cx.o_wt = (uint32_t)off[5];
    cx.o_red = (uint32_t)off[6]; cx.o_bar = (uint32_t)off[7]; cx.o_slots = (uint32_t)off[8];
    cx.SD = f.state_dim; cx.CD = f.cond_dim; cx.ncalls = f.n_calls;
    cx.tid = (int)threadIdx.x - 32 * RR_CW0; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
    cx.cwarp = cx.warp - RR_CW0;                    // compute-warp index 0..15 (meaningless on the special warps)
    cx.producer = cx.cwarp < 0 || cx.cwarp >= RR_NCOMP / 32;
    cx.q = cx.warp & 3; cx.cg = (cx.cwarp >> 2) & 3;
    cx.row = (cx.q << 5) + cx.lane;
    cx.scr = scratch + (size_t)blockIdx.x * NSLOT * f.state_dim * LDA;
    cx.stage = 0;
    cx.phase = (cx.warp == RR_WLOAD) ? 1u : 0u;     // the loader starts with every stage free
    cx.ph_d = 0; cx.dbuf = 0; cx.rstage = 0;
    cx.tr_n = 0;
    cx.tr_role = (blockIdx.x != 0 || cx.lane != 0) ? -1 : (cx.warp == RR_WMMA ? 1 : (cx.cwarp == 0 ? 0 : (cx.cwarp == RR_NCOMP / 32 - 1 ? 2 : -1)));
    if (threadIdx.x == 0) {
      for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(&cx.full()[s], 1 + RR_NCOMP / 32); mbar_init(&cx.empty()[s], 1); }   // loader + every compute warp
      mbar_init(cx.d_ready(), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (cx.warp == RR_WMMA) {    // the MMA warp owns the tensor memory: all 512 columns
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(cx.tmem_slot())), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int c = 0; c < f.n_calls; ++c) {
      for (int l = 0; l < f.net[c].n_layers; ++l)
        for (int n = threadIdx.x; n < KMAX; n += RR_NTHR)
          cx.sbias()[(c * NET_MAXL + l) * KMAX + n] = (n < f.net[c].Np[l]) ? f.net[c].b[l][n] : 0.0f;
      for (int i = threadIdx.x; i < f.net[c].t_dim * KMAX; i += RR_NTHR) {
        const int j = i / KMAX, n = i - j * KMAX;
        cx.swt()[(c * cx.tdim + j) * KMAX + n] = (n < f.net[c].Np[0]) ? f.net[c].Wt[(size_t)j * f.net[c].Np[0] + n] : 0.0f;
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cx.tmem = *cx.tmem_slot();
    cx.lane_addr = cx.tmem + ((uint32_t)(cx.q << 5) << 16);
  }

  static __device__ __forceinline__ void fini(CtxR& cx) {
    tc_fence_before();
    __syncthreads();
    if (cx.warp == RR_WMMA) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "r"(512));
    }
  }

  static __device__ __forceinline__ void advance(CtxR& cx) {
    if (++cx.stage == TC_NSTAGE) { cx.stage = 0; cx.phase ^= 1u; }
  }

  // layer-0 bias of every call of one evaluation with the (row-uniform) time features folded in
  static __device__ __forceinline__ void prep_beff(CtxR& cx, const FieldDev& f, const float* tfeat, float* buf) {
    for (int i = cx.tid; i < f.n_calls * KMAX; i += RR_NCOMP) {
      const int c = i / KMAX, n = i - c * KMAX;
      float b = cx.sbias()[(c * NET_MAXL) * KMAX + n];
      const float* wt = cx.swt() + (c * cx.tdim) * KMAX + n;
      for (int j = 0; j < f.net[c].t_dim; ++j) b = fmaf(wt[j * KMAX], tfeat[j], b);
      buf[i] = b;
    }
  }

  // ---- loader warp -----------------------------------------------------------------------------
  static __device__ __forceinline__ void load_net(CtxR& cx, const NetDev& net) {
    if (cx.lane != 0) return;
    for (int l = 0; l < net.n_layers; ++l) {
      const int K = net.K[l], Np = net.Np[l];
      for (int k0 = 0; k0 < K; k0 += KC) {
        const int rows = min(KC, K - k0);
        mbar_wait(&cx.empty()[cx.stage], cx.phase);
        const uint32_t bytes = (uint32_t)(2 * rows * Np) * sizeof(float);
        mbar_expect_tx(&cx.full()[cx.stage], bytes);
        bulk_g2s(cx.ring() + cx.stage * TC_STAGE_FLOATS, net.W[l] + (size_t)2 * k0 * Np, bytes, &cx.full()[cx.stage]);
        advance(cx);
      }
    }
  }

  // ---- MMA warp: chunk c of layer l as soon as its weights and its A columns are there ------------
  // NJ > 0: compile-time number of k-steps per chunk (4 for a full 32-row chunk) -> no per-MMA branches
  template <int NJ>
  static __device__ __forceinline__ void issue_chunk(uint32_t d_acc, uint32_t a_hi0, uint32_t a_lo0, uint64_t dh0, uint64_t dl0,
                                                     uint64_t kstep, uint32_t idesc, uint32_t acc0, int nj) {
#pragma unroll
    for (int j = 0; j < KC / 8; ++j) {
      if ((NJ > 0) ? (j < NJ) : (j < nj)) {
        tc_mma_ts(d_acc, a_hi0 + 8u * j, dl0 + (uint64_t)j * kstep, idesc, (j == 0) ? acc0 : 1u);
        tc_mma_ts(d_acc, a_lo0 + 8u * j, dh0 + (uint64_t)j * kstep, idesc, 1u);
      }
    }
#pragma unroll
    for (int j = 0; j < KC / 8; ++j)
      if ((NJ > 0) ? (j < NJ) : (j < nj)) tc_mma_ts(d_acc, a_hi0 + 8u * j, dh0 + (uint64_t)j * kstep, idesc, 1u);
  }
  static __device__ __forceinline__ void mma_net(CtxR& cx, const NetDev& net) {
    for (int l = 0; l < net.n_layers; ++l) {
      const int K = net.K[l], Np = net.Np[l];
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t lbo = (uint32_t)Np * 16u;
      const uint64_t kstep = (uint64_t)(lbo >> 3);
      const uint32_t d_acc = cx.tmem + cx.dbuf * 128u;
      cx.dbuf ^= 1u;
      uint32_t acc = 0;
      int ci = 0;
      for (int k0 = 0; k0 < K; k0 += KC, ++ci) {
        const int nj = min(KC, K - k0) >> 3;
        mbar_wait(&cx.full()[cx.stage], cx.phase);            // W chunk landed AND A columns [k0, k0+32) written
        tc_fence_after();
        RR_TRACE(cx, 100 + 10 * l + ci);
        const uint32_t hi_base = smem_u32(cx.ring() + cx.stage * TC_STAGE_FLOATS);
        const uint64_t dh0 = tc_desc(hi_base, lbo, 128u);
        const uint64_t dl0 = tc_desc(hi_base + (uint32_t)(nj * 8 * Np) * 4u, lbo, 128u);
        const uint32_t a_hi0 = cx.tmem + RR_COL_AHI + (uint32_t)k0, a_lo0 = cx.tmem + RR_COL_ALO + (uint32_t)k0;
        const bool lastc = (k0 + KC >= K);
        if (elect_one()) {
          if (nj == KC / 8) issue_chunk<KC / 8>(d_acc, a_hi0, a_lo0, dh0, dl0, kstep, idesc, acc, nj);
          else issue_chunk<0>(d_acc, a_hi0, a_lo0, dh0, dl0, kstep, idesc, acc, nj);
          if (lastc) tc_commit(cx.d_ready());                   // the accumulator of this layer is complete
        }
        __syncwarp();
        RR_TRACE(cx, 300 + ci);
        acc = 1u;
        advance(cx);
      }
      RR_TRACE(cx, 190 + l);
    }
  }

  // ---- compute warps ---------------------------------------------------------------------------
  // "my part of the next A chunk is in tensor memory": arrive on the ring stage that chunk will use
  static __device__ __forceinline__ void signal_chunk(CtxR& cx) {
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (cx.lane == 0) mbar_arrive(&cx.full()[cx.stage]);
    if (++cx.stage == TC_NSTAGE) cx.stage = 0;
  }
  // K = rows of the layer whose accumulator is awaited; its late ring stages are released here
  static __device__ __forceinline__ void wait_d_ready(CtxR& cx, int K) {
    mbar_wait(cx.d_ready(), cx.ph_d);
    cx.ph_d ^= 1u;
    tc_fence_after();
    for (int k0 = 0; k0 < K; k0 += KC) {
      if (cx.tid == 0) mbar_arrive(&cx.empty()[cx.rstage]);
      if (++cx.rstage == TC_NSTAGE) cx.rstage = 0;
    }
  }

  // layer-0 operand of call c from cx.ycur() / cx.condb(): this thread's row, its 8 columns per chunk.
  // Loads are unconditional (clamped index) so that all 8 are in flight together; a group of 8 columns that
  // holds no state column (conditional / zero padding only) skips the state loads.
  static __device__ __forceinline__ void build_A(CtxR& cx, const FieldDev& f, int c) {
    const NetDev& net = f.net[c];
    const int K0 = net.K[0], xd = net.x_dim, cd = net.c_dim;
    const float* yc = cx.ycur() + f.in_off[c] * LDA + cx.row;
    const float* cb = cx.condb() + cx.row;
    int ci = 0;
    for (int k0 = 0; k0 < K0; k0 += KC, ++ci) {
      const int k8 = k0 + 8 * cx.cg;
      if (k8 < K0) {                                             // warp-uniform
        uint32_t hi[8], lo[8];
        if (k8 + 8 <= xd) {                                      // state columns only (warp-uniform)
          float vx[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) vx[j] = yc[(k8 + j) * LDA];
#pragma unroll
          for (int j = 0; j < 8; ++j) tf32_split(vx[j], hi[j], lo[j]);
        } else if (k8 >= xd + cd) {                              // zero padding only
#pragma unroll
          for (int j = 0; j < 8; ++j) { hi[j] = 0u; lo[j] = 0u; }
        } else {
          float vx[8], vc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = k8 + j;
            vx[j] = (k8 < xd) ? yc[min(k, xd - 1) * LDA] : 0.0f;
            vc[j] = (cd > 0) ? cb[min(max(k - xd, 0), cd - 1) * LDA] : 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = k8 + j;
            const float val = (k < xd) ? vx[j] : ((k < xd + cd) ? vc[j] : 0.0f);
            tf32_split(val, hi[j], lo[j]);
          }
        }
        tc_st8(cx.lane_addr + RR_COL_AHI + k8, hi);
        tc_st8(cx.lane_addr + RR_COL_ALO + k8, lo);
      }
      signal_chunk(cx);
    }
  }

  // hidden layers: accumulator chunk -> + bias -> SiLU -> TF32 split -> next A operand chunk -> hand off
  static __device__ __forceinline__ void hidden(CtxR& cx, const NetDev& net, int c, const float* beff) {
    if (GEN) { FFB_ACT_DISPATCH(net.act, hidden_act<ACT>(cx, net, c, beff)); }
    else hidden_act<FFB_ACT_SILU>(cx, net, c, beff);
  }
  template <int ACT>
  static __device__ __forceinline__ void hidden_act(CtxR& cx, const NetDev& net, int c, const float* beff) {
    for (int l = 0; l + 1 < net.n_layers; ++l) {
      const int nc = net.Np[l] / KC;
      const float* bias = (l == 0) ? beff : cx.sbias() + (c * NET_MAXL + l) * KMAX;
      const uint32_t dcol = cx.lane_addr + cx.dbuf * 128u + 8u * cx.cg;
      cx.dbuf ^= 1u;
      wait_d_ready(cx, net.K[l]);
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
          if (ACT == FFB_ACT_SILU) {
            // SiLU + TF32 split on packed FP32 pairs (add/mul/fma.f32x2): 6.5 issue slots per element instead of 11,
            // which leaves the schedulers to the SFU (2 MUFU per element) and to the MMA warp
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
              const float2 z = __fadd2_rn(make_float2(__uint_as_float(m[ci & 1][u]), __uint_as_float(m[ci & 1][u + 1])),
                                          make_float2(bb[u], bb[u + 1]));
              const float2 x = __fmul2_rn(z, make_float2(-1.4426950408889634f, -1.4426950408889634f));
              float2 e, r;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(x.x));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(x.y));
              const float2 s = __fadd2_rn(e, make_float2(1.0f, 1.0f));
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(s.x));
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(s.y));
              const float2 a = __fmul2_rn(z, r);
              tf32_split2(a, hi[u], hi[u + 1], lo[u], lo[u + 1]);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float z = __uint_as_float(m[ci & 1][u]) + bb[u];
              tf32_split(act_fwd<ACT>(z), hi[u], lo[u]);
            }
          }
          tc_st8(cx.lane_addr + RR_COL_AHI + c0, hi);
          tc_st8(cx.lane_addr + RR_COL_ALO + c0, lo);
          signal_chunk(cx);
          RR_TRACE(cx, 201 + 10 * l + ci);
        }
      }
    }
  }

  // last layer: fn(c0, o[8]) with o[u] = raw output + bias of column c0 + u, for every 8-column group this
  // thread owns that holds a real column (columns >= N[last] of the group are padding)
  template <class F>
  static __device__ __forceinline__ void last(CtxR& cx, const NetDev& net, int c, const float* beff, F&& fn) {
    const int nl = net.n_layers, Nreal = net.N[nl - 1];
    const float* bias = (nl == 1) ? beff : cx.sbias() + (c * NET_MAXL + nl - 1) * KMAX;
    const uint32_t dcol = cx.lane_addr + cx.dbuf * 128u + 8u * cx.cg;
    cx.dbuf ^= 1u;
    wait_d_ready(cx, net.K[nl - 1]);
    RR_TRACE(cx, 290);
    for (int c0 = 8 * cx.cg; c0 < Nreal; c0 += KC) {              // warp-uniform trip count
      uint32_t m[8];
      tc_ld8(dcol + (uint32_t)(c0 - 8 * cx.cg), m);
      const float4 b0 = *reinterpret_cast<const float4*>(bias + c0);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + c0 + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      tc_wait_ld();
      float o[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = __uint_as_float(m[u]) + bb[u];
      fn(c0, o);
    }
    tc_fence_before();
  }

  // ---- one evaluation of the field at cx.ycur(): derivative -> slot dst -------------------------
  // Every warp of the CTA calls this; on return the compute warps of a lane quarter have passed a
  // quarter barrier, i.e. slot dst is complete for the rows of that quarter.  `overlap()` runs on the
  // compute warps right after the first layer-0 operand has been handed to the MMA warp (idle time).
  struct NoOverlap { __device__ __forceinline__ void operator()() const {} };
  template <bool SS, class OV = NoOverlap>
  static __device__ __forceinline__ void eval(CtxR& cx, const FieldDev& f, float ev_a, float ev_c, float ev_sigma,
                                              float ev_sign, const float* beff, int dst, unsigned call_mask = 3u,
                                              OV&& overlap = OV()) {
    eval_ev<SS>(cx, f, EvVals{ev_a, ev_c, ev_sigma, ev_sign}, beff, dst, call_mask, static_cast<OV&&>(overlap));
  }
  // EV: where the per-evaluation scalars come from.  They are fetched where the last-layer epilogue uses them, so a
  // source that costs a load (EvCtl: the device-resident controller block) holds no register across the hidden layers.
  template <bool SS, class EV, class OV = NoOverlap>
  static __device__ __forceinline__ void eval_ev(CtxR& cx, const FieldDev& f, const EV& ev, const float* beff, int dst,
                                                 unsigned call_mask = 3u, OV&& overlap = OV()) {
    bool first = true;
    for (int c = 0; c < f.n_calls; ++c) {
      if (!((call_mask >> c) & 1u)) continue;
      const NetDev& net = f.net[c];
      if (cx.warp == RR_WLOAD) { load_net(cx, net); continue; }
      if (cx.warp == RR_WMMA) { mma_net(cx, net); continue; }
      if (cx.producer) continue;             // spare warps
      rr_qbar(cx);                           // cx.ycur() of this lane quarter is final
      RR_TRACE(cx, 10);
      build_A(cx, f, c);
      RR_TRACE(cx, 11);
      if (first) { overlap(); first = false; }
      hidden(cx, net, c, beff + c * KMAX);
      float* kd = rr_slot<SS>(cx, dst) + cx.row;
      const float* yc = cx.ycur() + cx.row;
      const int ooff = f.out_off[c], Nreal = net.N[net.n_layers - 1];
      const bool score = (f.kind == FFB_FIELD_SCORE), use_sigma = f.use_sigma != 0, has_drift = f.has_drift != 0;
      last(cx, net, c, beff + c * KMAX, [&](int c0, const float (&o)[8]) {
        float yv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) yv[u] = (score && has_drift) ? yc[(ooff + min(c0 + u, Nreal - 1)) * LDA] : 0.0f;
        float xd_[8];
        const float sgn = ev.sign() * f.out_sign[c];
        const float ev_a = (score && has_drift) ? ev.a() : 0.0f, ev_c = score ? ev.c() : 0.0f;
        const float ev_sigma = (score && use_sigma) ? ev.sigma() : 1.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (score) {
            const float sc = use_sigma ? __fdiv_rn(o[u], ev_sigma) : o[u];
            const float lin = has_drift ? __fmul_rn(ev_a, yv[u]) : 0.0f;
            xd_[u] = __fsub_rn(lin, __fmul_rn(ev_c, sc)) * sgn;
          } else {
            xd_[u] = o[u] * sgn;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u < Nreal) kd[(ooff + c0 + u) * LDA] = xd_[u];
      });
    }
    RR_TRACE(cx, 12);
    if (!cx.producer) rr_qbar(cx);
    RR_TRACE(cx, 13);
  }
};
using EngineRR = EngineRR_<false>;

// ---- row-local stage algebra helpers --------------------------------------------------------------
// A thread owns, in its row, the state columns d with (d >> 3) & 3 == cg: blocks of 8 columns starting at
// d0 = 8 cg + 32 b.  rr_load8 reads a block of a [d][row] buffer with unconditional loads (columns past
// SD are clamped to SD - 1: valid memory, value unused) so the 8 loads are in flight together;
// rr_store8 writes the real columns only.
template <class F>
__device__ __forceinline__ void rr_for_blocks(const CtxR& cx, F&& fn) {
  for (int d0 = 8 * cx.cg; d0 < cx.SD; d0 += 32) fn(d0);
}
__device__ __forceinline__ void rr_load8(const CtxR& cx, const float* buf, int d0, float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = buf[min(d0 + u, cx.SD - 1) * LDA + cx.row];
}
__device__ __forceinline__ void rr_load8_if(const CtxR& cx, bool on, const float* buf, int d0, float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = on ? buf[min(d0 + u, cx.SD - 1) * LDA + cx.row] : 0.0f;
}
__device__ __forceinline__ void rr_store8(const CtxR& cx, float* buf, int d0, const float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (d0 + u < cx.SD) buf[(d0 + u) * LDA + cx.row] = v[u];
}

}  // namespace ffb

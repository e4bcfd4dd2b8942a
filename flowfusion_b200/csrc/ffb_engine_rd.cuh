// ffb_engine_rd.cuh -- "row-resident, dual-tile" tensor-core engine for fields WITHOUT tangent rows
// (sampling paths: PF-ODE, reverse SDE, flow sampling, symplectic).  Same arithmetic as ffb_engine_rr.cuh
// (tcgen05 kind::tf32, 3 products per k-step, cross products first, FP32 accumulate in tensor memory), but
// TWO independent 128-row tiles are resident on every SM so that one tile's MMAs fill the time the other
// tile spends in its epilogue (bias + activation + TF32 split) and in the row-local stage algebra:
//
//   * 16 compute warps = 2 groups x 8 warps.  Group g owns one tile at a time; inside the group warp w owns
//     TMEM lane quarter q = w & 3 (rows 32q .. 32q+31) and column parity cg = (w >> 2) & 1: the 8-column
//     blocks [16b + 8cg, 16b + 8cg + 8) of every layer output, every layer-0 operand and every state buffer.
//     A thread is one row x 8 columns of a block;
//   * tensor memory: group g holds one accumulator D (128 columns) and the high half A_hi of its A operand
//     (128 columns): D0 | A_hi0 | D1 | A_hi1 = all 512 columns.  The low half A_lo lives in SHARED memory
//     (64 KB per group, canonical no-swizzle K-major core-matrix image), so of the three products of a k-step
//     A_hi W_lo and A_hi W_hi are TS-form MMAs and A_lo W_hi is an SS-form MMA (same issue rate, measured
//     with csrc/tc_rate.cu);
//   * one loader warp and one MMA warp serve both groups, layer by layer in the fixed order
//     (layer l, group 0), (layer l, group 1), (layer l+1, group 0) ...: the weights of a layer are streamed
//     once per group through a 64 KB ring (cp.async.bulk, full/empty mbarriers, stages released by
//     tcgen05.commit), `a_ready[g]` (one arrive per compute warp of the group) says the group's A operand of the
//     layer is complete, `d_ready[g]` (tcgen05.commit) that its accumulator is;
//   * everything between two network evaluations is row-local exactly as in the single-tile engine; the only
//     synchronisation inside a trajectory is a 64-thread named barrier between the two warps that share a
//     lane quarter of a group.
//
// Shared memory: 2 x 64 KB (A_lo) + 64 KB (ring) leave ~35 KB, so the state slots of an integrator (and, for
// wide states, the current stage input and the conditional) live in a per-group global scratch that stays in
// L1/L2 (template parameter MEM: 0 = everything in shared memory, 1 = slots global, 2 = slots + stage input +
// conditional global).
#pragma once
#include "ffb_engine_rr.cuh"

namespace ffb {

constexpr int RD_NGROUP = 2;
constexpr int RD_GWARPS = 8;                          // compute warps per group
constexpr int RD_GTHR = RD_GWARPS * 32;               // compute threads per group
constexpr int RD_NCOMP = RD_NGROUP * RD_GTHR;         // 512
#ifndef RD_MMA_WARPS
#define RD_MMA_WARPS 2                                 // MMA-issuing warps: 1, or 2 that take the ring stages (chunks) in turn
#endif
constexpr int RD_NTHR = RD_NCOMP + 32 + 32 * RD_MMA_WARPS;   // + loader warp + MMA warp(s)
constexpr int RD_WLOAD = RD_NCOMP / 32;
constexpr int RD_WMMA = RD_NCOMP / 32 + 1;            // first MMA warp (owns the tensor-memory allocation)
#ifndef RD_KC
#define RD_KC 32                                       // weight rows (k) per ring stage: 32 or 16
#endif
#ifndef RD_NSTAGE
#define RD_NSTAGE (64 / RD_KC)                         // 64 KB ring
#endif
#ifndef RD_CLUSTER
#define RD_CLUSTER 2                                   // CTAs per cluster: every weight chunk is read from L2 once per cluster
#endif                                                 // (each CTA loads 1 / RD_CLUSTER of it and multicasts the piece to all)
#ifndef RD_TURN
#define RD_TURN (32 / RD_KC)                           // ring stages one MMA warp issues before it passes the turn (32 k-rows)
#endif
constexpr int RD_STAGE_FLOATS = 2 * RD_KC * KMAX;     // W_hi | W_lo rows of one stage
constexpr uint32_t RD_ALO_LBO = TM * 16u;             // bytes between two K core matrices of the A_lo image
// named barriers: 0 = __syncthreads, 1 + g = group g, 3 + 4 g + q = lane quarter q of group g, 11 = all compute warps
__device__ __forceinline__ void rd_allbar() { asm volatile("bar.sync 11, 512;" ::: "memory"); }

// D[tmem_d] (+)= A[smem desc] * B[smem desc]   (kind::tf32, cta_group::1, both operands from shared memory)
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (the tensor core reads A_lo through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- thread-block cluster helpers (RD_CLUSTER > 1) ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> shared bulk copy delivered to the same offset of every CTA in `mask`; each destination's mbarrier (same
// offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset of every CTA in `mask`
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct CtxD {
  uint32_t o_ring, o_alo, o_ycur, o_cond, o_sbias, o_beff, o_red, o_bar, o_slots;   // o_alo, o_ycur, o_cond, o_red, o_slots: this group's
  uint32_t alo_gstride;             // bytes between the A_lo images of the two groups (special warps address both)
  __device__ __forceinline__ float* ring() const { return reinterpret_cast<float*>(smem_base() + o_ring); }
  __device__ __forceinline__ float* sbias() const { return reinterpret_cast<float*>(smem_base() + o_sbias); }
  __device__ __forceinline__ float* beff() const { return reinterpret_cast<float*>(smem_base() + o_beff); }
  __device__ __forceinline__ double* red() const { return reinterpret_cast<double*>(smem_base() + o_red); }
  __device__ __forceinline__ uint64_t* full() const { return reinterpret_cast<uint64_t*>(smem_base() + o_bar); }
  __device__ __forceinline__ uint64_t* empty() const { return full() + RD_NSTAGE; }
  __device__ __forceinline__ uint64_t* a_ready() const { return full() + 2 * RD_NSTAGE; }
  __device__ __forceinline__ uint64_t* d_ready() const { return full() + 2 * RD_NSTAGE + RD_NGROUP; }
  __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(full() + 2 * RD_NSTAGE + 2 * RD_NGROUP); }
  float* scr;                       // this group's global scratch: [RD_SCR_SLOTS][SD][LDA] + [CD][LDA]
  int SD, CD, maxl, ncalls;
  int tid, lane, warp;              // tid: index inside the group (compute warps)
  int g, q, cg, row;                // group, lane quarter, column parity, tile row (= TMEM lane)
  bool producer;
  unsigned active;                  // special warps: groups that hold a tile in this round (bit g)
  uint32_t tmem, lane_addr;         // lane_addr: this thread's lane quarter, column 0 of its group
  uint32_t alo_row;                 // shared-space address of this thread's row in its group's A_lo image
  int stage; uint32_t phase;        // ring position (loader / MMA warp)
  uint32_t ph_d;                    // compute warps: parity of d_ready[g]
  uint32_t ph_a0, ph_a1;            // MMA warp: parity of a_ready[0], a_ready[1]
  int mw; uint32_t cctr, tctr;      // MMA warps: index of this MMA warp, ring stages (chunks) / turns visited so far
  uint32_t crank;                   // rank of this CTA in its cluster
  int tr_role, tr_n;                // debug timeline (-DFFB_TRACE)
};

// Debug timeline (compiled in with -DFFB_TRACE only): CTA 0 records clock64() at hand-off points, one private
// region per role (0: compute warp 0 of group 0, 1: MMA warp 0, 2: compute warp 0 of group 1, 3: loader, 4: MMA warp 1), no atomics.
#if defined(FFB_TRACE) && defined(FFB_TRACE_ROUNDS)      // only the per-tile markers (tags >= 900): a whole launch fits the buffer
#define RD_TRACE(cx, tag) do { if ((tag) >= 900) rd_trace(cx, tag); } while (0)
#elif defined(FFB_TRACE)
#define RD_TRACE(cx, tag) rd_trace(cx, tag)
#else
#define RD_TRACE(cx, tag) do {} while (0)
#endif
__device__ __forceinline__ void rd_trace(CtxD& cx, int tag) {
  if (cx.tr_role >= 0 && g_trace && cx.tr_n < RR_TRACE_CAP) {
    long long* p = g_trace + 2 * (cx.tr_role * RR_TRACE_CAP + cx.tr_n);
    long long t = clock64();
    if (tag >= 950) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));      // wall-clock nanoseconds
    p[0] = t; p[1] = tag;
    ++cx.tr_n;
  }
}

__device__ __forceinline__ void rd_gbar(const CtxD& cx) { asm volatile("bar.sync %0, 256;" ::"r"(1 + cx.g) : "memory"); }
__device__ __forceinline__ void rd_qbar(const CtxD& cx) { asm volatile("bar.sync %0, 64;" ::"r"(3 + 4 * cx.g + cx.q) : "memory"); }

template <int MEM>
__device__ __forceinline__ float* rd_slot(const CtxD& cx, int slot) {
  if (MEM == 0) return reinterpret_cast<float*>(smem_base() + cx.o_slots) + (size_t)slot * cx.SD * LDA;
  return cx.scr + (size_t)slot * cx.SD * LDA;
}
template <int MEM>
__device__ __forceinline__ float* rd_ycur(const CtxD& cx) {
  if (MEM < 2) return reinterpret_cast<float*>(smem_base() + cx.o_ycur);
  return cx.scr + (size_t)NSLOT * cx.SD * LDA;
}
// slots beyond the integrator's K / Y0 slots: NSLOT = stage input (global copy, MEM == 2), RD_XSLOT .. = partial sums
constexpr int RD_XSLOT = NSLOT + 1;
constexpr int RD_NXSLOT = 2;
constexpr int RD_SCR_SLOTS = RD_XSLOT + RD_NXSLOT;
template <int MEM>
__device__ __forceinline__ float* rd_cond(const CtxD& cx) {
  if (MEM < 2) return reinterpret_cast<float*>(smem_base() + cx.o_cond);
  return cx.scr + (size_t)RD_SCR_SLOTS * cx.SD * LDA;
}
// floats of one group's global scratch
__host__ __device__ inline size_t rd_scratch_floats(int SD, int CD) {
  return (size_t)RD_SCR_SLOTS * SD * LDA + (size_t)(CD > 0 ? CD : 1) * LDA;
}

// ka: widest A operand (max K over every layer of every call, a multiple of 8); maxl: most layers of a call;
// nbeff: layer-0 bias buffers (each ncalls x KMAX floats); mem: see MEM above.
// off[0..8]: ring, A_lo (x2), ycur (x2), cond (x2), biases, beff, reduction scratch (x2), barriers, slots (x2);
// gstride[i]: distance between the two groups' copies of region i
__host__ __device__ inline size_t smem_layout_rd(int SD, int CD, int ka, int maxl, int mem, int nslot, int ncalls, int nbeff,
                                                 size_t* off /*[12]*/, size_t* gstride /*[12]*/) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  auto take2 = [&](size_t bytes, size_t& stride) { stride = (bytes + 127) & ~size_t(127); size_t r = o; o += 2 * stride; return r; };
  size_t v[12], s[12];
  for (int i = 0; i < 12; ++i) { v[i] = 0; s[i] = 0; }
  v[0] = take(sizeof(float) * RD_NSTAGE * RD_STAGE_FLOATS);                       // weight ring
  v[1] = take2((size_t)(ka / 4) * RD_ALO_LBO, s[1]);                              // A_lo images
  v[2] = take2(mem < 2 ? sizeof(float) * SD * LDA : 0, s[2]);                     // ycur
  v[3] = take2(mem < 2 ? sizeof(float) * (CD > 0 ? CD : 1) * LDA : 0, s[3]);      // cond
  v[4] = take(sizeof(float) * ncalls * maxl * KMAX);                              // biases
  v[5] = take(sizeof(float) * nbeff * ncalls * KMAX);                             // layer-0 bias + time features, per evaluation
  v[6] = take2(sizeof(double) * RD_GWARPS * FFB_NPART, s[6]);                     // block-reduction scratch
  v[7] = take(sizeof(uint64_t) * (2 * RD_NSTAGE + 2 * RD_NGROUP + 2));            // mbarriers + TMEM base slot
  v[8] = take2(mem == 0 ? sizeof(float) * (size_t)nslot * SD * LDA : 0, s[8]);    // state slots
  if (off) for (int i = 0; i < 9; ++i) off[i] = v[i];
  if (gstride) for (int i = 0; i < 9; ++i) gstride[i] = s[i];
  return o;
}
__host__ __device__ inline int rd_field_ka(const FieldDev& f) {
  int ka = 8;
  for (int c = 0; c < f.n_calls; ++c)
    for (int l = 0; l < f.net[c].n_layers; ++l) ka = f.net[c].K[l] > ka ? f.net[c].K[l] : ka;
  return (ka + 7) & ~7;
}
__host__ __device__ inline int rd_field_maxl(const FieldDev& f) {
  int m = f.net[0].n_layers;
  if (f.n_calls > 1 && f.net[1].n_layers > m) m = f.net[1].n_layers;
  return m;
}

// GEN = false: SiLU networks (the hot instantiation); GEN = true: activation dispatched at run time (FFB_ACT_*).
// MEM: where the state slots / the stage input / the conditional live (see the header comment).
template <bool GEN, int MEM>
struct EngineRD_ {
  static constexpr int NTHR = RD_NTHR;

  static __device__ __forceinline__ void init(CtxD& cx, const FieldDev& f, float* scratch, int nslot, int nbeff) {
    size_t off[12], gs[12];
    cx.maxl = rd_field_maxl(f);
    smem_layout_rd(f.state_dim, f.cond_dim, rd_field_ka(f), cx.maxl, MEM, nslot, f.n_calls, nbeff, off, gs);
    cx.SD = f.state_dim; cx.CD = f.cond_dim; cx.ncalls = f.n_calls;
    cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
    cx.producer = cx.warp >= RD_NCOMP / 32;
    cx.g = cx.producer ? 0 : (cx.warp >> 3);
    cx.tid = (int)threadIdx.x - cx.g * RD_GTHR;
    cx.q = cx.warp & 3; cx.cg = (cx.warp >> 2) & 1;
    cx.row = (cx.q << 5) + cx.lane;
    cx.o_ring = (uint32_t)off[0];
    cx.alo_gstride = (uint32_t)gs[1];
    cx.o_alo = (uint32_t)(off[1] + cx.g * gs[1]);
    cx.o_ycur = (uint32_t)(off[2] + cx.g * gs[2]);
    cx.o_cond = (uint32_t)(off[3] + cx.g * gs[3]);
    cx.o_sbias = (uint32_t)off[4]; cx.o_beff = (uint32_t)off[5];
    cx.o_red = (uint32_t)(off[6] + cx.g * gs[6]);
    cx.o_bar = (uint32_t)off[7];
    cx.o_slots = (uint32_t)(off[8] + cx.g * gs[8]);
    cx.scr = scratch + ((size_t)blockIdx.x * RD_NGROUP + cx.g) * rd_scratch_floats(f.state_dim, f.cond_dim);
    cx.stage = 0;
    cx.phase = (cx.warp == RD_WLOAD) ? 1u : 0u;      // the loader starts with every stage free
    cx.ph_d = 0; cx.ph_a0 = 0; cx.ph_a1 = 0;
    cx.mw = cx.warp - RD_WMMA; cx.cctr = 0; cx.tctr = 0;
    cx.active = 0;
    cx.tr_n = 0;
    cx.tr_role = (blockIdx.x != 0 || cx.lane != 0) ? -1
               : (cx.warp == RD_WMMA ? 1 : (cx.warp == RD_WMMA + 1 ? 4 : (cx.warp == RD_WLOAD ? 3 : (cx.warp == 0 ? 0 : (cx.warp == RD_GWARPS ? 2 : -1)))));
    if (threadIdx.x == 0) {
      // a ring stage is refilled (by multicast from every CTA of the cluster) once EVERY CTA's MMAs have retired from it
      for (int s = 0; s < RD_NSTAGE; ++s) { mbar_init(&cx.full()[s], 1); mbar_init(&cx.empty()[s], RD_CLUSTER); }
      for (int g = 0; g < RD_NGROUP; ++g) { mbar_init(&cx.a_ready()[g], RD_GWARPS); mbar_init(&cx.d_ready()[g], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (cx.warp == RD_WMMA) {    // the MMA warp owns the tensor memory: all 512 columns
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(cx.tmem_slot())), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int c = 0; c < f.n_calls; ++c)
      for (int l = 0; l < f.net[c].n_layers; ++l)
        for (int n = threadIdx.x; n < KMAX; n += RD_NTHR)
          cx.sbias()[(c * cx.maxl + l) * KMAX + n] = (n < f.net[c].Np[l]) ? f.net[c].b[l][n] : 0.0f;
    tc_fence_before();
    __syncthreads();
    cx.crank = 0;
    if (RD_CLUSTER > 1) { cx.crank = cluster_ctarank(); cluster_sync_all(); }    // every CTA's mbarriers exist before any remote arrive
    tc_fence_after();
    cx.tmem = *cx.tmem_slot();
    cx.lane_addr = cx.tmem + ((uint32_t)(cx.q << 5) << 16) + 256u * (uint32_t)cx.g;
    cx.alo_row = smem_u32(smem_base() + cx.o_alo) + (uint32_t)cx.row * 16u;
  }

  static __device__ __forceinline__ void fini(CtxD& cx) {
    mma_drain(cx);
    tc_fence_before();
    __syncthreads();
    if (RD_CLUSTER > 1) cluster_sync_all();        // no CTA leaves while a peer may still multicast into it / arrive on its barriers
    if (cx.warp == RD_WMMA) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "r"(512));
    }
  }

  static __device__ __forceinline__ void advance(CtxD& cx) {
    if (++cx.stage == RD_NSTAGE) { cx.stage = 0; cx.phase ^= 1u; }
  }

  // layer-0 bias of every call of one evaluation with the (row-uniform) time features folded in.
  // Threads t0, t0 + nt, ... share the work; the time rows are read from global memory (L1/L2 resident).
  static __device__ __forceinline__ void prep_beff(const CtxD& cx, const FieldDev& f, const float* tfeat, float* buf, int t0, int nt) {
    for (int i = t0; i < f.n_calls * KMAX; i += nt) {
      const int c = i / KMAX, n = i - c * KMAX;
      const int Np0 = f.net[c].Np[0];
      float b = cx.sbias()[(c * cx.maxl) * KMAX + n];
      if (n < Np0) {
        const float* wt = f.net[c].Wt + n;
        for (int j = 0; j < f.net[c].t_dim; ++j) b = fmaf(__ldg(wt + (size_t)j * Np0), tfeat[j], b);
      }
      buf[i] = b;
    }
  }

  // ---- loader warp -----------------------------------------------------------------------------
  static __device__ __forceinline__ void load_net(CtxD& cx, const NetDev& net) {
    if (cx.lane != 0) return;
    for (int l = 0; l < net.n_layers; ++l) {
      const int K = net.K[l], Np = net.Np[l];
      for (int g = 0; g < RD_NGROUP; ++g) {
        if (!((cx.active >> g) & 1u)) continue;
        for (int k0 = 0; k0 < K; k0 += RD_KC) {
          const int rows = min(RD_KC, K - k0);
          mbar_wait(&cx.empty()[cx.stage], cx.phase);
          RD_TRACE(cx, 500 + 10 * l + g);
          const uint32_t bytes = (uint32_t)(rows * Np) * sizeof(float);        // of one half (hi or lo)
          mbar_expect_tx(&cx.full()[cx.stage], 2u * bytes);
          float* dst = cx.ring() + cx.stage * RD_STAGE_FLOATS;
          // packed image: per 32-row chunk [W_hi rows | W_lo rows], each K-major in groups of 4 rows
          const int kc0 = k0 & ~(KC - 1), crow = min(KC, K - kc0);              // the 32-row chunk this stage is part of
          const float* hi = net.W[l] + (size_t)2 * kc0 * Np + (size_t)(k0 - kc0) * Np;
          const float* lo = hi + (size_t)crow * Np;
          if (RD_CLUSTER == 1) {
            bulk_g2s(dst, hi, bytes, &cx.full()[cx.stage]);
            bulk_g2s(dst + rows * Np, lo, bytes, &cx.full()[cx.stage]);
          } else {
            // this CTA's share of the stage: piece `crank` of the hi block and of the lo block, delivered to every CTA
            const uint32_t pf = (uint32_t)(rows * Np) / RD_CLUSTER;             // floats per piece (rows * Np is a multiple of 256)
            const uint16_t mask = (uint16_t)((1u << RD_CLUSTER) - 1u);
            bulk_g2s_mc(dst + cx.crank * pf, hi + cx.crank * pf, pf * 4u, &cx.full()[cx.stage], mask);
            bulk_g2s_mc(dst + rows * Np + cx.crank * pf, lo + cx.crank * pf, pf * 4u, &cx.full()[cx.stage], mask);
          }
          advance(cx);
        }
      }
    }
  }

  // ---- MMA warp: (layer l, group 0), (layer l, group 1), (layer l+1, group 0) ... ------------------------------
  // one ring stage: NJ k-steps (compile-time when > 0), cross products first
  // one ring stage: NJ k-steps (compile-time when > 0); per k-step the two cross products (A_hi W_lo from tensor memory,
  // A_lo W_hi from shared memory), then the main products -- the order of the single-tile engine (csrc/tc_probe.cu)
  template <int NJ>
  static __device__ __forceinline__ void issue_stage(uint32_t d_acc, uint32_t a_hi0, uint64_t da_lo0, uint64_t dh0, uint64_t dl0,
                                                     uint64_t kstep, uint32_t idesc, uint32_t acc0, int nj) {
    constexpr uint64_t astep = (uint64_t)(2u * RD_ALO_LBO) >> 4;
#pragma unroll
    for (int j = 0; j < RD_KC / 8; ++j) {
      if ((NJ > 0) ? (j < NJ) : (j < nj)) {
        tc_mma_ts(d_acc, a_hi0 + 8u * j, dl0 + (uint64_t)j * kstep, idesc, (j == 0) ? acc0 : 1u);
        tc_mma_ss(d_acc, da_lo0 + (uint64_t)j * astep, dh0 + (uint64_t)j * kstep, idesc, 1u);
      }
    }
#pragma unroll
    for (int j = 0; j < RD_KC / 8; ++j)
      if ((NJ > 0) ? (j < NJ) : (j < nj)) tc_mma_ts(d_acc, a_hi0 + 8u * j, dh0 + (uint64_t)j * kstep, idesc, 1u);
  }
  // RD_MMA_WARPS == 2: the two MMA warps alternate ring stages.  Warp w issues the MMAs of every chunk that lands in
  // stage w: while one warp's 12 MMAs run, the other has already waited for its weights (and for a_ready when its chunk
  // opens a layer) and prepared its descriptors, so the hand-over costs one named-barrier wake-up instead of the whole
  // wait + fence + descriptor sequence (measured with -DFFB_TRACE: ~300 cycles per chunk on a single issuing warp).
  // The issue ORDER stays the order of the chunks (a turn token goes back and forth on named barriers 12 / 13), so the
  // accumulation order -- and therefore every bit of the result -- is the same as with one warp.
  static __device__ __forceinline__ void turn_wait(int mw) { asm volatile("bar.sync %0, 64;" ::"r"(12 + mw) : "memory"); }
  static __device__ __forceinline__ void turn_pass(int mw) { asm volatile("bar.arrive %0, 64;" ::"r"(12 + (mw ^ 1)) : "memory"); }
  static __device__ __forceinline__ void mma_net(CtxD& cx, const NetDev& net) {
    const uint32_t alo0 = smem_u32(smem_base() + cx.o_alo);       // group 0's image (the special warps have g = 0)
    for (int l = 0; l < net.n_layers; ++l) {
      const int K = net.K[l], Np = net.Np[l];
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t lbo = (uint32_t)Np * 16u;
      const uint64_t kstep = (uint64_t)(lbo >> 3);                 // descriptor increment of one k-step of W (2 lbo bytes >> 4)
      for (int g = 0; g < RD_NGROUP; ++g) {
        if (!((cx.active >> g) & 1u)) continue;
        const uint32_t d_acc = cx.tmem + 256u * (uint32_t)g;
        const uint32_t a_hi = d_acc + 128u;
        const uint32_t alo = alo0 + (uint32_t)g * cx.alo_gstride;
        uint32_t& ph = g ? cx.ph_a1 : cx.ph_a0;
        // a TURN = RD_TURN consecutive ring stages issued by one warp
        for (int k0 = 0; k0 < K; k0 += RD_KC * RD_TURN, ++cx.tctr) {
          const int nsub = min(RD_TURN, (K - k0 + RD_KC - 1) / RD_KC);
          const uint32_t c0 = cx.cctr;                                           // ring position of the turn's first stage
          cx.cctr += (uint32_t)nsub;
          if (RD_MMA_WARPS == 2 && (int)(cx.tctr & 1u) != cx.mw) continue;      // the other MMA warp's turn
#pragma unroll
          for (int u = 0; u < RD_TURN; ++u)
            if (u < nsub) mbar_wait(&cx.full()[(c0 + u) % RD_NSTAGE], ((c0 + u) / RD_NSTAGE) & 1u);   // weight rows landed
          RD_TRACE(cx, 300 + (k0 >> 4));
          if (RD_MMA_WARPS == 2) {
            // named barriers count WARPS: make sure the warp executes bar.sync in one piece
            __syncwarp();
            if (cx.tctr > 0) turn_wait(cx.mw);                    // the previous turn's MMAs have been issued
          }
          // The group's A operand of this layer is complete.  With two MMA warps this wait must come AFTER the turn token:
          // a warp opens only some of the layers, so it skips phases of a_ready[g], and a parity wait is only unambiguous
          // when the phase before the awaited one is known to be over -- which the token guarantees (the other warp
          // waited for that phase before it issued the turn that precedes this one).
          if (k0 == 0) { mbar_wait(&cx.a_ready()[g], ph); RD_TRACE(cx, 100 + 10 * l + g); }
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < RD_TURN; ++u) {
              if (u < nsub) {
                const int ks = k0 + u * RD_KC;
                const int nj = min(RD_KC, K - ks) >> 3;
                const uint32_t stage = (c0 + u) % RD_NSTAGE;
                const uint32_t hi_base = smem_u32(cx.ring() + stage * RD_STAGE_FLOATS);
                const uint64_t dh0 = tc_desc(hi_base, lbo, 128u);
                const uint64_t dl0 = tc_desc(hi_base + (uint32_t)(nj * 8 * Np) * 4u, lbo, 128u);
                const uint64_t da0 = tc_desc(alo + (uint32_t)(ks >> 2) * RD_ALO_LBO, RD_ALO_LBO, 128u);
                const uint32_t acc = ks ? 1u : 0u;
                if (nj == RD_KC / 8) issue_stage<RD_KC / 8>(d_acc, a_hi + (uint32_t)ks, da0, dh0, dl0, kstep, idesc, acc, nj);
                else issue_stage<0>(d_acc, a_hi + (uint32_t)ks, da0, dh0, dl0, kstep, idesc, acc, nj);
                // frees the ring stage (in every CTA of the cluster) when these MMAs retire
                if (RD_CLUSTER == 1) tc_commit(&cx.empty()[stage]);
                else tc_commit_mc(&cx.empty()[stage], (uint16_t)((1u << RD_CLUSTER) - 1u));
                if (ks + RD_KC >= K) tc_commit(&cx.d_ready()[g]);                // the group's accumulator of this layer is complete
              }
            }
          }
          __syncwarp();
          if (RD_MMA_WARPS == 2) {
            tc_fence_before();
            turn_pass(cx.mw);
          }
          RD_TRACE(cx, 400 + (k0 >> 4));
        }
        ph ^= 1u;
      }
    }
  }
  // with two MMA warps: consume the turn token the last chunk handed over (keeps the named barriers balanced)
  static __device__ __forceinline__ void mma_drain(CtxD& cx) {
    if (RD_MMA_WARPS == 2 && cx.warp >= RD_WMMA && cx.tctr > 0 && (int)(cx.tctr & 1u) == cx.mw) { __syncwarp(); turn_wait(cx.mw); }
  }

  // ---- compute warps ---------------------------------------------------------------------------
  // "my part of the group's next A operand is written": TMEM stores retired, shared-memory stores visible to the
  // async proxy, then one arrive per warp
  static __device__ __forceinline__ void signal_a(CtxD& cx) {
    tc_wait_st();
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (cx.lane == 0) mbar_arrive(&cx.a_ready()[cx.g]);
  }
  static __device__ __forceinline__ void wait_d(CtxD& cx) {
    mbar_wait(&cx.d_ready()[cx.g], cx.ph_d);
    cx.ph_d ^= 1u;
    tc_fence_after();
  }
  // this thread's row, columns [c0, c0 + 8) of the group's next A operand
  static __device__ __forceinline__ void store_a8(const CtxD& cx, int c0, const uint32_t (&hi)[8], const uint32_t (&lo)[8]) {
    tc_st8(cx.lane_addr + 128u + (uint32_t)c0, hi);
    const uint32_t p = cx.alo_row + (uint32_t)(c0 >> 2) * RD_ALO_LBO;
    sts128(p, lo[0], lo[1], lo[2], lo[3]);
    sts128(p + RD_ALO_LBO, lo[4], lo[5], lo[6], lo[7]);
  }

  // layer-0 operand of call c from the stage input / the conditional: this thread's row, its 8-column blocks
  static __device__ __forceinline__ void build_A(CtxD& cx, const FieldDev& f, int c) {
    const NetDev& net = f.net[c];
    const int K0 = net.K[0], xd = net.x_dim, cd = net.c_dim;
    const float* yc = rd_ycur<MEM>(cx) + f.in_off[c] * LDA + cx.row;
    const float* cb = rd_cond<MEM>(cx) + cx.row;
    for (int k8 = 8 * cx.cg; k8 < K0; k8 += 16) {                // warp-uniform
      uint32_t hi[8], lo[8];
      if (k8 + 8 <= xd) {                                        // state columns only
        float vx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) vx[j] = yc[(k8 + j) * LDA];
#pragma unroll
        for (int j = 0; j < 8; ++j) tf32_split(vx[j], hi[j], lo[j]);
      } else if (k8 >= xd + cd) {                                // zero padding only
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
      store_a8(cx, k8, hi, lo);
    }
    signal_a(cx);
    RD_TRACE(cx, 11);
  }

  // hidden layers: accumulator block -> + bias -> activation -> TF32 split -> next A operand; hand-off per layer
  static __device__ __forceinline__ void hidden(CtxD& cx, const NetDev& net, int c, const float* beff) {
    if (GEN) { FFB_ACT_DISPATCH(net.act, hidden_act<ACT>(cx, net, c, beff)); }
    else hidden_act<FFB_ACT_SILU>(cx, net, c, beff);
  }
  template <int ACT>
  static __device__ __forceinline__ void hidden_act(CtxD& cx, const NetDev& net, int c, const float* beff) {
    for (int l = 0; l + 1 < net.n_layers; ++l) {
      const int nb = net.Np[l] >> 4;                             // 8-column blocks of this thread: 16 b + 8 cg
      const float* bias = ((l == 0) ? beff : cx.sbias() + (c * cx.maxl + l) * KMAX) + 8 * cx.cg;
      const uint32_t dcol = cx.lane_addr + 8u * (uint32_t)cx.cg;
      wait_d(cx);
      RD_TRACE(cx, 200 + 10 * l);
      uint32_t m[2][8];
      tc_ld8(dcol, m[0]);
#pragma unroll
      for (int b = 0; b < KMAX / 16; ++b) {
        if (b < nb) {
          const int c0 = 16 * b + 8 * cx.cg;
          tc_wait_ld();
          if (b + 1 < nb) tc_ld8(dcol + 16u * (uint32_t)(b + 1), m[(b + 1) & 1]);
          const float4 b0 = *reinterpret_cast<const float4*>(bias + 16 * b);
          const float4 b1 = *reinterpret_cast<const float4*>(bias + 16 * b + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint32_t hi[8], lo[8];
          if (ACT == FFB_ACT_SILU) {
            // SiLU + TF32 split on packed FP32 pairs (add/mul/fma.f32x2); sigmoid on the SFU (ex2 + rcp)
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
              const float2 z = __fadd2_rn(make_float2(__uint_as_float(m[b & 1][u]), __uint_as_float(m[b & 1][u + 1])),
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
              const float z = __uint_as_float(m[b & 1][u]) + bb[u];
              tf32_split(act_fwd<ACT>(z), hi[u], lo[u]);
            }
          }
          store_a8(cx, c0, hi, lo);
        }
      }
      signal_a(cx);
      RD_TRACE(cx, 201 + 10 * l);
    }
  }

  // last layer: fn(c0, o[8]) with o[u] = raw output + bias of column c0 + u, for every 8-column block this
  // thread owns that holds a real column (columns >= N[last] of the block are padding)
  template <class F>
  static __device__ __forceinline__ void last(CtxD& cx, const NetDev& net, int c, const float* beff, F&& fn) {
    const int nl = net.n_layers, Nreal = net.N[nl - 1];
    const float* bias = (nl == 1) ? beff : cx.sbias() + (c * cx.maxl + nl - 1) * KMAX;
    wait_d(cx);
    RD_TRACE(cx, 290);
    for (int c0 = 8 * cx.cg; c0 < Nreal; c0 += 16) {             // warp-uniform trip count
      uint32_t m[8];
      tc_ld8(cx.lane_addr + (uint32_t)c0, m);
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

  // ---- one evaluation of the field at the stage input: derivative -> slot dst -------------------------
  // Every warp of the CTA calls this (the special warps serve every active group); on return the compute warps of a
  // lane quarter have passed their quarter barrier, i.e. slot dst is complete for the rows of that quarter.
  // `overlap()` runs on the compute warps right after the layer-0 operand has been handed to the MMA warp.
  struct NoOverlap { __device__ __forceinline__ void operator()() const {} };
  template <class OV = NoOverlap>
  static __device__ __forceinline__ void eval(CtxD& cx, const FieldDev& f, float ev_a, float ev_c, float ev_sigma, float ev_sign,
                                              const float* beff, int dst, unsigned call_mask = 3u, OV&& overlap = OV()) {
    eval_ev(cx, f, EvVals{ev_a, ev_c, ev_sigma, ev_sign}, beff, dst, call_mask, static_cast<OV&&>(overlap));
  }
  template <class EV, class OV = NoOverlap>
  static __device__ __forceinline__ void eval_ev(CtxD& cx, const FieldDev& f, const EV& ev, const float* beff, int dst,
                                                 unsigned call_mask = 3u, OV&& overlap = OV()) {
    bool first = true;
    for (int c = 0; c < f.n_calls; ++c) {
      if (!((call_mask >> c) & 1u)) continue;
      const NetDev& net = f.net[c];
      if (cx.warp == RD_WLOAD) { load_net(cx, net); continue; }
      if (cx.warp >= RD_WMMA) { mma_net(cx, net); continue; }
      rd_qbar(cx);                           // the stage input of this lane quarter is final
      RD_TRACE(cx, 10);
      build_A(cx, f, c);
      if (first) { overlap(); first = false; }
      hidden(cx, net, c, beff + c * KMAX);
      float* kd = rd_slot<MEM>(cx, dst) + cx.row;
      const float* yc = rd_ycur<MEM>(cx) + cx.row;
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
    RD_TRACE(cx, 12);
    if (!cx.producer) rd_qbar(cx);
    RD_TRACE(cx, 13);
  }
};

// ---- row-local stage algebra helpers --------------------------------------------------------------
// A thread owns, in its row, the state columns d with (d >> 3) & 1 == cg: blocks of 8 columns starting at
// d0 = 8 cg + 16 b.  Loads are unconditional (columns past SD clamped to SD - 1) so the 8 loads are in flight together.
template <class F>
__device__ __forceinline__ void rd_for_blocks(const CtxD& cx, F&& fn) {
  for (int d0 = 8 * cx.cg; d0 < cx.SD; d0 += 16) fn(d0);
}
__device__ __forceinline__ void rd_load8(const CtxD& cx, const float* buf, int d0, float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = buf[min(d0 + u, cx.SD - 1) * LDA + cx.row];
}
__device__ __forceinline__ void rd_load8_if(const CtxD& cx, bool on, const float* buf, int d0, float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = on ? buf[min(d0 + u, cx.SD - 1) * LDA + cx.row] : 0.0f;
}
__device__ __forceinline__ void rd_store8(const CtxD& cx, float* buf, int d0, const float (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (d0 + u < cx.SD) buf[(d0 + u) * LDA + cx.row] = v[u];
}

}  // namespace ffb

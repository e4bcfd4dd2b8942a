// ffb_engine_tc.cuh -- tensor-core tile engine: the fused MLP forward (+ forward-mode tangents)
// of a 128-row tile on tcgen05 (5th-gen tensor cores), FP32-faithful through a 3xTF32 split.
//
//   D[m][n] = sum_k A[m][k] W[n][k]      M = 128 rows of the tile, N, K <= 128
//
//   * activations never leave the SM: the A operand lives in TENSOR MEMORY (A_hi | A_lo, written
//     by the row-owning threads with tcgen05.st), accumulators live in TMEM too;
//   * weights are packed once (ffb_net_create) into the canonical no-swizzle K-major UMMA image,
//     already split W = W_hi + W_lo (both rounded to TF32 with cvt.rna), and streamed from L2
//     with cp.async.bulk into a 3-stage ring guarded by full/empty mbarriers;
//   * 3xTF32: D_main = A_hi W_hi ; D_cross = A_hi W_lo + A_lo W_hi (a separate accumulator: adding
//     the 2^-11-scaled cross terms into the big accumulator costs ~3x in accuracy because the
//     tensor core's accumulate truncates -- measured with csrc/tc_probe.cu); the epilogue adds
//     D_main + D_cross in FP32, then bias, SiLU (SFU ex2 + rcp), RN split, tcgen05.st of the next A;
//   * TMEM: D_main | D_cross | A_hi | A_lo = 4 x 128 columns = all 512 columns of the SM;
//   * warp roles: warps 0-7 epilogue + stage algebra (row = TMEM lane, two warps per lane quarter
//     split the columns), warp 8 = weight loader, warp 9 = MMA issuer (one thread) + TMEM owner.
//
// Tangent rows (divergence trace): rows [S, S(1+T)) carry d(activation)/dx_j; their epilogue is
// a multiply by silu'(z) of the matching primal row (gate buffer in shared memory).
#pragma once
#include "ffb_engine.cuh"

namespace ffb {

constexpr int TC_NTHR = NCOMP + 64;              // + loader warp + MMA warp
constexpr int TC_STAGE_FLOATS = 2 * KC * KMAX;   // W_hi | W_lo chunk of 32 k-rows: 32 KB
constexpr int TC_NSTAGE = 4;                     // one whole 128x128 layer (hi+lo) in flight
constexpr int ZS = LDA;                          // row stride of the z / gate buffers
constexpr uint32_t TM_COL_DMAIN = 0, TM_COL_DCROSS = 128, TM_COL_AHI = 256, TM_COL_ALO = 384;

// ---------------------------------------------------------------------------------------------
// tcgen05 PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(addr));
}
__device__ __forceinline__ void tc_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// one elected lane of a fully converged warp (the MMA / commit instructions are issued by one thread,
// but every operand is computed warp-uniformly so it can live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// D[tmem_d] (+)= A[tmem_a] * B[smem desc]   (kind::tf32, cta_group::1, A from tensor memory)
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// canonical no-swizzle K-major shared-memory matrix descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// round-to-nearest TF32 split: a ~= hi + lo, |a - hi - lo| <= 2^-24 |a|
__device__ __forceinline__ void tf32_split_rn(float a, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(a));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(a - __uint_as_float(hi)));
}
// hot-path variant: cvt.rna.tf32.f32 is emulated on sm_100a (FSETP + predicated IADD + LOP3 + a
// constant move per conversion).  hi = round-half-away on the integer image (2 ALU ops, identical to
// cvt.rna for finite inputs); lo = a - hi is exact in FP32 and is left unrounded: the tensor core
// ignores the low 13 mantissa bits of a TF32 operand, a truncation error of 2^-21 |a|.
__device__ __forceinline__ void tf32_split(float a, uint32_t& hi, uint32_t& lo) {
#ifdef FFB_SPLIT_CVT
  tf32_split_rn(a, hi, lo);
#else
  hi = (__float_as_uint(a) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
#endif
}

// optional timeline trace (debug): CTA 0 records clock64() at layer hand-offs
__device__ long long* g_trace = nullptr;
__device__ int g_trace_pos = 0;
__device__ __forceinline__ void trace(int tag, int layer) {
  if (g_trace && blockIdx.x == 0) {
    const int i = atomicAdd(&g_trace_pos, 1);
    if (i < 4096) { g_trace[2 * i] = clock64(); g_trace[2 * i + 1] = tag * 100 + layer; }
  }
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct CtxT {
  uint32_t o_ring, o_ycur, o_cond, o_prb, o_zb, o_gate, o_out, o_beff, o_sbias, o_klp, o_red, o_bar, o_slots, o_wt;
  int tdim;
  int slots_smem;
  __device__ __forceinline__ float* ring() const { return reinterpret_cast<float*>(smem_base() + o_ring); }
  __device__ __forceinline__ float* ycur() const { return reinterpret_cast<float*>(smem_base() + o_ycur); }
  __device__ __forceinline__ float* condb() const { return reinterpret_cast<float*>(smem_base() + o_cond); }
  __device__ __forceinline__ float* prb() const { return reinterpret_cast<float*>(smem_base() + o_prb); }
  __device__ __forceinline__ float* zb() const { return reinterpret_cast<float*>(smem_base() + o_zb); }
  __device__ __forceinline__ float* gate() const { return reinterpret_cast<float*>(smem_base() + o_gate); }
  __device__ __forceinline__ float* outb() const { return reinterpret_cast<float*>(smem_base() + o_out); }
  __device__ __forceinline__ float* stage_buf() const { return outb(); }
  __device__ __forceinline__ float* beff() const { return reinterpret_cast<float*>(smem_base() + o_beff); }
  __device__ __forceinline__ float* sbias() const { return reinterpret_cast<float*>(smem_base() + o_sbias); }
  __device__ __forceinline__ float* swt() const { return reinterpret_cast<float*>(smem_base() + o_wt); }
  __device__ __forceinline__ float* klp() const { return reinterpret_cast<float*>(smem_base() + o_klp); }
  __device__ __forceinline__ double* red() const { return reinterpret_cast<double*>(smem_base() + o_red); }
  __device__ __forceinline__ uint64_t* full() const { return reinterpret_cast<uint64_t*>(smem_base() + o_bar); }
  __device__ __forceinline__ uint64_t* empty() const { return full() + TC_NSTAGE; }
  __device__ __forceinline__ uint64_t* a_ready() const { return full() + 2 * TC_NSTAGE; }
  __device__ __forceinline__ uint64_t* d_ready() const { return full() + 2 * TC_NSTAGE + 1; }
  __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(full() + 2 * TC_NSTAGE + 2); }
  float* scr;
  int S, T, SD, CD;
  int tid, lane, warp;
  bool producer;        // every non-compute warp (kernels only test this)
  uint32_t tmem;        // TMEM base address (lane 0, column 0)
  uint32_t lane_addr;   // TMEM address of this thread's lane quarter
  int half;             // which half of the columns this compute warp owns
  int row;              // tile row (= TMEM lane) of this compute thread
  int stage; uint32_t phase;      // weight ring (loader / MMA warp)
  uint32_t ph_a, ph_d;            // parities of a_ready (MMA warp) / d_ready (compute warps)
};

__device__ __forceinline__ float* slot_ptr(const CtxT& cx, int slot) {
  float* base = cx.slots_smem ? reinterpret_cast<float*>(smem_base() + cx.o_slots) : cx.scr;
  return base + (size_t)slot * cx.SD * LDA;
}
// compile-time variant: SS = slots in shared memory (LDS/STS) or in the global scratch (LDG/STG);
// the run-time ternary above degrades every access to a generic LD/ST
template <bool SS>
__device__ __forceinline__ float* slot_ptr_t(const CtxT& cx, int slot) {
  if (SS) return reinterpret_cast<float*>(smem_base() + cx.o_slots) + (size_t)slot * cx.SD * LDA;
  return cx.scr + (size_t)slot * cx.SD * LDA;
}
// one lane polls the mbarrier, the warp then reconverges: 32x less polling traffic on the shared-memory
// pipe that also feeds the tensor core's B operand
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
#ifdef FFB_POLL_ONE
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
#else
  (void)lane;
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void pipe_advance(CtxT& cx) {
  if (++cx.stage == TC_NSTAGE) { cx.stage = 0; cx.phase ^= 1u; }
}

__host__ __device__ inline size_t smem_layout_tc(int SD, int CD, int T, int hutch, int slots_smem, int ncalls, int tdim,
                                                 size_t* off /*[16]*/) {
  const int S = TM / (1 + T);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  size_t v[16];
  v[0] = take(sizeof(float) * TC_NSTAGE * TC_STAGE_FLOATS);       // ring
  v[1] = take(sizeof(float) * SD * LDA);                          // ycur
  v[2] = take(sizeof(float) * (CD > 0 ? CD : 1) * LDA);           // cond
  v[3] = take(hutch ? sizeof(float) * SD * LDA : 0);              // probes
  v[4] = take(T > 0 ? sizeof(float) * S * ZS : 0);                // z / activation of primal rows
  v[5] = take(T > 0 ? sizeof(float) * S * ZS : 0);                // gate = silu'(z)
  v[6] = take(sizeof(float) * SD * LDA);                          // raw output of the last layer [n][row]
  v[7] = take(sizeof(float) * KMAX);                              // beff
  v[8] = take(sizeof(float) * ncalls * NET_MAXL * KMAX);    // biases, every layer of every network
  v[9] = take(T > 0 ? sizeof(float) * (NSLOT + 2) * TM : 0);      // klp
  v[10] = take(sizeof(double) * 8 * FFB_NPART);                   // red
  v[11] = take(sizeof(uint64_t) * (2 * TC_NSTAGE + 4));           // barriers + tmem slot
  v[12] = take(slots_smem ? sizeof(float) * NSLOT * SD * LDA : 0);
  v[13] = take(sizeof(float) * ncalls * (tdim > 0 ? tdim : 1) * KMAX);   // layer-0 time-feature rows
  if (off) for (int i = 0; i < 14; ++i) off[i] = v[i];
  return o;
}
__host__ __device__ inline int field_tdim(const FieldDev& f) {
  int t = f.net[0].t_dim;
  if (f.n_calls > 1 && f.net[1].t_dim > t) t = f.net[1].t_dim;
  return t;
}

struct EngineTC {
  using Ctx = CtxT;
  static constexpr int NTHR = TC_NTHR;

  static __device__ __forceinline__ void init(CtxT& cx, const FieldDev& f, float* scratch) {
    const int T = (f.div_mode == FFB_DIV_EXACT) ? f.net[0].x_dim : (f.div_mode == FFB_DIV_HUTCH ? 1 : 0);
    size_t off[16];
    cx.tdim = field_tdim(f);
    smem_layout_tc(f.state_dim, f.cond_dim, T, f.div_mode == FFB_DIV_HUTCH, f.slots_smem, f.n_calls, cx.tdim, off);
    cx.o_ring = (uint32_t)off[0]; cx.o_ycur = (uint32_t)off[1]; cx.o_cond = (uint32_t)off[2];
    cx.o_prb = (uint32_t)off[3]; cx.o_zb = (uint32_t)off[4]; cx.o_gate = (uint32_t)off[5];
    cx.o_out = (uint32_t)off[6]; cx.o_beff = (uint32_t)off[7]; cx.o_sbias = (uint32_t)off[8];
    cx.o_klp = (uint32_t)off[9]; cx.o_red = (uint32_t)off[10]; cx.o_bar = (uint32_t)off[11];
    cx.o_slots = (uint32_t)off[12];
    cx.o_wt = (uint32_t)off[13];
    cx.slots_smem = f.slots_smem;
    cx.T = T; cx.S = TM / (1 + T); cx.SD = f.state_dim; cx.CD = f.cond_dim;
    cx.tid = threadIdx.x; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
    cx.producer = cx.warp >= NCOMP / 32;
    cx.half = (cx.warp >> 2) & 1;
    cx.row = ((cx.warp & 3) << 5) + cx.lane;
    cx.scr = scratch + (size_t)blockIdx.x * NSLOT * f.state_dim * LDA;
    cx.stage = 0;
    cx.phase = (cx.warp == NCOMP / 32) ? 1u : 0u;     // loader starts with all stages free
    cx.ph_a = 0; cx.ph_d = 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(&cx.full()[s], 1); mbar_init(&cx.empty()[s], 1); }
      mbar_init(cx.a_ready(), NCOMP / 32);
      mbar_init(cx.d_ready(), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (cx.warp == NCOMP / 32 + 1) {    // the MMA warp owns the tensor memory: all 512 columns
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(cx.tmem_slot())), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // biases of every layer to shared memory once (real column order, zero padded)
    for (int c = 0; c < f.n_calls; ++c) {
      for (int l = 0; l < f.net[c].n_layers; ++l)
        for (int n = threadIdx.x; n < KMAX; n += TC_NTHR)
          cx.sbias()[(c * NET_MAXL + l) * KMAX + n] = (n < f.net[c].Np[l]) ? f.net[c].b[l][n] : 0.0f;
      for (int i = threadIdx.x; i < f.net[c].t_dim * KMAX; i += TC_NTHR) {
        const int j = i / KMAX, n = i - j * KMAX;
        cx.swt()[(c * cx.tdim + j) * KMAX + n] = (n < f.net[c].Np[0]) ? f.net[c].Wt[(size_t)j * f.net[c].Np[0] + n] : 0.0f;
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cx.tmem = *cx.tmem_slot();
    cx.lane_addr = cx.tmem + ((uint32_t)((cx.warp & 3) << 5) << 16);
  }

  static __device__ __forceinline__ void fini(CtxT& cx) {
    tc_fence_before();
    __syncthreads();
    if (cx.warp == NCOMP / 32 + 1) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "r"(512));
    }
  }

  // ---- loader warp: stream one network's W_hi|W_lo chunks through the ring -------------------
  static __device__ __forceinline__ void load_net(CtxT& cx, const NetDev& net) {
    if (cx.lane != 0) return;
    for (int l = 0; l < net.n_layers; ++l) {
      const int K = net.K[l], Np = net.Np[l];
      for (int k0 = 0; k0 < K; k0 += KC) {
        const int rows = min(KC, K - k0);
        mbar_wait(&cx.empty()[cx.stage], cx.phase);
        const uint32_t bytes = (uint32_t)(2 * rows * Np) * sizeof(float);
        mbar_expect_tx(&cx.full()[cx.stage], bytes);
        bulk_g2s(cx.ring() + cx.stage * TC_STAGE_FLOATS, net.W[l] + (size_t)2 * k0 * Np, bytes, &cx.full()[cx.stage]);
        pipe_advance(cx);
      }
    }
  }

  // ---- MMA warp: issue the 3xTF32 contraction of every layer ----------------------------------
  // Measured (csrc/tc_rate.cu): a tcgen05.mma kind::tf32 M128 N128 K8 retires every ~67 cycles
  // whatever the accumulator pattern, so issue overhead must stay well below that.  The warp runs
  // convergently so descriptors / addresses live in uniform registers (issuing from a divergent
  // single-lane region makes the compiler wrap every MMA in an ELECT + 5x R2UR waterfall loop),
  // and there is ONE elected region per ring stage (12 MMAs + commit), not one per MMA.
  static __device__ __forceinline__ void mma_net(CtxT& cx, const NetDev& net) {
    for (int l = 0; l < net.n_layers; ++l) {
      const int K = net.K[l], Np = net.Np[l];
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t lbo = (uint32_t)Np * 16u;        // bytes between the two 16-byte K chunks of one MMA
      const uint64_t kstep = (uint64_t)(lbo >> 3);    // descriptor increment of one MMA k-step (2*lbo bytes >> 4)
      const uint32_t d_main = cx.tmem + TM_COL_DMAIN, d_cross = cx.tmem + TM_COL_DCROSS;
      mbar_wait_warp(cx.a_ready(), cx.ph_a, cx.lane);   // the epilogue has written this layer's A operand
      tc_fence_after();
      cx.ph_a ^= 1u;
      if (cx.lane == 0) trace(1, l);
      uint32_t acc = 0;
      for (int k0 = 0; k0 < K; k0 += KC) {
        const int nj = min(KC, K - k0) >> 3;           // MMA k-steps in this stage (1..4)
        mbar_wait_warp(&cx.full()[cx.stage], cx.phase, cx.lane);
        tc_fence_after();
        const uint32_t hi_base = smem_u32(cx.ring() + cx.stage * TC_STAGE_FLOATS);
        const uint64_t dh0 = tc_desc(hi_base, lbo, 128u);
        const uint64_t dl0 = tc_desc(hi_base + (uint32_t)(nj * 8 * Np) * 4u, lbo, 128u);
        const uint32_t a_hi0 = cx.tmem + TM_COL_AHI + (uint32_t)k0, a_lo0 = cx.tmem + TM_COL_ALO + (uint32_t)k0;
        uint64_t* ebar = &cx.empty()[cx.stage];
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < KC / 8; ++j) {
            if (j < nj) {
              const uint64_t d_hi = dh0 + (uint64_t)j * kstep, d_lo = dl0 + (uint64_t)j * kstep;
              tc_mma_ts(d_main, a_hi0 + 8u * j, d_hi, idesc, (j == 0) ? acc : 1u);
              tc_mma_ts(d_cross, a_hi0 + 8u * j, d_lo, idesc, (j == 0) ? acc : 1u);
              tc_mma_ts(d_cross, a_lo0 + 8u * j, d_hi, idesc, 1u);
            }
          }
          tc_commit(ebar);                              // frees the ring stage when these MMAs retire
        }
        __syncwarp();
        acc = 1u;
        pipe_advance(cx);
      }
      if (elect_one()) tc_commit(cx.d_ready());        // accumulators of this layer are complete
      __syncwarp();
      if (cx.lane == 0) trace(2, l);
    }
  }

  // compute warps: signal "A operand of the next layer is in tensor memory"
  static __device__ __forceinline__ void signal_a_ready(CtxT& cx) {
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (cx.lane == 0) mbar_arrive(cx.a_ready());
  }
  static __device__ __forceinline__ void wait_d_ready(CtxT& cx) {
    mbar_wait_warp(cx.d_ready(), cx.ph_d, cx.lane);
    cx.ph_d ^= 1u;
    tc_fence_after();
  }

  // ---- evaluate the vector field at cx.ycur(); derivative -> slot dst, divergence -> klp[dst] ----
  template <bool SS>
  static __device__ __forceinline__ void eval(CtxT& cx, const FieldDev& f, const ffb_eval_scalars& ev, int dst,
                                              unsigned call_mask = 3u) {
    for (int c = 0; c < f.n_calls; ++c) {
      if (!((call_mask >> c) & 1u)) continue;
      const NetDev& net = f.net[c];
      if (cx.warp == NCOMP / 32) { load_net(cx, net); continue; }
      if (cx.warp == NCOMP / 32 + 1) { mma_net(cx, net); continue; }
      const int S = cx.S, T = cx.T, live = S * (1 + T);
      const int r = cx.row, h = cx.half;
      const int Np0 = net.Np[0], xd = net.x_dim, cd = net.c_dim, K0 = net.K[0];
      if (cx.tid == 0) trace(7, 0);
      // layer-0 bias with the (row-uniform) time features folded in
      for (int n = cx.tid; n < Np0; n += NCOMP) {
        float b = cx.sbias()[(c * NET_MAXL) * KMAX + n];
        const float* wt = cx.swt() + (c * cx.tdim) * KMAX + n;
        for (int j = 0; j < net.t_dim; ++j) b = fmaf(wt[j * KMAX], ev.tfeat[j], b);
        cx.beff()[n] = b;
      }
      // layer-0 A operand: this thread's row, column groups of 8 interleaved between the two halves
      {
        const int jt = (r >= S && r < live) ? (r / S - 1) : -1;      // tangent index of this row
        const int sidx = (r < S) ? r : (r - (jt + 1) * S);
        for (int g = h; g < K0 / 8; g += 2) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = g * 8 + q;
            float val = 0.0f;
            if (r < S) {
              if (k < xd) val = cx.ycur()[(f.in_off[c] + k) * LDA + r];
              else if (k < xd + cd) val = cx.condb()[(k - xd) * LDA + r];
            } else if (jt >= 0 && k < xd) {
              val = (f.div_mode == FFB_DIV_EXACT) ? (k == jt ? 1.0f : 0.0f) : cx.prb()[k * LDA + sidx];
            }
            tf32_split(val, hi[q], lo[q]);
          }
          tc_st8(cx.lane_addr + TM_COL_AHI + g * 8, hi);
          tc_st8(cx.lane_addr + TM_COL_ALO + g * 8, lo);
        }
      }
      bar_compute();                      // beff visible to every epilogue thread
      if (cx.tid == 0) trace(8, 0);
      signal_a_ready(cx);
      if (cx.tid == 0) trace(0, 0);

      const int nl = net.n_layers;
      for (int l = 0; l < nl; ++l) {
        const bool last = (l == nl - 1);
        const int Np = net.Np[l], Nreal = net.N[l];
        const float* bias = (l == 0) ? cx.beff() : cx.sbias() + (c * NET_MAXL + l) * KMAX;
        const int cbeg = h * (Np >> 1), cend = cbeg + (Np >> 1);     // Np is a multiple of 32
        wait_d_ready(cx);
        if (cx.tid == 0) trace(3, l);
        if (last) {
          // raw network output (primal: + bias) -> outb[n][row]
          for (int c0 = cbeg; c0 < cend; c0 += 16) {
            uint32_t m[16], x[16];
            tc_ld16(cx.lane_addr + TM_COL_DMAIN + c0, m);
            tc_ld16(cx.lane_addr + TM_COL_DCROSS + c0, x);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int n = c0 + q;
              float z = __uint_as_float(m[q]) + __uint_as_float(x[q]);
              if (r < S) z += bias[n];
              if (n < Nreal) cx.outb()[n * LDA + r] = z;
            }
          }
          tc_fence_before();
        } else if (T == 0) {
          // software-pipelined: the TMEM loads of column group g+1 are in flight while group g is
          // pushed through bias + SiLU + TF32 split and stored back as the next layer's A operand
          const int ng = (cend - cbeg) >> 4;            // 1..4 groups of 16 columns
          uint32_t m[2][16], x[2][16];
          tc_ld16(cx.lane_addr + TM_COL_DMAIN + cbeg, m[0]);
          tc_ld16(cx.lane_addr + TM_COL_DCROSS + cbeg, x[0]);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ng) {
              const int c0 = cbeg + 16 * g;
              tc_wait_ld();
              if (g + 1 < ng) {
                tc_ld16(cx.lane_addr + TM_COL_DMAIN + c0 + 16, m[(g + 1) & 1]);
                tc_ld16(cx.lane_addr + TM_COL_DCROSS + c0 + 16, x[(g + 1) & 1]);
              }
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int q = 0; q < 16; q += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias + c0 + q);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float z = (__uint_as_float(m[g & 1][q + u]) + __uint_as_float(x[g & 1][q + u])) + bb[u];
                  tf32_split(z * sigmoidf_fast(z), hi[q + u], lo[q + u]);
                }
              }
              tc_st16(cx.lane_addr + TM_COL_AHI + c0, hi);
              tc_st16(cx.lane_addr + TM_COL_ALO + c0, lo);
            }
          }
          if (cx.tid == 0) trace(4, l);
          signal_a_ready(cx);
          if (cx.tid == 0) trace(0, l + 1);
        } else {
          // primal rows publish z; all threads share the SiLU / gate work; tangent rows are gated
          float* zb = cx.zb();
          float* gt = cx.gate();
          // tcgen05.ld is warp-collective (.sync.aligned): branch on warp-uniform conditions only
          if (((cx.warp & 3) << 5) < S) {
            for (int c0 = cbeg; c0 < cend; c0 += 16) {
              uint32_t m[16], x[16];
              tc_ld16(cx.lane_addr + TM_COL_DMAIN + c0, m);
              tc_ld16(cx.lane_addr + TM_COL_DCROSS + c0, x);
              tc_wait_ld();
              if (r < S) {
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                  float4 z4;
                  z4.x = (__uint_as_float(m[q]) + __uint_as_float(x[q])) + bias[c0 + q];
                  z4.y = (__uint_as_float(m[q + 1]) + __uint_as_float(x[q + 1])) + bias[c0 + q + 1];
                  z4.z = (__uint_as_float(m[q + 2]) + __uint_as_float(x[q + 2])) + bias[c0 + q + 2];
                  z4.w = (__uint_as_float(m[q + 3]) + __uint_as_float(x[q + 3])) + bias[c0 + q + 3];
                  *reinterpret_cast<float4*>(zb + r * ZS + c0 + q) = z4;
                }
              }
            }
          }
          bar_compute();
          for (int idx = cx.tid; idx < S * Np; idx += NCOMP) {
            const int s = idx / Np, n = idx - s * Np;
            const float z = zb[s * ZS + n];
            const float sg = sigmoidf_fast(z);
            zb[s * ZS + n] = z * sg;
            gt[s * ZS + n] = sg * (1.0f + z * (1.0f - sg));
          }
          bar_compute();
          const int s = (r < S) ? r : (r < live ? r % S : -1);
          for (int c0 = cbeg; c0 < cend; c0 += 16) {
            uint32_t hi[16], lo[16], m[16], x[16];
            tc_ld16(cx.lane_addr + TM_COL_DMAIN + c0, m);      // every lane loads (warp-collective)
            tc_ld16(cx.lane_addr + TM_COL_DCROSS + c0, x);
            tc_wait_ld();
            if (r < S) {
#pragma unroll
              for (int q = 0; q < 16; ++q) tf32_split(zb[s * ZS + c0 + q], hi[q], lo[q]);
            } else if (s >= 0) {
#pragma unroll
              for (int q = 0; q < 16; ++q)
                tf32_split((__uint_as_float(m[q]) + __uint_as_float(x[q])) * gt[s * ZS + c0 + q], hi[q], lo[q]);
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) { hi[q] = 0u; lo[q] = 0u; }
            }
            tc_st16(cx.lane_addr + TM_COL_AHI + c0, hi);
            tc_st16(cx.lane_addr + TM_COL_ALO + c0, lo);
          }
          signal_a_ready(cx);
        }
      }
      bar_compute();                      // outb complete
      if (cx.tid == 0) trace(5, 0);
      // ---- field transform on the primal rows (same arithmetic as the FFMA engine) ----------------
      const int Dout = net.N[nl - 1];
      float* kd = slot_ptr_t<SS>(cx, dst) + f.out_off[c] * LDA;
      const float* ob = cx.outb();
      const float* yc = cx.ycur() + f.out_off[c] * LDA;
      const float sgn = ev.sign * f.out_sign[c];
      {
        const int rr = cx.tid & (TM - 1);
        if (rr < S) {
          for (int d = cx.tid >> 7; d < Dout; d += NCOMP / TM) {
            const float o = ob[d * LDA + rr];
            float xd_;
            if (f.kind == FFB_FIELD_SCORE) {
              const float sc = f.use_sigma ? __fdiv_rn(o, ev.sigma) : o;
              const float lin = f.has_drift ? __fmul_rn(ev.a, yc[d * LDA + rr]) : 0.0f;
              xd_ = __fsub_rn(lin, __fmul_rn(ev.c, sc));
            } else {
              xd_ = o;
            }
            kd[d * LDA + rr] = xd_ * sgn;
          }
        }
      }
      if (T > 0) {
        for (int s = cx.tid; s < S; s += NCOMP) {
          float tr = 0.0f;
          if (f.div_mode == FFB_DIV_EXACT) {
            for (int j = 0; j < T; ++j) tr += ob[j * LDA + (j + 1) * S + s];
          } else {
            for (int d = 0; d < xd; ++d) tr = fmaf(cx.prb()[d * LDA + s], ob[d * LDA + S + s], tr);
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
      if (cx.tid == 0) trace(6, 0);
    }
  }
};

}  // namespace ffb

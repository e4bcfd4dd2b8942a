// tc_rate.cu -- micro-benchmark: cycles per tcgen05.mma (kind::tf32, M=128, K=8) for accumulate chains.
//   NACC = number of independent accumulators used round-robin; SS = A operand from shared memory (else TMEM)
// Run: ./tc_rate   (prints a table)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <bool SS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (SS)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ int g_bstride = 16;     // descriptor step between the B operands of consecutive MMAs (16-byte units)
template <int N, int NACC, bool SS>
__global__ void rate(int reps, long long* out, int bg, int commit_every) {
  __shared__ uint64_t bar2;
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<float*>(smem)[i] = (bg & 8) ? ((float)(h & 0xFFFFFF) / 8388608.0f - 1.0f) : 0.001f * (i % 7);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  if ((bg & 8) && warp >= 2 && warp < 6) {     // random A operand: 96 columns at 384.., this warp's lane quarter
    const uint32_t la = tb + ((uint32_t)((warp & 3) << 5) << 16) + 384u;
    for (int c = 0; c < 96; c += 8) {
      uint32_t v[8];
      for (int u = 0; u < 8; ++u) { uint32_t h = (threadIdx.x * 97u + c + u) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; v[u] = __float_as_uint((float)(h & 0xFFFFFF) / 8388608.0f - 1.0f) & 0xFFFFE000u; }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(la + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  bg &= 7;
  if (threadIdx.x >= 64 && bg) {
    // background load from the other warps until the MMA thread is done
    const int w = (threadIdx.x >> 5) & 3;
    const uint32_t la = tb + ((uint32_t)(w << 5) << 16) + 128u;      // columns 128.. of this warp's lane quarter
    float acc = 0.f; uint32_t v[8] = {0,0,0,0,0,0,0,0};
    int it = 0;
    while (!stop_flag && it < 200000) {
      ++it;
      if (bg == 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(la + (uint32_t)(it & 7) * 8u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(la + 64u + (uint32_t)(it & 7) * 8u), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else if (bg == 2) {
        float4 x;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(smem_u32(smem) + 32768u + (uint32_t)((threadIdx.x * 16 + it * 512) & 16383)));
        acc += x.x + x.w;
      } else {
        float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc)); acc = e * 0.5f;
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc + 1.0f)); acc = e;
      }
    }
    if (acc == 123.f) out[5] = (long long)acc + v[0];
  }
  if (threadIdx.x == 32) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sb = smem_u32(smem);
    const uint64_t db = make_desc(sb, N * 16, 128);
    const uint64_t da = make_desc(sb + 16384, 128 * 16, 128);
    const int bstr = g_bstride;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int i = 0; i < 12; ++i)
        mma<SS>(tb + (uint32_t)(i % NACC) * 128u, tb + 384u + (uint32_t)i * 8u, da, db + (uint64_t)(i * bstr), idesc, (r | (i >= NACC)) ? 1u : 0u);
      if (commit_every && (r % commit_every) == commit_every - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop_flag = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
template <int N, int NACC, bool SS>
void run(long long* d, int grid = 1, int bgwarps = 0, int bg = 0, int ce = 0, int reps = 8) {
  cudaFuncSetAttribute(rate<N, NACC, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    rate<N, NACC, SS><<<grid, 64 + 32 * bgwarps, 64 * 1024>>>(reps, d, bg, ce);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (rep) printf("commit_every=%d grid=%3d bgwarps=%2d bg=%d  N=%3d nacc=%d %s: issue %6lld complete %6lld cyc for %d MMAs -> %.1f cyc/MMA (%s)\n", ce, grid, bgwarps, bg, N, NACC, SS ? "SS" : "TS", h[0], h[1],
                    reps * 12, (double)h[1] / (reps * 12), cudaGetErrorString(e));
  }
}
// long runs on every SM: does the MMA rate (cycles per MMA) or the SM clock give way under sustained load?
template <int N, bool SS>
static void run_long(long long* d, int bgwarps, int bg, int reps) {
  cudaFuncSetAttribute(rate<N, 1, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  long long h[2];
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    rate<N, 1, SS><<<148, 64 + 32 * bgwarps, 64 * 1024>>>(reps, d, bg, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("LONG grid=148 bgwarps=%2d bg=%d N=%3d %s reps=%d: %.1f cyc/MMA, kernel %.3f ms -> SM clock %.0f MHz (%s)\n", bgwarps, bg, N, SS ? "SS" : "TS",
           reps, (double)h[1] / (reps * 12.0), ms, (double)h[1] / (ms * 1e3), cudaGetErrorString(e));
  }
}
int main(int argc, char** argv) {
  long long* d; cudaMalloc(&d, 16);
  if (argc > 1 && argv[1][0] == 'B') {      // B operands of consecutive MMAs 256 B apart (overlapping windows) vs 4 KB apart (distinct)
    for (int st : {16, 256}) {
      cudaMemcpyToSymbol(g_bstride, &st, sizeof(st));
      printf("B-operand stride %d bytes\n", st * 16);
      run<128, 1, false>(d, 148, 0, 0, 0, 512); run<128, 1, true>(d, 148, 0, 0, 0, 512); run<64, 1, false>(d, 148, 0, 0, 0, 512);
    }
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'L') {
    run_long<128, false>(d, 0, 0, 16384);
    run_long<128, false>(d, 16, 3, 16384);
    run_long<128, false>(d, 16, 1, 16384);
    run_long<128, true>(d, 16, 3, 16384);
    return 0;
  }
  run<128, 1, false>(d); run<128, 2, false>(d); run<128, 3, false>(d);
  run<128, 1, false>(d, 148); run<128, 1, false>(d, 1, 16, 1); run<128, 1, false>(d, 1, 16, 2); run<128, 1, false>(d, 1, 16, 3);
  run<128, 1, false>(d, 1, 4, 0, 0, 64); run<128, 1, false>(d, 148, 4, 0, 0, 64); run<128, 1, false>(d, 1, 4, 8, 0, 64); run<128, 1, false>(d, 148, 4, 8, 0, 64); run<128, 1, false>(d, 148, 4, 8, 0, 512);
  run<128, 1, false>(d, 1, 0, 0, 1); run<128, 1, false>(d, 1, 0, 0, 2); run<128, 1, false>(d, 1, 16, 1, 1);
  run<128, 1, false>(d, 148, 16, 1); run<128, 1, false>(d, 148, 16, 3);
  run<128, 1, true>(d);  run<128, 3, true>(d);
  run<64, 1, false>(d);  run<64, 2, false>(d);  run<64, 3, false>(d);
  run<32, 1, false>(d);  run<32, 2, false>(d);  run<32, 3, false>(d); run<32, 1, true>(d);
  run<256, 1, false>(d); run<256, 1, true>(d);
  return 0;
}

// tc_probe.cu -- standalone probe of the tcgen05 pieces the tensor-core tile engine relies on:
//   * A operand in TMEM (written with tcgen05.st by the row-owning threads), B operand in shared
//     memory in the canonical NO-SWIZZLE K-major layout, kind::tf32, FP32 accumulate in TMEM;
//   * 3xTF32 split (hi*hi + hi*lo + lo*hi) accuracy vs a float64 reference;
//   * M=128, N in {128, 16}, K in {128, 24}.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
// Run:   ./tc_probe <N> <K> <lbo_bytes> <sbo_bytes> <passes>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait_timeout(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < 2000000; ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // layout_type = 0: no swizzle
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// out[m][n] = sum_k A[m][k] * W[n][k]
__global__ void __launch_bounds__(160, 1)
probe(const float* __restrict__ A, const float* __restrict__ Bhi_img, const float* __restrict__ Blo_img,
      float* __restrict__ out, int N, int K, uint32_t lbo, uint32_t sbo, int passes, int* status, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sBhi = reinterpret_cast<float*>(smem);
  float* sBlo = sBhi + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < N * K; i += blockDim.x) { sBhi[i] = Bhi_img[i]; sBlo[i] = Blo_img[i]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const uint32_t colD = 0, colAhi = 128, colAlo = 256;

  if (warp < 4) {
    // row m = tid: write A_hi / A_lo rows into TMEM, 8 columns at a time
    const int m = tid;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t hi[8], lo[8];
      for (int j = 0; j < 8; ++j) {
        const float a = A[m * K + k0 + j];
        if (mode & 1) {
          uint32_t h, l;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(a - __uint_as_float(h)));
          hi[j] = h; lo[j] = l;
        } else {
          const float h = tf32_hi(a);
          hi[j] = __float_as_uint(h);
          lo[j] = __float_as_uint(a - h);
        }
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + colAhi + k0),
                   "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + colAlo + k0),
                   "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();

  if (warp == 4 && lane == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t kstep_bytes = 2 * lbo;   // one MMA consumes K = 8 tf32 = two 16-byte K chunks
    uint32_t acc = 0, acc2 = 0;
    const uint32_t colD2 = 384;
    if (mode & 4) {
      // single accumulator, pass-ordered: (mode & 8) ? main first : cross first
      for (int pass = 0; pass < 2; ++pass) {
        const bool do_main = ((mode & 8) != 0) == (pass == 0);
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint64_t dhi = make_desc(smem_u32(sBhi) + ks * kstep_bytes, lbo, sbo);
          const uint64_t dlo = make_desc(smem_u32(sBlo) + ks * kstep_bytes, lbo, sbo);
          const uint32_t a_hi = tb + colAhi + ks * 8, a_lo = tb + colAlo + ks * 8;
          for (int p = (do_main ? 0 : 1); p < (do_main ? 1 : 3); ++p) {
            const uint32_t a = (p == 2) ? a_lo : a_hi;
            const uint64_t d = (p == 1) ? dlo : dhi;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                ::"r"(tb + colD), "r"(a), "l"(d), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
            acc = 1;
          }
        }
      }
    } else
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t dhi = make_desc(smem_u32(sBhi) + ks * kstep_bytes, lbo, sbo);
      const uint64_t dlo = make_desc(smem_u32(sBlo) + ks * kstep_bytes, lbo, sbo);
      const uint32_t a_hi = tb + colAhi + ks * 8, a_lo = tb + colAlo + ks * 8;
      for (int p = 0; p < passes; ++p) {
        const uint32_t a = (p == 2) ? a_lo : a_hi;
        const uint64_t d = (p == 1) ? dlo : dhi;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
            ::"r"(tb + (((mode & 2) && p > 0) ? colD2 : colD)), "r"(a), "l"(d), "r"(idesc),
              "r"(((mode & 2) && p > 0) ? acc2 : acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        if ((mode & 2) && p > 0) acc2 = 1; else acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  if (warp < 4) {
    if (!mbar_wait_timeout(&bar, 0)) { if (lane == 0) atomicOr(status, 1); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    for (int n0 = 0; n0 < N; n0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(lane_addr + colD + n0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (mode & 2) {
        uint32_t w[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "r"(lane_addr + 384 + n0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      }
      for (int j = 0; j < 8; ++j) out[tid * N + n0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}

static int g_rn = 0;
static float tf32_hi_h(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  if (g_rn) u += 0x1000u;            // round to nearest (ties away) on the 13 dropped bits
  u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 128;
  const int K = argc > 2 ? atoi(argv[2]) : 128;
  uint32_t lbo = argc > 3 ? (uint32_t)atoi(argv[3]) : (uint32_t)(N * 16);
  uint32_t sbo = argc > 4 ? (uint32_t)atoi(argv[4]) : 128u;
  const int passes = argc > 5 ? atoi(argv[5]) : 3;
  const int mode = argc > 6 ? atoi(argv[6]) : 0;
  const int M = 128;
  g_rn = mode & 1;
  std::vector<float> A(M * K), W(N * K), imgh(N * K), imgl(N * K), out(M * N, -1.f);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 4.f - 2.f;
  for (auto& v : W) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
  // canonical no-swizzle K-major image: W[n][k] at kc*(N*16 B) + (n/8)*128 B + (n%8)*16 B + (k%4)*4 B
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const size_t off = (size_t)(k / 4) * (N * 4) + (size_t)(n / 8) * 32 + (size_t)(n % 8) * 4 + (k % 4);
      const float h = tf32_hi_h(W[n * K + k]);
      imgh[off] = h;
      imgl[off] = g_rn ? tf32_hi_h(W[n * K + k] - h) : W[n * K + k] - h;
    }
  float *dA, *dh, *dl, *dout; int* dst;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dh, imgh.size() * 4); cudaMalloc(&dl, imgl.size() * 4);
  cudaMalloc(&dout, out.size() * 4); cudaMalloc(&dst, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dh, imgh.data(), imgh.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dl, imgl.data(), imgl.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0, 4);
  cudaMemset(dout, 0xff, out.size() * 4);
  const size_t smem = (size_t)2 * N * K * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 160, smem>>>(dA, dh, dl, dout, N, K, lbo, sbo, passes, dst, mode);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0, fp32err = 0, se = 0, sf = 0, sr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0; float f = 0.f;
      for (int k = 0; k < K; ++k) { r += (double)A[m * K + k] * W[n * K + k]; f = fmaf(A[m * K + k], W[n * K + k], f); }
      maxerr = fmax(maxerr, fabs(r - out[m * N + n]));
      fp32err = fmax(fp32err, fabs(r - f));
      maxref = fmax(maxref, fabs(r));
      se += (r - out[m * N + n]) * (r - out[m * N + n]); sf += (r - f) * (r - f); sr += r * r;
    }
  printf("mode=%d rms err %.3e (fp32 chain %.3e) rms ref %.3e  mean signed err %.3e\n", mode, sqrt(se / (M * N)), sqrt(sf / (M * N)), sqrt(sr / (M * N)), 0.0);
  printf("N=%d K=%d lbo=%u sbo=%u passes=%d : cuda=%s status=%d max|err|=%.3e (fp32 fma chain err %.3e) max|ref|=%.3e  out[0][0..3]=%g %g %g %g\n",
         N, K, lbo, sbo, passes, cudaGetErrorString(e), st, maxerr, fp32err, maxref, out[0], out[1], out[2], out[3]);
  return 0;
}

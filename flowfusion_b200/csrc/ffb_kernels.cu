// ffb_kernels.cu -- integrator kernels and the C ABI of libffb200.so (see include/ffb200.h).
//
// Kernels (all persistent: grid = min(#tiles, #SMs), one 288-thread CTA per SM):
//   k_field_eval   one evaluation of the vector field (+ torchdiffeq initial-step norms)
//   k_dopri5       one attempted Dormand-Prince step: 6 fused evaluations, FSAL, error norm
//                  partials in FP64, dense-output interpolant at t_end
//   k_fixed        whole fixed-grid trajectory on-chip: euler / midpoint / rk4(3/8) /
//                  Euler-Maruyama / leapfrog
//   k_reduce, k_gauss_logprob, k_philox, k_ffma_peak, k_pack_*   small helpers
//
// Reference semantics followed (paths relative to the reference checkout):
//   torchdiffeq dopri5 / fixed-grid drivers  -> restated in oracle/torchdiffeq/_solver.py
//   diffusion.py:258-279 (PF-ODE drift), :543-562 (Euler-Maruyama), :483-503 / :327-334 (trace)
//   flow.py:109-166, :553-652 ; symplectic.py:99-123, :186-201
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"
#include "ffb_engine_tc.cuh"
#include "ffb_control.cuh"
#include "ffb_rd.h"
#include "ffb_wide.h"

using namespace ffb;

#include "ffb_kernels_generic.cuh"

#include "ffb_kernels_rr.cuh"
#include "ffb_engine_rrt.cuh"

// =============================================================================================
// helpers
// =============================================================================================
// sums[q] = sum over tiles of partials[tile][q], in a fixed order (deterministic).  Thread i owns quantity q = i & 15
// of the tiles (i >> 4) + 64 k: a warp reads two whole 128-byte rows per load and keeps 8 loads in flight (the former
// one-warp-per-quantity loop was a chain of dependent L2 reads: 80 us for the 7 813 tiles of 1 M rows).
__device__ __forceinline__ void reduce_partials_block(const double* __restrict__ partials, int64_t ntiles, double* sh /*[1024]*/,
                                                      double* out /*[FFB_NPART], shared or global*/) {
  const int q = threadIdx.x & (FFB_NPART - 1);
  const int64_t r = threadIdx.x >> 4;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t t = r;
  for (; t + 7 * 64 < ntiles; t += 8 * 64) {
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] += partials[(t + 64 * u) * FFB_NPART + q];
  }
  for (; t < ntiles; t += 64) s[0] += partials[t * FFB_NPART + q];
  sh[threadIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (threadIdx.x < FFB_NPART) {
    double a = 0.0;
    for (int k = 0; k < 64; ++k) a += sh[k * FFB_NPART + threadIdx.x];
    out[threadIdx.x] = a;
  }
}
__global__ void __launch_bounds__(1024) k_reduce(const double* __restrict__ partials, int64_t ntiles, double* __restrict__ sums) {
  __shared__ double sh[1024];
  reduce_partials_block(partials, ntiles, sh, sums);
}

__global__ void k_gauss_logprob(const float* __restrict__ x, const float* __restrict__ add, float* __restrict__ out,
                                int64_t batch, int dim, float sigma, float log_norm) {
  // one warp per row: sum_d(-(x^2)/(2 sigma^2) - log sigma - 0.5 log 2pi)  (torch Normal.log_prob order)
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= batch) return;
  const float var2 = 2.0f * sigma * sigma;
  float s = 0.0f;
  for (int d = lane; d < dim; d += 32) {
    const float v = x[row * dim + d];
    s += -(v * v) / var2 - log_norm;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s + (add ? add[row] : 0.0f);
}

__global__ void k_philox(float* __restrict__ out, int64_t batch, int dim, uint64_t seed, uint64_t offset, int step,
                         int64_t row_offset) {
  const int ng = (dim + 3) >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * ng) return;
  const int64_t r = idx / ng;
  const int grp = (int)(idx - r * ng);
  const float4 z = philox_normal4(seed, offset, row_offset + r, step, grp);
  const float zz[4] = {z.x, z.y, z.z, z.w};
  for (int q = 0; q < 4; ++q)
    if (grp * 4 + q < dim) out[r * dim + grp * 4 + q] = zz[q];
}

__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters) {
  float2 acc[16];
  const float2 a = make_float2(1.0f + 1e-7f * threadIdx.x, 1.0f - 1e-7f * threadIdx.x);
  const float2 b = make_float2(1e-9f, -1e-9f);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2((float)i, (float)-i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;
}

// weight packing -------------------------------------------------------------------------------
// real column n of packed column p for a layer with padded width Np
__host__ __device__ inline int packed_to_real(int p, int Np) {
  const int C = Np / 16;
  if (C == 8) {
    const int g = p / 64, rem = p % 64, tx = rem / 4, jj = rem % 4;
    return (g * 4 + jj) * 16 + tx;
  }
  const int tx = p / C, j = p % C;
  return j * 16 + tx;
}
// dst[k][p] = W[n(p)][src_col(k)], zero outside; rows_map: k -> source column (or -1)
__global__ void k_pack_weight(const float* __restrict__ W, int in_features, int out_features, float* __restrict__ dst,
                              int K, int Np, int x_col, int x_dim, int c_col, int c_dim, int layer0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * Np) return;
  const int k = idx / Np, p = idx - k * Np;
  const int n = packed_to_real(p, Np);
  int col = -1;
  if (layer0) {
    if (k < x_dim) col = x_col + k;
    else if (k < x_dim + c_dim) col = c_col + (k - x_dim);
  } else if (k < in_features) {
    col = k;
  }
  dst[idx] = (n < out_features && col >= 0) ? W[(size_t)n * in_features + col] : 0.0f;
}
__global__ void k_pack_time(const float* __restrict__ W, int in_features, int out_features, float* __restrict__ dst,
                            int t_dim, int Np, int t_col) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= t_dim * Np) return;
  const int j = idx / Np, p = idx - j * Np;
  const int n = packed_to_real(p, Np);
  dst[idx] = (n < out_features) ? W[(size_t)n * in_features + t_col + j] : 0.0f;
}
__global__ void k_pack_bias(const float* __restrict__ b, int out_features, float* __restrict__ dst, int Np) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np) return;
  const int n = packed_to_real(p, Np);
  dst[p] = (n < out_features) ? b[n] : 0.0f;
}

// ---- tensor-core image: per chunk of 32 k-rows [W_hi | W_lo], each in the canonical no-swizzle
// K-major UMMA layout: element (n, k) at kc*(Np*16 B) + (n/8)*128 B + (n%8)*16 B + (k%4)*4 B, kc = k/4
__global__ void k_pack_weight_tc(const float* __restrict__ W, int in_features, int out_features,
                                 float* __restrict__ dst, int K, int Np, int x_col, int x_dim, int c_col, int c_dim,
                                 int layer0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * Np) return;
  const int k = idx / Np, n = idx - k * Np;
  int col = -1;
  if (layer0) {
    if (k < x_dim) col = x_col + k;
    else if (k < x_dim + c_dim) col = c_col + (k - x_dim);
  } else if (k < in_features) {
    col = k;
  }
  const float w = (n < out_features && col >= 0) ? W[(size_t)n * in_features + col] : 0.0f;
  uint32_t hi, lo;
  tf32_split_rn(w, hi, lo);
  const int c = k / KC, kk = k - c * KC, rows = min(KC, K - c * KC);
  const size_t base = (size_t)c * 2 * KC * Np;
  const size_t off = (size_t)(kk >> 2) * (Np * 4) + (size_t)(n >> 3) * 32 + (size_t)(n & 7) * 4 + (kk & 3);
  dst[base + off] = __uint_as_float(hi);
  dst[base + (size_t)rows * Np + off] = __uint_as_float(lo);
}
__global__ void k_pack_time_tc(const float* __restrict__ W, int in_features, int out_features, float* __restrict__ dst,
                               int t_dim, int Np, int t_col) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= t_dim * Np) return;
  const int j = idx / Np, n = idx - j * Np;
  dst[idx] = (n < out_features) ? W[(size_t)n * in_features + t_col + j] : 0.0f;
}
__global__ void k_pack_bias_tc(const float* __restrict__ b, int out_features, float* __restrict__ dst, int Np) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Np) return;
  dst[n] = (n < out_features) ? b[n] : 0.0f;
}

// ---- wide-engine image (ffb_engine_wide.cuh): [Np / CW][K][CW], CW = min(Np, 128); inside a chunk lane tx of a warp owns
// the packed columns tx*C .. tx*C + C-1 (C = CW / 32), which are the real columns j*32 + tx
__global__ void k_pack_weight_wide(const float* __restrict__ W, int in_features, int out_features, float* __restrict__ dst,
                                   int K, int Np, int CW, int x_col, int x_dim, int c_col, int c_dim, int layer0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * Np) return;
  const int nc = idx / (K * CW), rem = idx - nc * (K * CW);
  const int k = rem / CW, p = rem - k * CW;
  const int C = CW / 32, tx = p / C, j = p - tx * C;
  const int n = nc * CW + j * 32 + tx;
  int col = -1;
  if (layer0) {
    if (k < x_dim) col = x_col + k;
    else if (k < x_dim + c_dim) col = c_col + (k - x_dim);
  } else if (k < in_features) {
    col = k;
  }
  dst[idx] = (n < out_features && col >= 0) ? W[(size_t)n * in_features + col] : 0.0f;
}

// =============================================================================================
// device-side dopri5 controller (ffb_control.cuh): one thread, between two attempt kernels
// =============================================================================================
__global__ void k_dopri5_control(const __grid_constant__ ffb_dopri5_ctl_params p, const double* __restrict__ sums,
                                 ffb_dopri5_ctl* __restrict__ ctl, int after) {
  if (threadIdx.x == 0 && blockIdx.x == 0) ffbctl::control_turn(p, sums, *ctl, after);
}
// single-GPU shortcut: the tile reduction and the controller turn in one launch (sums is still written, for the caller)
__global__ void __launch_bounds__(1024) k_dopri5_reduce_control(const __grid_constant__ ffb_dopri5_ctl_params p,
        const double* __restrict__ partials, int64_t ntiles, double* __restrict__ sums, ffb_dopri5_ctl* __restrict__ ctl) {
  __shared__ double sh[1024];
  __shared__ double tot[FFB_NPART];
  if (ctl->done != FFB_CTL_RUNNING) {                  // uniform: nothing in flight to judge, only the turn is counted
    if (threadIdx.x == 0) ffbctl::notify_host(*ctl);
    return;
  }
  reduce_partials_block(partials, ntiles, sh, tot);
  __syncthreads();
  if (threadIdx.x < FFB_NPART) sums[threadIdx.x] = tot[threadIdx.x];
  if (threadIdx.x == 0) ffbctl::control_turn(p, tot, *ctl, 1);
}
__global__ void k_time_program(const __grid_constant__ ffb_time_program prog, const float* __restrict__ times, int n,
                               float sign, ffb_eval_scalars* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ffbctl::program_row(prog, times[i], sign, out + i);
}

// =============================================================================================
// host side: C ABI
// =============================================================================================
struct ffb_net {
  NetDev dev;   // FP32 FFMA2 engine image
  NetDev tc;    // tensor-core (3xTF32) engine image
  WideNet wide;             // wide-engine image (host copy of the descriptor; every network has one)
  WideNet* wide_dev;        // the same in device memory (what FieldDev::wide points to)
  bool narrow;              // true: dev / tc hold images (<= 8 layers, widths <= 128); false: one-layer summaries only
  std::vector<void*> allocs;
  int64_t flops;
};

// engine selection: 1 = tensor cores (default: dual-tile engine for fields without tangent rows, tangent-row engine
// for the log-likelihood paths), 3 = the same with the single-tile chunk-pipelined engine instead of the dual-tile one
// (FFB_ENGINE=rr), 4 = the same with the dual-tile engine for every dopri5 attempt it can hold, whatever the batch
// (FFB_ENGINE=rd; tests), 2 = the older whole-layer hand-off tile engine (FFB_ENGINE=tc_tile), 0 = FP32 FFMA2 (FFB_ENGINE=ffma),
// 5 = the wide engine for every field (FFB_ENGINE=wide; tests -- by default it only takes networks wider than 128 or deeper than 8)
static int g_engine = -1;
static int engine() {
  if (g_engine < 0) {
    const char* e = getenv("FFB_ENGINE");
    if (e && (!strcmp(e, "ffma") || !strcmp(e, "0"))) g_engine = 0;
    else if (e && (!strcmp(e, "tc_tile") || !strcmp(e, "2"))) g_engine = 2;
    else if (e && (!strcmp(e, "rr") || !strcmp(e, "3"))) g_engine = 3;
    else if (e && (!strcmp(e, "rd") || !strcmp(e, "4"))) g_engine = 4;
    else if (e && (!strcmp(e, "wide") || !strcmp(e, "5"))) g_engine = 5;
    else g_engine = 1;
  }
  return g_engine;
}
static bool chunk_engines() { return engine() == 1 || engine() == 3 || engine() == 4; }
// debug: timeline trace buffer (2 * 4096 int64) or NULL to disable
extern "C" int ffb_debug_trace(long long* buf) {
  int zero = 0;
  cudaMemcpyToSymbol(ffb::g_trace, &buf, sizeof(buf));
  cudaMemcpyToSymbol(ffb::g_trace_pos, &zero, sizeof(zero));
  return 0;
}
extern "C" int ffb_set_engine(int e) { g_engine = (e >= 2 && e <= 5) ? e : (e ? 1 : 0); return g_engine; }
extern "C" int ffb_get_engine(void) { return engine(); }

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return fail(FFB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
  } while (0)

// helpers shared with the other translation unit (csrc/ffb_staged.cu), declared in ffb_common.cuh
int ffb_fail(int code, const std::string& msg) { return fail(code, msg); }
void ffb_count_launches(int n) { g_launches += n; }

extern "C" int ffb_abi_version(void) { return FFB_ABI_VERSION; }
extern "C" const char* ffb_last_error(void) { return g_err.c_str(); }
extern "C" int64_t ffb_launch_count(void) { return g_launches.load(); }

extern "C" int ffb_device_info(int32_t* sm_count, int32_t* smem_optin, int32_t* cc_major, int32_t* cc_minor,
                               int32_t* clock_khz) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  int v = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem_optin) *smem_optin = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major) *cc_major = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
  if (cc_minor) *cc_minor = v;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev));
  if (clock_khz) *clock_khz = v;
  return FFB_OK;
}

static int pad_width(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }

extern "C" int ffb_net_create(const ffb_net_desc* d, void* stream_, ffb_net** out) {
  if (!d || !out) return fail(FFB_ERR_ARG, "ffb_net_create: null argument");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (d->n_layers < 1 || d->n_layers > FFB_MAX_LAYERS) return fail(FFB_ERR_ARG, "ffb_net_create: 1..16 Linear layers supported");
  if (d->t_dim < 0 || d->t_dim > FFB_MAX_TFEAT) return fail(FFB_ERR_ARG, "ffb_net_create: at most 32 time-feature columns");
  if (d->x_dim < 1 || d->x_dim + d->c_dim + d->t_dim != d->in_features)
    return fail(FFB_ERR_ARG, "ffb_net_create: x_dim + c_dim + t_dim must equal in_features");
  if (d->x_dim + d->c_dim > FFB_MAX_STATE) return fail(FFB_ERR_ARG, "ffb_net_create: state + conditional columns exceed 128");
  for (int l = 0; l < d->n_layers; ++l)
    if (d->widths[l] < 1 || d->widths[l] > FFB_MAX_WIDTH)
      return fail(FFB_ERR_ARG, "ffb_net_create: layer widths must be in 1..512");
  if (d->widths[d->n_layers - 1] > FFB_MAX_STATE) return fail(FFB_ERR_ARG, "ffb_net_create: the output layer is at most 128 wide");
  if (d->activation < FFB_ACT_SILU || d->activation > FFB_ACT_GELU) return fail(FFB_ERR_ARG, "ffb_net_create: unknown activation");
  ffb_net* net = new ffb_net();
  NetDev& nd = net->dev;
  NetDev& nt = net->tc;
  memset(&nd, 0, sizeof(nd));
  memset(&nt, 0, sizeof(nt));
  memset(&net->wide, 0, sizeof(net->wide));
  net->wide_dev = nullptr;
  net->narrow = d->n_layers <= NET_MAXL;
  for (int l = 0; l < d->n_layers; ++l) if (d->widths[l] > KMAX) net->narrow = false;
  net->flops = 0;
  int in_f = d->in_features;
  for (int l = 0; l < d->n_layers; ++l) { net->flops += 2LL * in_f * d->widths[l]; in_f = d->widths[l]; }
  auto cleanup = [&]() { for (void* p : net->allocs) cudaFree(p); delete net; };
  auto dalloc = [&](size_t nfloat) -> float* {
    float* p = nullptr;
    if (cudaMalloc(&p, sizeof(float) * (nfloat ? nfloat : 1)) != cudaSuccess) return nullptr;
    net->allocs.push_back(p);
    return p;
  };
  const int td = d->t_dim > 0 ? d->t_dim : 1;
  // ---- FP32 FFMA2 image and tensor-core image (networks the 128-wide engines hold) -------------------------------
  nd.t_dim = nt.t_dim = d->t_dim; nd.x_dim = nt.x_dim = d->x_dim; nd.c_dim = nt.c_dim = d->c_dim;
  nd.act = nt.act = d->activation;
  if (net->narrow) {
    nd.n_layers = nt.n_layers = d->n_layers;
    in_f = d->in_features;
    for (int l = 0; l < d->n_layers; ++l) {
      const int N = d->widths[l], Np = pad_width(N);
      const int K = (l == 0) ? ((d->x_dim + d->c_dim + 3) & ~3) : nd.Np[l - 1];
      nd.K[l] = K; nd.N[l] = N; nd.Np[l] = Np;
      float *w = dalloc((size_t)K * Np), *b = dalloc(Np);
      if (!w || !b) { cleanup(); return fail(FFB_ERR_CUDA, "ffb_net_create: cudaMalloc failed"); }
      k_pack_weight<<<(K * Np + 255) / 256, 256, 0, stream>>>(d->weight[l], in_f, N, w, K, Np, d->x_col, d->x_dim,
                                                             d->c_col, d->c_dim, l == 0);
      k_pack_bias<<<1, 128, 0, stream>>>(d->bias[l], N, b, Np);
      g_launches += 2;
      nd.W[l] = w; nd.b[l] = b;
      if (l == 0) {
        float* wt = dalloc((size_t)td * Np);
        if (!wt) { cleanup(); return fail(FFB_ERR_CUDA, "cudaMalloc"); }
        if (d->t_dim > 0) {
          k_pack_time<<<(d->t_dim * Np + 255) / 256, 256, 0, stream>>>(d->weight[0], in_f, N, wt, d->t_dim, Np, d->t_col);
          g_launches += 1;
        }
        nd.Wt = wt;
      }
      in_f = N;
    }
    in_f = d->in_features;
    for (int l = 0; l < d->n_layers; ++l) {
      const int N = d->widths[l], Np = (N + 31) & ~31;
      const int K = (l == 0) ? ((d->x_dim + d->c_dim + 7) & ~7) : nt.Np[l - 1];
      nt.K[l] = K; nt.N[l] = N; nt.Np[l] = Np;
      float *w = dalloc((size_t)2 * K * Np), *b = dalloc(Np);
      if (!w || !b) { cleanup(); return fail(FFB_ERR_CUDA, "ffb_net_create: cudaMalloc failed"); }
      k_pack_weight_tc<<<(K * Np + 255) / 256, 256, 0, stream>>>(d->weight[l], in_f, N, w, K, Np, d->x_col, d->x_dim,
                                                                d->c_col, d->c_dim, l == 0);
      k_pack_bias_tc<<<1, 128, 0, stream>>>(d->bias[l], N, b, Np);
      g_launches += 2;
      nt.W[l] = w; nt.b[l] = b;
      if (l == 0) {
        float* wt = dalloc((size_t)td * Np);
        if (!wt) { cleanup(); return fail(FFB_ERR_CUDA, "cudaMalloc"); }
        if (d->t_dim > 0) {
          k_pack_time_tc<<<(d->t_dim * Np + 255) / 256, 256, 0, stream>>>(d->weight[0], in_f, N, wt, d->t_dim, Np, d->t_col);
          g_launches += 1;
        }
        nt.Wt = wt;
      }
      in_f = N;
    }
  } else {
    // one-layer summary: what the host-side checks and the generic kernels read (dims, output width, activation)
    nd.n_layers = nt.n_layers = 1;
    nd.N[0] = nt.N[0] = d->widths[d->n_layers - 1];
    nd.Np[0] = nt.Np[0] = pad_width(nd.N[0]);
    nd.K[0] = nt.K[0] = (d->x_dim + d->c_dim + 7) & ~7;
  }
  // ---- wide-engine image (every network: FFB_ENGINE=wide runs any field on it) -------------------------------------
  WideNet& nw = net->wide;
  nw.n_layers = d->n_layers;
  nw.t_dim = d->t_dim; nw.x_dim = d->x_dim; nw.c_dim = d->c_dim; nw.act = d->activation;
  in_f = d->in_features;
  for (int l = 0; l < d->n_layers; ++l) {
    const int N = d->widths[l];
    const int Np = N <= 32 ? 32 : (N <= 64 ? 64 : ((N + 127) & ~127));
    const int CW = std::min(Np, 128);
    const int K = (l == 0) ? ((d->x_dim + d->c_dim + 3) & ~3) : nw.Np[l - 1];
    nw.K[l] = K; nw.N[l] = N; nw.Np[l] = Np;
    nw.max_k = std::max(nw.max_k, K);
    float *w = dalloc((size_t)K * Np), *b = dalloc(Np);
    if (!w || !b) { cleanup(); return fail(FFB_ERR_CUDA, "ffb_net_create: cudaMalloc failed"); }
    k_pack_weight_wide<<<(K * Np + 255) / 256, 256, 0, stream>>>(d->weight[l], in_f, N, w, K, Np, CW, d->x_col, d->x_dim,
                                                                d->c_col, d->c_dim, l == 0);
    k_pack_bias_tc<<<(Np + 127) / 128, 128, 0, stream>>>(d->bias[l], N, b, Np);
    g_launches += 2;
    nw.W[l] = w; nw.b[l] = b;
    if (l == 0) {
      float* wt = dalloc((size_t)td * Np);
      if (!wt) { cleanup(); return fail(FFB_ERR_CUDA, "cudaMalloc"); }
      if (d->t_dim > 0) {
        k_pack_time_tc<<<(d->t_dim * Np + 255) / 256, 256, 0, stream>>>(d->weight[0], in_f, N, wt, d->t_dim, Np, d->t_col);
        g_launches += 1;
      }
      nw.Wt = wt;
    }
    in_f = N;
  }
  void* wdev = nullptr;
  if (cudaMalloc(&wdev, sizeof(WideNet)) != cudaSuccess) { cleanup(); return fail(FFB_ERR_CUDA, "ffb_net_create: cudaMalloc failed"); }
  net->allocs.push_back(wdev);
  net->wide_dev = reinterpret_cast<WideNet*>(wdev);
  // net->wide lives as long as the handle, so the asynchronous copy may read it after this call returns
  if (cudaMemcpyAsync(wdev, &net->wide, sizeof(WideNet), cudaMemcpyHostToDevice, stream) != cudaSuccess) {
    cleanup();
    return fail(FFB_ERR_CUDA, "ffb_net_create: descriptor upload failed");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { cleanup(); return fail(FFB_ERR_CUDA, std::string("ffb_net_create: ") + cudaGetErrorString(e)); }
  *out = net;
  return FFB_OK;
}

extern "C" void ffb_net_destroy(ffb_net* net) {
  if (!net) return;
  for (void* p : net->allocs) cudaFree(p);
  delete net;
}

extern "C" int64_t ffb_net_flops(const ffb_net* net) { return net ? net->flops : 0; }

// ---- field validation / conversion -----------------------------------------------------------
static int tangents_of(const ffb_field* f) {
  return f->div_mode == FFB_DIV_EXACT ? f->net[0]->dev.x_dim : (f->div_mode == FFB_DIV_HUTCH ? 1 : 0);
}

static size_t field_smem(const FieldDev& fd, int T, int slots) {
  return engine() ? smem_layout_tc(fd.state_dim, fd.cond_dim, T, fd.div_mode == FFB_DIV_HUTCH, slots, fd.n_calls, field_tdim(fd), nullptr)
                  : smem_layout(fd.state_dim, fd.cond_dim, T, fd.div_mode == FFB_DIV_HUTCH, slots, nullptr);
}

static int make_field(const ffb_field* f, FieldDev* out) {
  if (!f) return fail(FFB_ERR_ARG, "null field");
  if (f->n_calls < 1 || f->n_calls > 2) return fail(FFB_ERR_ARG, "field: n_calls must be 1 or 2");
  if (f->state_dim < 1 || f->state_dim > FFB_MAX_STATE) return fail(FFB_ERR_ARG, "field: state_dim must be in 1..128");
  if (f->cond_dim < 0 || f->cond_dim > FFB_MAX_WIDTH) return fail(FFB_ERR_ARG, "field: bad cond_dim");
  memset(out, 0, sizeof(*out));
  out->n_calls = f->n_calls;
  for (int c = 0; c < f->n_calls; ++c) {
    if (!f->net[c]) return fail(FFB_ERR_ARG, "field: null net");
    const NetDev& nd = f->net[c]->dev;
    const int dout = nd.N[nd.n_layers - 1];
    if (nd.c_dim != f->cond_dim) return fail(FFB_ERR_ARG, "field: net conditional width differs from field cond_dim");
    if (f->in_off[c] < 0 || f->in_off[c] + nd.x_dim > f->state_dim) return fail(FFB_ERR_ARG, "field: input block outside the state");
    if (f->out_off[c] < 0 || f->out_off[c] + dout > f->state_dim) return fail(FFB_ERR_ARG, "field: output block outside the state");
    out->net[c] = engine() ? f->net[c]->tc : nd;
    out->wide[c] = f->net[c]->wide_dev;
    out->wide_maxk = std::max(out->wide_maxk, f->net[c]->wide.max_k);
    out->in_off[c] = f->in_off[c];
    out->out_off[c] = f->out_off[c];
    out->out_sign[c] = f->out_sign[c];
  }
  if (f->div_mode != FFB_DIV_NONE) {
    const NetDev& nd = f->net[0]->dev;
    if (f->n_calls != 1 || nd.x_dim != f->state_dim || nd.N[nd.n_layers - 1] != f->state_dim)
      return fail(FFB_ERR_ARG, "field: divergence needs a single network mapping the state to itself");
    if (f->div_mode == FFB_DIV_EXACT && 1 + nd.x_dim > TM)
      return fail(FFB_ERR_ARG, "field: exact trace supports at most 127 state columns");
  }
  out->state_dim = f->state_dim; out->cond_dim = f->cond_dim; out->kind = f->kind;
  out->use_sigma = f->use_sigma; out->has_drift = f->has_drift; out->div_mode = f->div_mode;
  // keep the Y0 / K1..K7 slots in shared memory whenever the tile still fits in one SM
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t with_slots = field_smem(*out, tangents_of(f), 1);
  out->slots_smem = (optin > 0 && with_slots <= (size_t)optin) ? 1 : 0;
  return FFB_OK;
}

// true when the field runs on the wide engine (ffb_engine_wide.cuh): a network wider than 128 or deeper than 8 layers,
// or every field under FFB_ENGINE=wide
static bool use_wide(const ffb_field* f) {
  if (engine() == 5) return true;
  for (int c = 0; c < f->n_calls; ++c) if (f->net[c] && !f->net[c]->narrow) return true;
  return false;
}

static int smem_optin();
// Tangent-row engine plan: samples per tile (<= what the 128 rows hold, reduced until the per-sample state fits
// shared memory) and the dynamic shared-memory size; 0 samples = the field cannot run on it.
static int rrt_plan(const ffb_field* f, int state_dim, int cond_dim, int tdim, int nslot, int nbeff, size_t* smem_out) {
  if (f->div_mode == FFB_DIV_NONE) return 0;
  int S = rrt_rowmap(tangents_of(f), -1, nullptr);
  for (; S >= 1; --S) {
    const size_t smem = smem_layout_rrt(state_dim, cond_dim, rrt_ld(S), f->div_mode == FFB_DIV_HUTCH, nslot, tdim, nbeff, nullptr);
    if (smem + 256 <= (size_t)smem_optin()) { if (smem_out) *smem_out = smem; return S; }   // 256: the kernels' static shared memory
  }
  return 0;
}
static int field_tdim_host(const ffb_field* f) {
  int t = f->net[0]->tc.t_dim;
  if (f->n_calls > 1 && f->net[1] && f->net[1]->tc.t_dim > t) t = f->net[1]->tc.t_dim;
  return t;
}
// Tiles of `batch` rows: the caller sizes `partials` with this, so it is the maximum over the tile shapes
// the field's kernels use (dense rows for the tile engine, quarter-packed samples for the tangent engine).
extern "C" int64_t ffb_num_tiles(const ffb_field* f, int64_t batch) {
  if (!f || !f->net[0]) return -1;
  const int S = TM / (1 + tangents_of(f));
  int64_t n = (batch + S - 1) / S;
  if (f->div_mode != FFB_DIV_NONE && !use_wide(f)) {
    const int td = field_tdim_host(f);
    const int configs[2][2] = {{NSLOT, 6}, {3, 1}};          // dopri5 attempt, single evaluation
    for (auto& c : configs) {
      const int St = rrt_plan(f, f->state_dim, f->cond_dim, td, c[0], c[1], nullptr);
      if (St > 0) n = std::max<int64_t>(n, (batch + St - 1) / St);
    }
  }
  return n;
}

static int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n > 0 ? n : 148;
}
int ffb_num_sms() { return num_sms(); }

extern "C" size_t ffb_scratch_bytes(const ffb_field* f) {
  if (!f) return 0;
  return std::max((size_t)num_sms() * NSLOT * f->state_dim * LDA * sizeof(float), rd_scratch_bytes(f->state_dim, f->cond_dim));
}

template <typename Kern, typename Args>
static int launch_tiles(Kern kern, int nthr, const char* name, const ffb_field* f, const FieldDev& fd, const Args& a,
                        int64_t batch, cudaStream_t stream) {
  const int T = tangents_of(f);
  for (int c = 0; c < fd.n_calls; ++c)
    if (fd.net[c].act != FFB_ACT_SILU)
      return fail(FFB_ERR_ARG, std::string(name) + ": non-SiLU activations run on the chunk-pipelined tensor-core engines only "
                  "(not on FFB_ENGINE=ffma / tc_tile, fixed grids with a divergence, or samples wider than a tile)");
  const size_t smem = field_smem(fd, T, fd.slots_smem);
  int dev = 0, optin = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ((int)smem > optin)
    return fail(FFB_ERR_ARG, std::string(name) + ": tile needs " + std::to_string(smem) + " B of shared memory, device allows " + std::to_string(optin));
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int Sd = TM / (1 + T);
  const int64_t ntiles = (batch + Sd - 1) / Sd;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  kern<<<grid, nthr, smem, stream>>>(fd, a, ntiles);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

// ---- row-resident tensor-core kernels (fields without tangent rows) -----------------------------
// FFB_ENGINE=tc_tile (or ffb_set_engine(2)) keeps the older whole-layer hand-off tile engine for A/B runs
static bool use_rr(const FieldDev& fd) { return chunk_engines() && fd.div_mode == FFB_DIV_NONE; }
// The dual-tile engine (two resident tiles per SM) takes the dopri5 attempts of fields without tangent rows when there
// are at least two tiles per SM (below that the single-tile engine spreads the tiles over more SMs) and the stage input
// fits shared memory beside its operand images.  Both engines issue the same MMAs in the same order: same bits.
static bool use_rd(const FieldDev& fd, int64_t batch) {
  if (fd.div_mode != FFB_DIV_NONE || !(engine() == 1 || engine() == 4) || !rd_dopri5_fits(fd)) return false;
  return engine() == 4 || (batch + TM - 1) / TM >= 2 * (int64_t)num_sms();
}
// true when a network of the field has a non-SiLU activation (selects the kernels with the run-time dispatch)
static bool gen_act(const FieldDev& fd) {
  for (int c = 0; c < fd.n_calls; ++c) if (fd.net[c].act != FFB_ACT_SILU) return true;
  return false;
}

static int smem_optin() {
  static int v = 0;
  if (!v) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); }
  return v;
}
// decide where the state slots live (shared memory when the tile still fits) and return the block size
static size_t rr_pick_smem(FieldDev* fd, int nslot, int nbeff) {
  const int td = field_tdim(*fd);
  const size_t with_slots = smem_layout_rr(fd->state_dim, fd->cond_dim, nslot, fd->n_calls, td, nbeff, nullptr);
  fd->slots_smem = (with_slots <= (size_t)smem_optin()) ? 1 : 0;
  return fd->slots_smem ? with_slots : smem_layout_rr(fd->state_dim, fd->cond_dim, 0, fd->n_calls, td, nbeff, nullptr);
}
template <typename Kern, typename Args>
static int launch_rr(Kern kern, size_t smem, const char* name, const FieldDev& fd, const Args& a, int64_t batch,
                     cudaStream_t stream) {
  if ((int)smem > smem_optin())
    return fail(FFB_ERR_ARG, std::string(name) + ": tile needs " + std::to_string(smem) + " B of shared memory, device allows " + std::to_string(smem_optin()));
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (batch + TM - 1) / TM;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  kern<<<grid, RR_NTHR, smem, stream>>>(fd, a, ntiles);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

// ---- tangent-row engine (log-likelihood paths) ------------------------------------------------------------
// returns 0 when the field cannot run on it; otherwise sets fd->rrt_cap and returns the shared-memory size
static size_t rrt_smem(const ffb_field* f, FieldDev* fd, int nslot, int nbeff) {
  if (!chunk_engines() || fd->div_mode == FFB_DIV_NONE) return 0;
  size_t smem = 0;
  const int S = rrt_plan(f, fd->state_dim, fd->cond_dim, field_tdim(*fd), nslot, nbeff, &smem);
  if (S <= 0) return 0;
  fd->rrt_cap = S;
  return smem;
}
template <typename Kern, typename Args>
static int launch_rrt(Kern kern, size_t smem, const char* name, const FieldDev& fd, const Args& a, int64_t batch,
                      cudaStream_t stream) {
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int S = fd.rrt_cap;
  const int64_t ntiles = (batch + S - 1) / S;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  kern<<<grid, RR_NTHR, smem, stream>>>(fd, a, ntiles);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  (void)name;
  return FFB_OK;
}

extern "C" int ffb_field_eval(const ffb_field* f, const ffb_eval_args* a, void* stream) {
  FieldDev fd;
  int rc = make_field(f, &fd);
  if (rc) return rc;
  if (!a || !a->y || !a->scratch) return fail(FFB_ERR_ARG, "ffb_field_eval: y and scratch are required");
  if (fd.cond_dim && !a->cond) return fail(FFB_ERR_ARG, "ffb_field_eval: cond is required");
  if (fd.div_mode == FFB_DIV_HUTCH && !a->probes) return fail(FFB_ERR_ARG, "ffb_field_eval: probes are required");
  if (a->norms && !a->partials) return fail(FFB_ERR_ARG, "ffb_field_eval: partials buffer required for norms");
  if (a->norms == 2 && (!a->fbase || (fd.div_mode != FFB_DIV_NONE && !a->dlpbase)))
    return fail(FFB_ERR_ARG, "ffb_field_eval: norms=2 needs fbase (and dlpbase with a divergence)");
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  if (a->jac && fd.div_mode != FFB_DIV_EXACT) return fail(FFB_ERR_ARG, "ffb_field_eval: jac needs div_mode FFB_DIV_EXACT");
  if (use_wide(f)) {
    if (a->jac) return fail(FFB_ERR_ARG, "ffb_field_eval: jac is written by the tangent-row tensor-core engine only (widths <= 128)");
    return wide_launch_field_eval(fd, *a, st_);
  }
  if (const size_t smt = rrt_smem(f, &fd, 3, 1))
    return gen_act(fd) ? launch_rrt(k_field_eval_rrt<true>, smt, "ffb_field_eval", fd, *a, a->batch, st_)
                       : launch_rrt(k_field_eval_rrt<false>, smt, "ffb_field_eval", fd, *a, a->batch, st_);
  if (a->jac) return fail(FFB_ERR_ARG, "ffb_field_eval: jac is written by the tangent-row tensor-core engine only");
  if (use_rr(fd)) {
    const size_t smem = rr_pick_smem(&fd, 3, 1);
    if (gen_act(fd)) {
      if (fd.slots_smem) return launch_rr(k_field_eval_rr<true, true>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
      return launch_rr(k_field_eval_rr<false, true>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
    }
    if (fd.slots_smem) return launch_rr(k_field_eval_rr<true, false>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
    return launch_rr(k_field_eval_rr<false, false>, smem, "ffb_field_eval", fd, *a, a->batch, st_);
  }
  if (engine()) {
    if (fd.slots_smem) return launch_tiles(k_field_eval<EngineTC, true>, EngineTC::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
    return launch_tiles(k_field_eval<EngineTC, false>, EngineTC::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
  }
  if (fd.slots_smem) return launch_tiles(k_field_eval<EngineFFMA, true>, EngineFFMA::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
  return launch_tiles(k_field_eval<EngineFFMA, false>, EngineFFMA::NTHR, "ffb_field_eval", f, fd, *a, a->batch, st_);
}

extern "C" int ffb_dopri5_attempt(const ffb_field* f, const ffb_dopri5_args* a, void* stream) {
  FieldDev fd;
  int rc = make_field(f, &fd);
  if (rc) return rc;
  if (!a || !a->y0 || !a->f0 || !a->y1 || !a->f1 || !a->partials || !a->scratch)
    return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: y0, f0, y1, f1, partials, scratch are required");
  if (fd.div_mode != FFB_DIV_NONE && (!a->lp0 || !a->dlp0 || !a->lp1 || !a->dlp1))
    return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: log-det buffers are required with a divergence");
  if ((a->final || a->ctl) && (!a->y_out || (fd.div_mode != FFB_DIV_NONE && !a->lp_out)))
    return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: final step needs y_out (and lp_out)");
  if (fd.cond_dim && !a->cond) return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: cond is required");
  if (fd.div_mode == FFB_DIV_HUTCH && !a->probes) return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: probes are required");
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  const bool dyn = a->ctl != nullptr;
  if (use_wide(f)) {
    if (dyn) return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: args->ctl is not available on the wide engine (ffb_dopri5_ctl_supported)");
    return wide_launch_dopri5(fd, *a, st_);
  }
  if (const size_t smt = rrt_smem(f, &fd, NSLOT, 6)) {
    if (dyn) return gen_act(fd) ? launch_rrt(k_dopri5_rrt<true, true>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_)
                                : launch_rrt(k_dopri5_rrt<false, true>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_);
    return gen_act(fd) ? launch_rrt(k_dopri5_rrt<true, false>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_)
                       : launch_rrt(k_dopri5_rrt<false, false>, smt, "ffb_dopri5_attempt", fd, *a, a->batch, st_);
  }
  if (use_rd(fd, a->batch)) return rd_launch_dopri5(fd, *a, st_);
  if (use_rr(fd)) {
    const size_t smem = rr_pick_smem(&fd, NSLOT, 6);
#define FFB_RR_DOPRI5(SS_, GEN_)                                                                                      \
  (dyn ? launch_rr(k_dopri5_rr<SS_, GEN_, true>, smem, "ffb_dopri5_attempt", fd, *a, a->batch, st_)                    \
       : launch_rr(k_dopri5_rr<SS_, GEN_, false>, smem, "ffb_dopri5_attempt", fd, *a, a->batch, st_))
    if (gen_act(fd)) return fd.slots_smem ? FFB_RR_DOPRI5(true, true) : FFB_RR_DOPRI5(false, true);
    return fd.slots_smem ? FFB_RR_DOPRI5(true, false) : FFB_RR_DOPRI5(false, false);
#undef FFB_RR_DOPRI5
  }
  if (dyn) return fail(FFB_ERR_ARG, "ffb_dopri5_attempt: args->ctl needs the chunk-pipelined tensor-core engines (ffb_dopri5_ctl_supported)");
  if (engine()) {
    if (fd.slots_smem) return launch_tiles(k_dopri5<EngineTC, true>, EngineTC::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
    return launch_tiles(k_dopri5<EngineTC, false>, EngineTC::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
  }
  if (fd.slots_smem) return launch_tiles(k_dopri5<EngineFFMA, true>, EngineFFMA::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
  return launch_tiles(k_dopri5<EngineFFMA, false>, EngineFFMA::NTHR, "ffb_dopri5_attempt", f, fd, *a, a->batch, st_);
}

extern "C" int ffb_integrate_fixed(const ffb_field* f, const ffb_fixed_args* a, void* stream) {
  FieldDev fd;
  int rc = make_field(f, &fd);
  if (rc) return rc;
  if (!a || !a->x0 || !a->x_out || !a->step_table || !a->ev_table || !a->scratch || !a->status)
    return fail(FFB_ERR_ARG, "ffb_integrate_fixed: x0, x_out, step_table, ev_table, scratch, status are required");
  if (a->method < FFB_M_EULER || a->method > FFB_M_LEAPFROG) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: unknown method");
  if (a->method == FFB_M_LEAPFROG && fd.n_calls != 2) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: leapfrog needs a 2-network (q,p) field");
  if ((a->method == FFB_M_EM || a->method == FFB_M_LEAPFROG) && fd.div_mode != FFB_DIV_NONE)
    return fail(FFB_ERR_ARG, "ffb_integrate_fixed: no divergence with this method");
  if (fd.cond_dim && !a->cond) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: cond is required");
  if (fd.div_mode == FFB_DIV_HUTCH && !a->probes) return fail(FFB_ERR_ARG, "ffb_integrate_fixed: probes are required");
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  if (use_wide(f)) return wide_launch_fixed(fd, *a, st_);
  if (const size_t smt = rrt_smem(f, &fd, rr_fixed_slots(a->method), 1))
    return gen_act(fd) ? launch_rrt(k_fixed_rrt<true>, smt, "ffb_integrate_fixed", fd, *a, a->batch, st_)
                       : launch_rrt(k_fixed_rrt<false>, smt, "ffb_integrate_fixed", fd, *a, a->batch, st_);
  if (use_rr(fd)) {
    const size_t smem = rr_pick_smem(&fd, rr_fixed_slots(a->method), 8);
    if (gen_act(fd)) {
      if (fd.slots_smem) return launch_rr(k_fixed_rr<true, true>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
      return launch_rr(k_fixed_rr<false, true>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
    }
    if (fd.slots_smem) return launch_rr(k_fixed_rr<true, false>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
    return launch_rr(k_fixed_rr<false, false>, smem, "ffb_integrate_fixed", fd, *a, a->batch, st_);
  }
  if (engine()) {
    if (fd.slots_smem) return launch_tiles(k_fixed<EngineTC, true>, EngineTC::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
    return launch_tiles(k_fixed<EngineTC, false>, EngineTC::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
  }
  if (fd.slots_smem) return launch_tiles(k_fixed<EngineFFMA, true>, EngineFFMA::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
  return launch_tiles(k_fixed<EngineFFMA, false>, EngineFFMA::NTHR, "ffb_integrate_fixed", f, fd, *a, a->batch, st_);
}

extern "C" int ffb_dopri5_ctl_supported(const ffb_field* f) {
  FieldDev fd;
  if (make_field(f, &fd)) return 0;
  if (use_wide(f)) return 0;
  if (rrt_smem(f, &fd, NSLOT, 6)) return 1;
  return use_rr(fd) ? 1 : 0;
}

extern "C" int ffb_dopri5_control(const ffb_dopri5_ctl_params* p, double* sums, const double* partials, int64_t n_tiles,
                                  ffb_dopri5_ctl* ctl, int32_t after_attempt, void* stream) {
  if (!p || !ctl || (after_attempt && !sums)) return fail(FFB_ERR_ARG, "ffb_dopri5_control: null argument");
  if (p->n_grid < 0 || p->n_grid > FFB_CTL_MAX_GRID) return fail(FFB_ERR_ARG, "ffb_dopri5_control: at most 16 step_t points");
  if (p->prog.n_freq < 0 || p->prog.n_freq > FFB_MAX_FREQ) return fail(FFB_ERR_ARG, "ffb_dopri5_control: bad n_freq");
  if (after_attempt && partials)
    k_dopri5_reduce_control<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, partials, n_tiles, sums, ctl);
  else
    k_dopri5_control<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, sums, ctl, after_attempt);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_dopri5_control_host(const ffb_dopri5_ctl_params* p, const double* sums, ffb_dopri5_ctl* ctl,
                                       int32_t after_attempt) {
  if (!p || !ctl || (after_attempt && !sums)) return fail(FFB_ERR_ARG, "ffb_dopri5_control_host: null argument");
  if (p->n_grid < 0 || p->n_grid > FFB_CTL_MAX_GRID) return fail(FFB_ERR_ARG, "ffb_dopri5_control_host: at most 16 step_t points");
  if (p->prog.n_freq < 0 || p->prog.n_freq > FFB_MAX_FREQ) return fail(FFB_ERR_ARG, "ffb_dopri5_control_host: bad n_freq");
  ffbctl::control_turn(*p, sums, *ctl, after_attempt);
  return FFB_OK;
}

extern "C" int ffb_time_program_rows(const ffb_time_program* prog, const float* times, int32_t n, float sign,
                                     ffb_eval_scalars* out, int32_t on_device) {
  if (!prog || !times || !out || n < 0) return fail(FFB_ERR_ARG, "ffb_time_program_rows: bad argument");
  if (prog->n_freq < 0 || prog->n_freq > FFB_MAX_FREQ) return fail(FFB_ERR_ARG, "ffb_time_program_rows: bad n_freq");
  if (n == 0) return FFB_OK;
  if (!on_device) {
    for (int i = 0; i < n; ++i) ffbctl::program_row(*prog, times[i], sign, out + i);
    return FFB_OK;
  }
  float* dt = nullptr; ffb_eval_scalars* dout = nullptr;
  CUDA_TRY(cudaMalloc(&dt, sizeof(float) * n));
  if (cudaMalloc(&dout, sizeof(ffb_eval_scalars) * n) != cudaSuccess) { cudaFree(dt); return fail(FFB_ERR_CUDA, "cudaMalloc"); }
  cudaMemcpy(dt, times, sizeof(float) * n, cudaMemcpyHostToDevice);
  k_time_program<<<(n + 127) / 128, 128>>>(*prog, dt, n, sign, dout);
  g_launches += 1;
  cudaError_t e = cudaMemcpy(out, dout, sizeof(ffb_eval_scalars) * n, cudaMemcpyDeviceToHost);
  cudaFree(dt); cudaFree(dout);
  if (e != cudaSuccess) return fail(FFB_ERR_CUDA, std::string("ffb_time_program_rows: ") + cudaGetErrorString(e));
  return FFB_OK;
}

extern "C" int ffb_reduce_partials(const double* partials, int64_t n_tiles, double* sums, void* stream) {
  if (!partials || !sums) return fail(FFB_ERR_ARG, "ffb_reduce_partials: null argument");
  k_reduce<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partials, n_tiles, sums);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_gaussian_logprob(const float* x, const float* add, float* out, int64_t batch, int32_t dim,
                                    float sigma, void* stream) {
  if (!x || !out || dim < 1) return fail(FFB_ERR_ARG, "ffb_gaussian_logprob: bad argument");
  if (batch == 0) return FFB_OK;
  const float log_norm = logf(sigma) + 0.918938533204672742f;   // log sigma + log sqrt(2 pi)
  const int wpb = 8;
  k_gauss_logprob<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, add, out, batch, dim, sigma, log_norm);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_philox_normal(float* out, int64_t batch, int32_t dim, uint64_t seed, uint64_t offset, int32_t step,
                                 int64_t row_offset, void* stream) {
  if (!out || dim < 1) return fail(FFB_ERR_ARG, "ffb_philox_normal: bad argument");
  const int64_t n = batch * ((dim + 3) / 4);
  if (n == 0) return FFB_OK;
  k_philox<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, batch, dim, seed,
                                                                                             offset, step, row_offset);
  g_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_ffma_peak(int32_t iters, float* tflops, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  float* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 4));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int grid = num_sms() * 8;
  k_ffma_peak<<<grid, 256, 0, stream>>>(d, iters);   // warm-up
  CUDA_TRY(cudaEventRecord(e0, stream));
  k_ffma_peak<<<grid, 256, 0, stream>>>(d, iters);
  CUDA_TRY(cudaEventRecord(e1, stream));
  g_launches += 2;
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  const double flop = (double)grid * 256.0 * (double)iters * 16.0 * 4.0;   // 16 FFMA2 = 32 FMA = 64 FLOP
  if (tflops) *tflops = (float)(flop / (ms * 1e-3) / 1e12);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return FFB_OK;
}


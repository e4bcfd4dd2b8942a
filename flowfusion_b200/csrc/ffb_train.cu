// ffb_train.cu -- fused training step of one MLP (the training-side losses,
// diffusion.py:1369-1463 denoising / likelihood-weighted score matching, flow.py:191-256 and :679-747 flow matching).
//
// Every one of the reference's losses is, for per-row scalars alpha_b and a per-element offset beta_{b,d} that do not
// depend on the weights,
//        loss = scale * sum_{b,d} ( alpha_b * net(X)_{b,d} + beta_{b,d} )^2
// (DSM: alpha = 1 or sigma_b, beta = z;  likelihood weighting: alpha = g_b/sigma_b or g_b, beta = (g_b/sigma_b) z;
//  flow matching: alpha = 1, beta = x0 - xT), so one entry point serves them all:  ffb_train_step  returns the loss and
// d loss / d (every weight and bias) [and optionally d loss / d X] in four launches, whatever the depth:
//   k_train_pack     weights -> the forward image (k-major, per 128-column chunk) and the transposed image for the
//                    backward sweep, both in the wide engine's layout (the weights change every optimiser step);
//   k_train_fwdbwd   persistent CTAs walk the batch in passes of 32 rows with the wide engine's contraction
//                    (ffb_engine_wide.cuh: a warp owns 4 rows end to end, weights stream through the cp.async.bulk ring,
//                    packed FFMA2): forward keeping the pre-activations z_l in shared memory, the loss residual, then the
//                    backward sweep delta_{l-1} = (W_l^T delta_l) * act'(z_{l-1}); h_l and delta_l go to global memory
//                    (L2-resident at training batch sizes) for the weight gradients, FP64 loss partial per pass;
//   k_train_dw       dW_l = delta_l^T h_{l-1}, db_l = sum_b delta_l: 64 x 64 patches (4 x 4 per thread, FFMA2), the batch split over grid.y;
//   k_train_reduce   sums the split partials in a fixed order into the caller's gradient tensors (torch layout) and the
//                    pass partials into the loss: deterministic, no atomics.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <string>

#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"
#include "ffb_engine_wide.cuh"

namespace ffb {

#define TR_CUDA_TRY(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ffb_fail(FFB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

constexpr int TR_MAXL = FFB_MAX_LAYERS;

// Everything the kernels need to know about one training step (passed by value).
struct TrainPlan {
  int n_layers, in_features, act, want_dx;
  int fwd_only;         // forward pass only (out = net(X)): no pre-activations kept, no backward sweep, no weight gradients
  int64_t batch;
  int N[TR_MAXL];       // real output width of layer l
  int Kin[TR_MAXL];     // real input width of layer l
  int Np[TR_MAXL];      // padded output width (32, 64 or a multiple of 128)
  int Kp[TR_MAXL];      // padded input width: layer 0: in_features rounded up to 4; else Np[l-1]
  int KB0;              // padded width of dX (want_dx): in_features padded like an output
  const float* W[TR_MAXL];    // caller's weights, [N][Kin] row-major
  const float* B[TR_MAXL];    // caller's biases
  float* Wf[TR_MAXL];         // forward image  [Np/CW][Kp][CW]
  float* Wb[TR_MAXL];         // backward image [Kp'/CW'][Np][CW'] (l >= 1; l = 0 only with want_dx, Kp' = KB0)
  float* bp[TR_MAXL];         // padded bias [Np]
  float* hin[TR_MAXL];        // l >= 1: input of layer l = activation of layer l-1, (B, Np[l-1]) row-major
  float* dl[TR_MAXL];         // delta at the output of layer l, (B, Np[l]) row-major
  const float* x_in;          // (B, in_features)
  const float* alpha;         // (B,) or NULL
  const float* beta;          // (B, N[L-1])
  const float* cot;           // (B, N[L-1]) or NULL: VJP mode, delta_L = scale * cot and loss = scale * sum(cot * net)
  float* out;                 // (B, N[L-1]) or NULL: the network output
  float scale;
  float* grad_x;              // (B, in_features) or NULL
  double* pass_loss;          // [npass]
  float* dw_part;             // [splits][n_param]
  int64_t n_param;            // sum_l N*Kin + N
  int64_t off_w[TR_MAXL], off_b[TR_MAXL];   // offsets of layer l's dW / db inside one split's block
  int splits;
  int zoff[TR_MAXL];          // float offset of z_l inside the z region of shared memory (hidden layers)
  int maxk;                   // widest operand
};

static inline int tr_padw(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : ((n + 127) & ~127)); }

// ---------------------------------------------------------------------------------------------------------------
// pack: one launch for every image of every layer
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_train_pack(const __grid_constant__ TrainPlan p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int l = 0; l < p.n_layers; ++l) {
    const int N = p.N[l], Kin = p.Kin[l], Np = p.Np[l], Kp = p.Kp[l];
    const float* __restrict__ W = p.W[l];
    {   // forward image: contraction over k (inputs), columns n
      const int CW = min(Np, WD_CW), C = CW / 32;
      float* __restrict__ dst = p.Wf[l];
      for (int64_t idx = t0; idx < (int64_t)Kp * Np; idx += stride) {
        const int nc = (int)(idx / ((int64_t)Kp * CW)), rem = (int)(idx - (int64_t)nc * Kp * CW);
        const int k = rem / CW, q = rem - k * CW, tx = q / C, j = q - tx * C;
        const int n = nc * CW + j * 32 + tx;
        dst[idx] = (n < N && k < Kin) ? W[(size_t)n * Kin + k] : 0.0f;
      }
    }
    if (l >= 1 || p.want_dx) {   // backward image: contraction over n (outputs), columns k
      const int KB = (l == 0) ? p.KB0 : Kp;
      const int CW = min(KB, WD_CW), C = CW / 32;
      float* __restrict__ dst = p.Wb[l];
      for (int64_t idx = t0; idx < (int64_t)Np * KB; idx += stride) {
        const int kc = (int)(idx / ((int64_t)Np * CW)), rem = (int)(idx - (int64_t)kc * Np * CW);
        const int n = rem / CW, q = rem - n * CW, tx = q / C, j = q - tx * C;
        const int k = kc * CW + j * 32 + tx;
        dst[idx] = (n < N && k < Kin) ? W[(size_t)n * Kin + k] : 0.0f;
      }
    }
    for (int64_t n = t0; n < Np; n += stride) p.bp[l][n] = (n < N) ? p.B[l][n] : 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward + loss + backward sweep over passes of 32 rows
// ---------------------------------------------------------------------------------------------------------------
// shared memory: [z region | operand ping-pong 2 x maxk x WD_RS | ring | barriers | FP64 reduction scratch]
__host__ __device__ inline size_t train_smem(int zfloats, int maxk, size_t* off /*[5]*/) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  size_t v[5];
  v[0] = take(sizeof(float) * (zfloats > 0 ? zfloats : 1));
  v[1] = take(sizeof(float) * 2 * maxk * WD_RS);
  v[2] = take(sizeof(float) * WD_NSTAGE * WD_STAGE_FLOATS);
  v[3] = take(sizeof(uint64_t) * 2 * WD_NSTAGE);
  v[4] = take(sizeof(double) * 8);
  if (off) for (int i = 0; i < 5; ++i) off[i] = v[i];
  return o;
}

// stream one packed image ([cols / CW][K][CW]) through the ring (producer warp)
__device__ __forceinline__ void train_produce(CtxW& cx, const float* img, int K, int cols) {
  const int CW = min(cols, WD_CW);
  for (int nc = 0; nc < cols / CW; ++nc) {
    const float* base = img + (size_t)nc * K * CW;
    for (int k0 = 0; k0 < K; k0 += WD_KC) {
      const int rows = min(WD_KC, K - k0);
      if (cx.lane == 0) {
        mbar_wait(&cx.empty()[cx.stage], cx.phase);
        const uint32_t bytes = (uint32_t)(rows * CW) * sizeof(float);
        mbar_expect_tx(&cx.full()[cx.stage], bytes);
        bulk_g2s(cx.ring() + cx.stage * WD_STAGE_FLOATS, base + (size_t)k0 * CW, bytes, &cx.full()[cx.stage]);
      }
      EngineWide::advance(cx);
    }
  }
}

// One chunk of C x 32 output columns for this warp's 4 rows, then `epi(n, v[4])` per owned column n.
template <int C, class EPI>
__device__ __forceinline__ void train_chunk(CtxW& cx, const float* opin, int K, int nc, EPI&& epi) {
  float2 acc[2][C];
  EngineWide::gemm_chunk<C>(cx, opin, K, acc);
#pragma unroll
  for (int j = 0; j < C; ++j) {
    float v[4] = {acc[0][j].x, acc[0][j].y, acc[1][j].x, acc[1][j].y};
    epi(nc * (32 * C) + j * 32 + cx.lane, v);
  }
}
template <class EPI>
__device__ __forceinline__ void train_layer(CtxW& cx, const float* opin, int K, int cols, EPI&& epi) {
  const int CW = min(cols, WD_CW);
  for (int nc = 0; nc < cols / CW; ++nc) {
    if (CW == 128) train_chunk<4>(cx, opin, K, nc, epi);
    else if (CW == 64) train_chunk<2>(cx, opin, K, nc, epi);
    else train_chunk<1>(cx, opin, K, nc, epi);
  }
  __syncwarp();
}

template <int ACT>
__device__ __forceinline__ void train_passes(CtxW& cx, const TrainPlan& p, float* zreg, float* op, int64_t npass, double& loss) {
  const int L = p.n_layers;
  const int r0w = cx.warp * 4;
  for (int64_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const int64_t row0 = pass * WD_R + r0w;          // first of this warp's 4 rows
    float* A = op;
    float* Bf = op + (size_t)p.maxk * WD_RS;
    // layer-0 operand: this warp's rows of X, transposed to k-major
    for (int k = cx.lane; k < p.Kp[0]; k += 32) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (k < p.in_features && row0 + i < p.batch) ? p.x_in[(row0 + i) * p.in_features + k] : 0.0f;
      *reinterpret_cast<float4*>(A + (size_t)k * WD_RS + r0w) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncwarp();
    // ---- forward ------------------------------------------------------------------------------------------------
    for (int l = 0; l < L; ++l) {
      const int Np = p.Np[l];
      const float* bias = p.bp[l];
      if (l < L - 1) {
        float* z = zreg + p.zoff[l];
        float* hg = p.hin[l + 1];
        train_layer(cx, A, p.Kp[l], Np, [&](int n, float (&v)[4]) {
          const float bj = bias[n];
          float h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[i] += bj; h[i] = act_fwd<ACT>(v[i]); }
          *reinterpret_cast<float4*>(Bf + (size_t)n * WD_RS + r0w) = make_float4(h[0], h[1], h[2], h[3]);
          if (!p.fwd_only) {
            *reinterpret_cast<float4*>(z + (size_t)n * WD_RS + r0w) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) if (row0 + i < p.batch) hg[(row0 + i) * Np + n] = h[i];
          }
        });
      } else {
        const int Nout = p.N[l];
        float* dg = p.dl[l];
        train_layer(cx, A, p.Kp[l], Np, [&](int n, float (&v)[4]) {
          const float bj = bias[n];
          float d[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            d[i] = 0.0f;
            if (n < Nout && row0 + i < p.batch) {
              const float o = v[i] + bj;
              if (p.out) p.out[(row0 + i) * Nout + n] = o;
              if (p.fwd_only) continue;
              if (p.cot) {
                const float ct = p.cot[(row0 + i) * Nout + n];
                loss += (double)ct * (double)o;
                d[i] = p.scale * ct;
              } else {
                const float al = p.alpha ? p.alpha[row0 + i] : 1.0f;
                const float r = fmaf(al, o, p.beta[(row0 + i) * Nout + n]);
                loss += (double)r * (double)r;
                d[i] = 2.0f * p.scale * al * r;
              }
            }
          }
          if (p.fwd_only) return;
          *reinterpret_cast<float4*>(Bf + (size_t)n * WD_RS + r0w) = make_float4(d[0], d[1], d[2], d[3]);
#pragma unroll
          for (int i = 0; i < 4; ++i) if (row0 + i < p.batch) dg[(row0 + i) * Np + n] = d[i];
        });
      }
      float* t = A; A = Bf; Bf = t;
    }
    if (p.fwd_only) continue;
    // ---- backward sweep: A holds delta_{L-1} (k-major over the output columns) --------------------------------------
    for (int l = L - 1; l >= 1; --l) {
      const float* z = zreg + p.zoff[l - 1];
      float* dg = p.dl[l - 1];
      const int Kp = p.Kp[l];                          // = Np[l-1]
      train_layer(cx, A, p.Np[l], Kp, [&](int k, float (&v)[4]) {
        const float4 z4 = *reinterpret_cast<const float4*>(z + (size_t)k * WD_RS + r0w);
        const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
        float d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { float a, g; act_fwd_grad<ACT>(zz[i], a, g); d[i] = v[i] * g; }
        *reinterpret_cast<float4*>(Bf + (size_t)k * WD_RS + r0w) = make_float4(d[0], d[1], d[2], d[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) if (row0 + i < p.batch) dg[(row0 + i) * Kp + k] = d[i];
      });
      float* t = A; A = Bf; Bf = t;
    }
    if (p.want_dx) {
      train_layer(cx, A, p.Np[0], p.KB0, [&](int k, float (&v)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k < p.in_features && row0 + i < p.batch) p.grad_x[(row0 + i) * p.in_features + k] = v[i];
      });
    }
  }
}

__global__ void __launch_bounds__(NTHR, 1) k_train_fwdbwd(const __grid_constant__ TrainPlan p, const int64_t npass, const int zfloats) {
  CtxW cx;
  size_t off[5];
  train_smem(zfloats, p.maxk, off);
  float* zreg = reinterpret_cast<float*>(smem_base() + off[0]);
  float* op = reinterpret_cast<float*>(smem_base() + off[1]);
  cx.o_ring = (uint32_t)off[2]; cx.o_bar = (uint32_t)off[3];
  double* red = reinterpret_cast<double*>(smem_base() + off[4]);
  cx.tid = threadIdx.x; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
  cx.producer = (cx.warp == NCOMP / 32);
  cx.stage = 0; cx.phase = cx.producer ? 1u : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WD_NSTAGE; ++s) { mbar_init(&cx.full()[s], 1); mbar_init(&cx.empty()[s], NCOMP / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cx.producer) {
    for (int64_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
      for (int l = 0; l < p.n_layers; ++l) train_produce(cx, p.Wf[l], p.Kp[l], p.Np[l]);
      if (p.fwd_only) continue;
      for (int l = p.n_layers - 1; l >= 1; --l) train_produce(cx, p.Wb[l], p.Np[l], p.Kp[l]);
      if (p.want_dx) train_produce(cx, p.Wb[0], p.Np[0], p.KB0);
    }
    return;
  }
  double loss = 0.0;
  FFB_ACT_DISPATCH(p.act, (train_passes<ACT>(cx, p, zreg, op, npass, loss)));
  // one FP64 partial per pass would need a reduction per pass; a CTA's passes are summed here and written once per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_down_sync(0xffffffffu, loss, o);
  if (cx.lane == 0) red[cx.warp] = loss;
  bar_compute();
  if (cx.tid == 0) {
    double s = 0.0;
    for (int w = 0; w < NCOMP / 32; ++w) s += red[w];
    p.pass_loss[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradients: 64 x 64 patches of dW_l (+ db_l from the patches with k0 = 0), batch split over grid.y.
// A thread owns 4 x 4 outputs: per batch row one LDS.128 of delta (broadcast to the 16 threads that share the rows of
// the patch), one LDS.128 of h, 8 packed FFMA2.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DW_P = 64;             // patch edge
constexpr int DW_LD = DW_P + 4;      // shared-memory row stride (floats): float4-aligned, conflict-free for the two access patterns
__host__ __device__ inline int dw_patches(const TrainPlan& p) {
  int n = 0;
  for (int l = 0; l < p.n_layers; ++l) n += ((p.N[l] + DW_P - 1) / DW_P) * ((p.Kin[l] + DW_P - 1) / DW_P);
  return n;
}
__global__ void __launch_bounds__(256) k_train_dw(const __grid_constant__ TrainPlan p) {
  __shared__ __align__(16) float sd[32][DW_LD];
  __shared__ __align__(16) float sh[32][DW_LD];
  // which layer / patch is this block?
  int l = 0, pid = blockIdx.x;
  for (; l < p.n_layers; ++l) {
    const int np = ((p.N[l] + DW_P - 1) / DW_P) * ((p.Kin[l] + DW_P - 1) / DW_P);
    if (pid < np) break;
    pid -= np;
  }
  if (l >= p.n_layers) return;
  const int kb = (p.Kin[l] + DW_P - 1) / DW_P;
  const int n0 = (pid / kb) * DW_P, k0 = (pid % kb) * DW_P;
  const int N = p.N[l], Kin = p.Kin[l];
  const float* __restrict__ dl = p.dl[l];
  const int ldd = p.Np[l];
  const float* __restrict__ hin = (l == 0) ? p.x_in : p.hin[l];
  const int ldh = (l == 0) ? p.in_features : p.Np[l - 1];
  const int64_t per = (p.batch + p.splits - 1) / p.splits;
  const int64_t b0 = (int64_t)blockIdx.y * per, b1 = min(p.batch, b0 + per);
  const int tn = threadIdx.x >> 4, tk = threadIdx.x & 15;     // outputs n0 + 4 tn + {0..3}, k0 + 4 tk + {0..3}
  float2 acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = make_float2(0.f, 0.f); acc[i][1] = make_float2(0.f, 0.f); }
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
  const int lr = threadIdx.x >> 6, lc = threadIdx.x & 63;
  for (int64_t bb = b0; bb < b1; bb += 32) {
#pragma unroll
    for (int r = lr; r < 32; r += 4) {
      const int64_t b = bb + r;
      sd[r][lc] = (b < b1 && n0 + lc < ldd) ? dl[b * ldd + n0 + lc] : 0.0f;
      sh[r][lc] = (b < b1 && k0 + lc < Kin) ? hin[b * ldh + k0 + lc] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float4 d4 = *reinterpret_cast<const float4*>(&sd[r][4 * tn]);
      const float4 h4 = *reinterpret_cast<const float4*>(&sh[r][4 * tk]);
      const float2 h01 = make_float2(h4.x, h4.y), h23 = make_float2(h4.z, h4.w);
      const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 di = make_float2(dd[i], dd[i]);
        acc[i][0] = __ffma2_rn(di, h01, acc[i][0]);
        acc[i][1] = __ffma2_rn(di, h23, acc[i][1]);
        accb[i] += dd[i];
      }
    }
    __syncthreads();
  }
  float* out = p.dw_part + (size_t)blockIdx.y * p.n_param;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * tn + i;
    if (n >= N) continue;
    const float v[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * tk + j;
      if (k < Kin) out[p.off_w[l] + (size_t)n * Kin + k] = v[j];
    }
    if (k0 == 0 && tk == 0) out[p.off_b[l] + n] = accb[i];
  }
}

struct TrainOut {
  float* gw[TR_MAXL];
  float* gb[TR_MAXL];
  double* loss;
  int nblocks_loss;
};
__global__ void k_train_reduce(const __grid_constant__ TrainPlan p, const __grid_constant__ TrainOut o) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int l = 0; l < p.n_layers; ++l) {
    const int64_t nw = (int64_t)p.N[l] * p.Kin[l];
    for (int64_t i = t0; i < nw; i += stride) {
      float s = 0.0f;
      for (int q = 0; q < p.splits; ++q) s += p.dw_part[(size_t)q * p.n_param + p.off_w[l] + i];
      o.gw[l][i] = s;
    }
    for (int64_t i = t0; i < p.N[l]; i += stride) {
      float s = 0.0f;
      for (int q = 0; q < p.splits; ++q) s += p.dw_part[(size_t)q * p.n_param + p.off_b[l] + i];
      o.gb[l][i] = s;
    }
  }
  if (t0 == 0) {
    double s = 0.0;
    for (int i = 0; i < o.nblocks_loss; ++i) s += p.pass_loss[i];
    *o.loss = s * (double)p.scale;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Hamiltonian leapfrog: the gradient dH/d(q, p) of a scalar-output MLP is the forward pass + the backward sweep to the
// inputs of the kernel above with delta_L = 1, kept entirely on-chip; one launch carries a pass of 32 trajectories
// through every kick-drift-kick step (BASELINE.json north_star / configs[4]).  Extension: the reference's own symplectic
// networks output dq/dt and dp/dt directly and run on the two-network field of ffb_integrate_fixed.
// ---------------------------------------------------------------------------------------------------------------
struct HamArgs {
  int64_t batch;
  int D, C, n_steps;
  float dt;
  const float* z0;
  const float* cond;
  float* z_out;
  float* h_out;
};
// shared memory: the training layout + the state rows [q | p | cond] and the gradient rows, both k-major
__host__ __device__ inline size_t ham_smem(int zfloats, int maxk, int kp0, int kb0, size_t* off /*[7]*/) {
  size_t v[7];
  size_t o = train_smem(zfloats, maxk, v);
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 127) & ~size_t(127); return r; };
  v[5] = take(sizeof(float) * kp0 * WD_RS);
  v[6] = take(sizeof(float) * kb0 * WD_RS);
  if (off) for (int i = 0; i < 7; ++i) off[i] = v[i];
  return o;
}

// H and dH/d(inputs) at the state rows S of this warp: forward keeping z_l, delta_L = 1, backward sweep to the inputs -> G
template <int ACT>
__device__ __forceinline__ void ham_grad(CtxW& cx, const TrainPlan& p, float* zreg, float* op, const float* S, float* G, float (&hval)[4]) {
  const int L = p.n_layers, r0w = cx.warp * 4;
  float* A = op;
  float* Bf = op + (size_t)p.maxk * WD_RS;
  for (int k = cx.lane; k < p.Kp[0]; k += 32)
    *reinterpret_cast<float4*>(A + (size_t)k * WD_RS + r0w) = *reinterpret_cast<const float4*>(S + (size_t)k * WD_RS + r0w);
  __syncwarp();
  for (int l = 0; l < L; ++l) {
    const float* bias = p.bp[l];
    if (l < L - 1) {
      float* z = zreg + p.zoff[l];
      train_layer(cx, A, p.Kp[l], p.Np[l], [&](int n, float (&v)[4]) {
        const float bj = bias[n];
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] += bj; h[i] = act_fwd<ACT>(v[i]); }
        *reinterpret_cast<float4*>(z + (size_t)n * WD_RS + r0w) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(Bf + (size_t)n * WD_RS + r0w) = make_float4(h[0], h[1], h[2], h[3]);
      });
    } else {
      train_layer(cx, A, p.Kp[l], p.Np[l], [&](int n, float (&v)[4]) {
        const float one = (n == 0) ? 1.0f : 0.0f;          // delta_L: d H / d (output column 0)
        if (n == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) hval[i] = v[i] + bias[0];
        }
        *reinterpret_cast<float4*>(Bf + (size_t)n * WD_RS + r0w) = make_float4(one, one, one, one);
      });
    }
    float* t = A; A = Bf; Bf = t;
  }
  for (int l = L - 1; l >= 1; --l) {
    const float* z = zreg + p.zoff[l - 1];
    train_layer(cx, A, p.Np[l], p.Kp[l], [&](int k, float (&v)[4]) {
      const float4 z4 = *reinterpret_cast<const float4*>(z + (size_t)k * WD_RS + r0w);
      const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
      float d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { float a, g; act_fwd_grad<ACT>(zz[i], a, g); d[i] = v[i] * g; }
      *reinterpret_cast<float4*>(Bf + (size_t)k * WD_RS + r0w) = make_float4(d[0], d[1], d[2], d[3]);
    });
    float* t = A; A = Bf; Bf = t;
  }
  train_layer(cx, A, p.Np[0], p.KB0, [&](int k, float (&v)[4]) {
    *reinterpret_cast<float4*>(G + (size_t)k * WD_RS + r0w) = make_float4(v[0], v[1], v[2], v[3]);
  });
}

template <int ACT>
__device__ __forceinline__ void ham_passes(CtxW& cx, const TrainPlan& p, const HamArgs& a, float* zreg, float* op, float* S, float* G,
                                           int64_t npass) {
  const int D = a.D, C = a.C, r0w = cx.warp * 4;
  const float dt = a.dt, half = 0.5f * a.dt;
  for (int64_t pass = blockIdx.x; pass < npass; pass += gridDim.x) {
    const int64_t row0 = pass * WD_R + r0w;
    for (int k = cx.lane; k < p.Kp[0]; k += 32) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x = 0.0f;
        if (row0 + i < a.batch) {
          if (k < 2 * D) x = a.z0[(row0 + i) * 2 * D + k];
          else if (k < 2 * D + C) x = a.cond[(row0 + i) * C + (k - 2 * D)];
        }
        v[i] = x;
      }
      *reinterpret_cast<float4*>(S + (size_t)k * WD_RS + r0w) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncwarp();
    float h0[4], h1[4];
    // target[k] += coef * G[src[k]] on this warp's rows
    auto axpy = [&](int dst0, int src0, float coef) {
      __syncwarp();
      for (int k = cx.lane; k < D; k += 32) {
        float4 s4 = *reinterpret_cast<float4*>(S + (size_t)(dst0 + k) * WD_RS + r0w);
        const float4 g4 = *reinterpret_cast<const float4*>(G + (size_t)(src0 + k) * WD_RS + r0w);
        s4.x = fmaf(coef, g4.x, s4.x); s4.y = fmaf(coef, g4.y, s4.y); s4.z = fmaf(coef, g4.z, s4.z); s4.w = fmaf(coef, g4.w, s4.w);
        *reinterpret_cast<float4*>(S + (size_t)(dst0 + k) * WD_RS + r0w) = s4;
      }
      __syncwarp();
    };
    for (int step = 0; step < a.n_steps; ++step) {
      ham_grad<ACT>(cx, p, zreg, op, S, G, step == 0 ? h0 : h1);
      axpy(D, 0, -half);                        // kick:  p -= dt/2 dH/dq (q, p)
      ham_grad<ACT>(cx, p, zreg, op, S, G, h1);
      axpy(0, D, dt);                           // drift: q += dt dH/dp (q, p_half)
      ham_grad<ACT>(cx, p, zreg, op, S, G, h1);
      axpy(D, 0, -half);                        // kick:  p -= dt/2 dH/dq (q_new, p_half)
    }
    ham_grad<ACT>(cx, p, zreg, op, S, G, h1);    // H at the end (and at the start when there are no steps)
    if (a.n_steps == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) h0[i] = h1[i];
    }
    __syncwarp();
    for (int k = cx.lane; k < 2 * D; k += 32) {
      const float4 s4 = *reinterpret_cast<const float4*>(S + (size_t)k * WD_RS + r0w);
      const float v[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) if (row0 + i < a.batch) a.z_out[(row0 + i) * 2 * D + k] = v[i];
    }
    if (a.h_out && cx.lane == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) if (row0 + i < a.batch) { a.h_out[(row0 + i) * 2] = h0[i]; a.h_out[(row0 + i) * 2 + 1] = h1[i]; }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(NTHR, 1) k_ham_leapfrog(const __grid_constant__ TrainPlan p, const __grid_constant__ HamArgs a,
                                                          const int64_t npass, const int zfloats) {
  CtxW cx;
  size_t off[7];
  ham_smem(zfloats, p.maxk, p.Kp[0], p.KB0, off);
  float* zreg = reinterpret_cast<float*>(smem_base() + off[0]);
  float* op = reinterpret_cast<float*>(smem_base() + off[1]);
  cx.o_ring = (uint32_t)off[2]; cx.o_bar = (uint32_t)off[3];
  float* S = reinterpret_cast<float*>(smem_base() + off[5]);
  float* G = reinterpret_cast<float*>(smem_base() + off[6]);
  cx.tid = threadIdx.x; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
  cx.producer = (cx.warp == NCOMP / 32);
  cx.stage = 0; cx.phase = cx.producer ? 1u : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WD_NSTAGE; ++s) { mbar_init(&cx.full()[s], 1); mbar_init(&cx.empty()[s], NCOMP / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cx.producer) {
    for (int64_t pass = blockIdx.x; pass < npass; pass += gridDim.x)
      for (int e = 0; e < 3 * a.n_steps + 1; ++e) {
        for (int l = 0; l < p.n_layers; ++l) train_produce(cx, p.Wf[l], p.Kp[l], p.Np[l]);
        for (int l = p.n_layers - 1; l >= 1; --l) train_produce(cx, p.Wb[l], p.Np[l], p.Kp[l]);
        train_produce(cx, p.Wb[0], p.Np[0], p.KB0);
      }
    return;
  }
  FFB_ACT_DISPATCH(p.act, (ham_passes<ACT>(cx, p, a, zreg, op, S, G, npass)));
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
static int tr_smem_optin() {
  static int v = 0;
  if (!v) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); }
  return v;
}

struct TrainLayout {
  size_t total_floats;       // floats of work memory before the FP64 loss partials
  size_t loss_off_bytes;     // byte offset of the loss partials
  size_t bytes;
  int zfloats, maxk, splits, grid;
};

// fills the shape part of the plan and the layout of the work buffer; pointers are set by the caller from `work`
static int train_plan(const ffb_net_desc* d, int64_t batch, int want_dx, TrainPlan* p, TrainLayout* lay, float* work) {
  if (!d) return ffb_fail(FFB_ERR_ARG, "ffb_train: null network description");
  if (d->n_layers < 1 || d->n_layers > FFB_MAX_LAYERS) return ffb_fail(FFB_ERR_ARG, "ffb_train: 1..16 Linear layers supported");
  if (d->in_features < 1 || d->in_features > FFB_MAX_WIDTH) return ffb_fail(FFB_ERR_ARG, "ffb_train: in_features must be in 1..512");
  for (int l = 0; l < d->n_layers; ++l)
    if (d->widths[l] < 1 || d->widths[l] > FFB_MAX_WIDTH) return ffb_fail(FFB_ERR_ARG, "ffb_train: layer widths must be in 1..512");
  if (d->activation < FFB_ACT_SILU || d->activation > FFB_ACT_GELU) return ffb_fail(FFB_ERR_ARG, "ffb_train: unknown activation");
  if (batch < 0) return ffb_fail(FFB_ERR_ARG, "ffb_train: negative batch");
  memset(p, 0, sizeof(*p));
  const bool fwd = (want_dx == 2);          // forward-only sizing: weight images and biases, nothing per row
  want_dx = (want_dx == 1) ? 1 : 0;
  p->n_layers = d->n_layers; p->in_features = d->in_features; p->act = d->activation; p->want_dx = want_dx; p->batch = batch;
  p->fwd_only = fwd ? 1 : 0;
  int in_f = d->in_features, zf = 0, maxk = 0;
  int64_t np = 0;
  for (int l = 0; l < d->n_layers; ++l) {
    p->N[l] = d->widths[l]; p->Kin[l] = in_f;
    p->Np[l] = tr_padw(d->widths[l]);
    p->Kp[l] = (l == 0) ? ((in_f + 3) & ~3) : p->Np[l - 1];
    maxk = std::max(maxk, std::max(p->Kp[l], p->Np[l]));
    if (l < d->n_layers - 1) { p->zoff[l] = zf; zf += p->Np[l] * WD_RS; }
    p->off_w[l] = np; np += (int64_t)p->N[l] * in_f;
    p->off_b[l] = np; np += p->N[l];
    in_f = d->widths[l];
  }
  p->KB0 = tr_padw(d->in_features);
  if (want_dx) maxk = std::max(maxk, p->KB0);
  p->maxk = maxk; p->n_param = np;
  const int sms = ffb_num_sms();
  const int64_t npass = (batch + WD_R - 1) / WD_R;
  lay->grid = (int)std::max<int64_t>(1, std::min<int64_t>(npass, sms));
  lay->zfloats = zf; lay->maxk = maxk;
  // enough (patch, split) blocks of k_train_dw to cover the SMs, at least 256 rows per split
  const int patches = dw_patches(*p);
  int splits = (2 * sms + patches - 1) / patches;
  splits = (int)std::max<int64_t>(1, std::min<int64_t>(splits, (batch + 255) / 256));
  splits = std::min(splits, 64);
  p->splits = lay->splits = splits;
  // work buffer layout (floats)
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) & ~size_t(63); return r; };
  for (int l = 0; l < d->n_layers; ++l) {
    const size_t a = take((size_t)p->Kp[l] * p->Np[l]);
    const size_t KB = (l == 0) ? (size_t)p->KB0 : (size_t)p->Kp[l];
    const size_t b = (l >= 1 || want_dx) ? take((size_t)p->Np[l] * KB) : 0;
    const size_t c = take(p->Np[l]);
    const size_t h = (l >= 1 && !fwd) ? take((size_t)batch * p->Np[l - 1]) : 0;
    const size_t dd = fwd ? 0 : take((size_t)batch * p->Np[l]);
    if (work) {
      p->Wf[l] = work + a; p->Wb[l] = (l >= 1 || want_dx) ? work + b : nullptr; p->bp[l] = work + c;
      p->hin[l] = (l >= 1 && !fwd) ? work + h : nullptr; p->dl[l] = fwd ? nullptr : work + dd;
    }
  }
  const size_t dwp = fwd ? 0 : take((size_t)splits * np);
  if (work) p->dw_part = work + dwp;
  lay->total_floats = o;
  lay->loss_off_bytes = (o * sizeof(float) + 255) & ~size_t(255);
  lay->bytes = lay->loss_off_bytes + sizeof(double) * (size_t)lay->grid;
  if (work) p->pass_loss = reinterpret_cast<double*>(reinterpret_cast<char*>(work) + lay->loss_off_bytes);
  return FFB_OK;
}

}  // namespace ffb

using namespace ffb;

extern "C" size_t ffb_train_work_bytes(const ffb_net_desc* net, int64_t batch, int32_t want_grad_x) {
  TrainPlan p; TrainLayout lay;
  if (train_plan(net, batch, want_grad_x, &p, &lay, nullptr)) return 0;
  return lay.bytes;
}

extern "C" int ffb_train_step(const ffb_net_desc* net, const ffb_train_args* a, void* stream_) {
  if (!net || !a) return ffb_fail(FFB_ERR_ARG, "ffb_train_step: null argument");
  // forward only: no loss / gradient outputs at all, just out = net(X) (a network evaluation at per-row inputs, e.g. the
  // reference's score(t, x) with one time per sample, diffusion.py:82-121)
  const bool fwd_only = a->out && !a->loss && !a->grad_x && !a->grad_w[0] && !a->grad_b[0];
  if (!a->x_in || !a->work || (!fwd_only && ((!a->beta && !a->cot) || !a->loss)))
    return ffb_fail(FFB_ERR_ARG, "ffb_train_step: x_in, beta (or cot), loss and work are required (or only `out` for a forward pass)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  TrainPlan p; TrainLayout lay;
  int rc = train_plan(net, a->batch, fwd_only ? 2 : (a->grad_x != nullptr ? 1 : 0), &p, &lay, a->work);
  if (rc) return rc;
  if (fwd_only) lay.zfloats = 0;
  TrainOut o;
  memset(&o, 0, sizeof(o));
  for (int l = 0; l < net->n_layers; ++l) {
    if (!net->weight[l] || !net->bias[l] || (!fwd_only && (!a->grad_w[l] || !a->grad_b[l])))
      return ffb_fail(FFB_ERR_ARG, "ffb_train_step: weight, bias, grad_w and grad_b are required for every layer");
    p.W[l] = net->weight[l]; p.B[l] = net->bias[l];
    o.gw[l] = a->grad_w[l]; o.gb[l] = a->grad_b[l];
  }
  p.x_in = a->x_in; p.alpha = a->alpha; p.beta = a->beta; p.scale = a->scale; p.grad_x = a->grad_x;
  p.cot = a->cot; p.out = a->out;
  o.loss = a->loss; o.nblocks_loss = lay.grid;
  const size_t smem = train_smem(lay.zfloats, lay.maxk, nullptr);
  if ((int)smem > tr_smem_optin())
    return ffb_fail(FFB_ERR_ARG, "ffb_train_step: the network needs " + std::to_string(smem) + " B of shared memory (pre-activations of "
                    "every hidden layer x 32 rows + two operand buffers), the device allows " + std::to_string(tr_smem_optin()));
  const int64_t npass = (a->batch + WD_R - 1) / WD_R;
  k_train_pack<<<2 * ffb_num_sms(), 256, 0, st>>>(p);
  TR_CUDA_TRY(cudaFuncSetAttribute(k_train_fwdbwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_train_fwdbwd<<<lay.grid, NTHR, smem, st>>>(p, npass, lay.zfloats);
  if (!fwd_only) {
    k_train_dw<<<dim3(dw_patches(p), p.splits), 256, 0, st>>>(p);
    k_train_reduce<<<ffb_num_sms(), 256, 0, st>>>(p, o);
  }
  ffb_count_launches(fwd_only ? 2 : 4);
  TR_CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_hamiltonian_leapfrog(const ffb_net_desc* net, const ffb_hamiltonian_args* a, void* stream_) {
  if (!net || !a) return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: null argument");
  if (!a->z0 || !a->z_out || !a->work) return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: z0, z_out and work are required");
  if (a->dim < 1 || a->cond_dim < 0 || 2 * a->dim + a->cond_dim != net->in_features)
    return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: the network's in_features must be 2 * dim + cond_dim ([q | p | cond])");
  if (net->n_layers < 1 || net->widths[net->n_layers - 1] != 1)
    return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: H is a scalar: the last layer must have one output");
  if (a->cond_dim > 0 && !a->cond) return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: cond is required");
  if (a->n_steps < 0) return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: negative n_steps");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  TrainPlan p; TrainLayout lay;
  int rc = train_plan(net, 0, 1, &p, &lay, a->work);       // batch 0: only the weight images live in `work`
  if (rc) return rc;
  for (int l = 0; l < net->n_layers; ++l) {
    if (!net->weight[l] || !net->bias[l]) return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: weight and bias are required for every layer");
    p.W[l] = net->weight[l]; p.B[l] = net->bias[l];
  }
  HamArgs h;
  h.batch = a->batch; h.D = a->dim; h.C = a->cond_dim; h.n_steps = a->n_steps; h.dt = a->dt;
  h.z0 = a->z0; h.cond = a->cond; h.z_out = a->z_out; h.h_out = a->h_out;
  const size_t smem = ham_smem(lay.zfloats, lay.maxk, p.Kp[0], p.KB0, nullptr);
  if ((int)smem > tr_smem_optin())
    return ffb_fail(FFB_ERR_ARG, "ffb_hamiltonian_leapfrog: the network needs " + std::to_string(smem) + " B of shared memory, the device allows " +
                    std::to_string(tr_smem_optin()));
  const int64_t npass = (a->batch + WD_R - 1) / WD_R;
  if (npass == 0) return FFB_OK;
  k_train_pack<<<2 * ffb_num_sms(), 256, 0, st>>>(p);
  TR_CUDA_TRY(cudaFuncSetAttribute(k_ham_leapfrog, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(npass, ffb_num_sms());
  k_ham_leapfrog<<<grid, NTHR, smem, st>>>(p, h, npass, lay.zfloats);
  ffb_count_launches(2);
  TR_CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

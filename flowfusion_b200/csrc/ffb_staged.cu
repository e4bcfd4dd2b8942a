// ffb_staged.cu -- staged (evaluation-at-a-time) solves for the Hutch++ and XTrace divergence estimators
// (reference diffusion.py:336-400 and :402-481).
//
// Why staged.  Both estimators multiply by the transposed field Jacobian A = (d x_dot / d x)^T twice with a
// per-sample thin QR in between (the second round's vectors are the Q of the first round's products), so an
// evaluation cannot stay inside the fused dopri5 attempt kernel the way the exact / Hutchinson traces do.
// Instead one evaluation is
//     ffb_field_eval(FFB_DIV_EXACT, jac = J)   the tangent-row tensor-core engine writes the field AND the full
//                                               network Jacobian (its D tangent rows, 4 D^2 bytes per sample)
//     k_trace_coop                              the estimator algebra, 16 or 32 LANES PER SAMPLE: lane j holds row j of
//                                               J^T in registers, vectors are distributed over the lanes (see below)
// and a dopri5 attempt is 6 x (k_rk_combine, field_eval, k_trace_coop) + k_rk_finish.  The Jacobian round trip
// (2 x 4 D^2 B per sample and evaluation, 2 KB at D = 16) costs ~3 % of the evaluation's tensor time; the algebra
// runs at full-GPU parallelism instead of on the 7 owner threads of a tangent tile.
//
// With the whole Jacobian on hand every product A v is a D x D mat-vec, so the estimators cost one exact-trace
// evaluation plus O(D^2 (r + m)) flops -- they reproduce the reference's numbers for given probes (parity), they
// are not cheaper than the exact trace at these D.  A reverse-mode engine (r + m sweeps instead of D tangents)
// is the follow-up for D >> 32.
//
// A one-thread-per-sample version of the algebra (trace_estimate_one) is __host__ __device__:
// ffb_trace_estimate_host runs it on the CPU (tests/test_trace_estimators.py checks it against the oracle here,
// without a GPU) and FFB_TRACE_KERNEL=thread launches it on the GPU (A/B; 8x slower than the cooperative kernel).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>

#include "ffb200.h"
#include "ffb_common.cuh"

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ffb_fail(FFB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
  } while (0)

namespace ffb {

// The per-sample algebra is templated on the padded sizes of its work arrays (DMAX >= D, KMAX >= rank): the arrays
// are indexed at run time, so they live in local memory, and with one [32][8] size for all (~4.4 KB per thread) the
// footprint of a resident block set overflowed L1 and the kernel ran at L2 latency (6.4 ms per 10^6 samples at
// D = 16, rank 1).  Sized to the rank it stays in L1.

// A v = J_f^T v for x_dot = a x - c net(x) [/ sigma]  (diffusion.py:233-238, 276-279), or = J_net^T v
struct TraceOp {
  const float* A;      // [j][n] = d net_n / d x_j
  int D;
  bool score, use_sigma, has_drift;
  float a, c, sigma;
  __host__ __device__ inline void apply(const float* v, float* out) const {
    for (int j = 0; j < D; ++j) {
      float acc = 0.0f;
      for (int n = 0; n < D; ++n) acc = fmaf(A[j * D + n], v[n], acc);
      if (score) {
        const float s = use_sigma ? acc / sigma : acc;
        out[j] = (has_drift ? a * v[j] : 0.0f) - c * s;
      } else {
        out[j] = acc;
      }
    }
  }
};

// Householder QR in place (LAPACK geqrf conventions): R in the upper triangle of Y (D x k, row stride TR_MAXK),
// the reflectors below the diagonal (v_jj = 1 implicit), their scales in tau.
template <int TR_MAXK>
__host__ __device__ inline void tr_qr(float* Y, float* tau, int D, int k) {
  for (int j = 0; j < k; ++j) {
    const float alpha = Y[j * TR_MAXK + j];
    float ss = 0.0f;
    for (int d = j + 1; d < D; ++d) ss = fmaf(Y[d * TR_MAXK + j], Y[d * TR_MAXK + j], ss);
    float t = 0.0f;
    if (ss != 0.0f) {
      const float beta = -copysignf(sqrtf(fmaf(alpha, alpha, ss)), alpha);
      t = (beta - alpha) / beta;
      const float scal = 1.0f / (alpha - beta);
      for (int d = j + 1; d < D; ++d) Y[d * TR_MAXK + j] *= scal;
      Y[j * TR_MAXK + j] = beta;
    }
    tau[j] = t;
    if (t != 0.0f) {
      for (int c = j + 1; c < k; ++c) {
        float w = Y[j * TR_MAXK + c];
        for (int d = j + 1; d < D; ++d) w = fmaf(Y[d * TR_MAXK + j], Y[d * TR_MAXK + c], w);
        w *= t;
        Y[j * TR_MAXK + c] -= w;
        for (int d = j + 1; d < D; ++d) Y[d * TR_MAXK + c] = fmaf(-w, Y[d * TR_MAXK + j], Y[d * TR_MAXK + c]);
      }
    }
  }
}
// thin Q (D x k) = H_0 ... H_{k-1} [I_k; 0]  (LAPACK orgqr)
template <int TR_MAXK>
__host__ __device__ inline void tr_formq(const float* Y, const float* tau, float* Q, int D, int k) {
  for (int d = 0; d < D; ++d)
    for (int i = 0; i < k; ++i) Q[d * TR_MAXK + i] = (d == i) ? 1.0f : 0.0f;
  for (int j = k - 1; j >= 0; --j) {
    if (tau[j] == 0.0f) continue;
    for (int c = j; c < k; ++c) {
      float w = Q[j * TR_MAXK + c];
      for (int d = j + 1; d < D; ++d) w = fmaf(Y[d * TR_MAXK + j], Q[d * TR_MAXK + c], w);
      w *= tau[j];
      Q[j * TR_MAXK + c] -= w;
      for (int d = j + 1; d < D; ++d) Q[d * TR_MAXK + c] = fmaf(-w, Y[d * TR_MAXK + j], Q[d * TR_MAXK + c]);
    }
  }
}

// Hutch++ (diffusion.py:336-400): Y = A S, Q = qr(Y), tr(Q^T A Q) + mean_l u_l^T A u_l, u_l = (I - Q Q^T) g_l
template <int TR_MAXD, int TR_MAXK>
__host__ __device__ inline float hutchpp_one(const ffb_trace_args& a, const TraceOp& op, int64_t b) {
  const int D = a.dim, k = a.rank;
  float Y[TR_MAXD * TR_MAXK], Q[TR_MAXD * TR_MAXK], tau[TR_MAXK], v[TR_MAXD], w[TR_MAXD], qg[TR_MAXK];
  for (int i = 0; i < k; ++i) {
    const float* s = a.S + ((int64_t)i * a.batch + b) * D;
    for (int d = 0; d < D; ++d) v[d] = s[d];
    op.apply(v, w);
    for (int d = 0; d < D; ++d) Y[d * TR_MAXK + i] = w[d];
  }
  tr_qr<TR_MAXK>(Y, tau, D, k);
  tr_formq<TR_MAXK>(Y, tau, Q, D, k);
  float trace_lr = 0.0f;
  for (int i = 0; i < k; ++i) {                                   // sum_i q_i^T A q_i   (:376-381)
    for (int d = 0; d < D; ++d) v[d] = Q[d * TR_MAXK + i];
    op.apply(v, w);
    for (int d = 0; d < D; ++d) trace_lr = fmaf(v[d], w[d], trace_lr);
  }
  float trace_res = 0.0f;
  for (int l = 0; l < a.nvec; ++l) {                              // residual probes (:383-396)
    const float* g = a.G + ((int64_t)l * a.batch + b) * D;
    for (int i = 0; i < k; ++i) {
      float acc = 0.0f;
      for (int d = 0; d < D; ++d) acc = fmaf(Q[d * TR_MAXK + i], g[d], acc);
      qg[i] = acc;
    }
    for (int d = 0; d < D; ++d) {
      float proj = 0.0f;
      for (int i = 0; i < k; ++i) proj = fmaf(Q[d * TR_MAXK + i], qg[i], proj);
      v[d] = g[d] - proj;
    }
    op.apply(v, w);
    for (int d = 0; d < D; ++d) trace_res = fmaf(v[d], w[d], trace_res);
  }
  return trace_lr + trace_res / (float)a.nvec;                    // :398
}

// The k x k tail of XTrace (diffusion.py:453-477) on H = Q^T Z, W = Q^T O, T = Z^T O and the upper-triangular R
// (row stride K): shared by the per-sample path and the cooperative kernel.
template <int K>
__host__ __device__ inline float xtrace_tail(const float* H, const float* W, const float* T, const float* R, int k) {
  float St[K * K], X[K * K];
  // St = inv(R) by back substitution (:453), rows scaled to unit 2-norm (:455)
  for (int j = 0; j < k; ++j) {
    for (int i = k - 1; i >= 0; --i) {
      if (i > j) { St[i * K + j] = 0.0f; continue; }
      float acc = (i == j) ? 1.0f : 0.0f;
      for (int l = i + 1; l <= j; ++l) acc = fmaf(-R[i * K + l], St[l * K + j], acc);
      St[i * K + j] = acc / R[i * K + i];
    }
  }
  for (int i = 0; i < k; ++i) {
    float ss = 0.0f;
    for (int j = 0; j < k; ++j) ss = fmaf(St[i * K + j], St[i * K + j], ss);
    const float nrm = sqrtf(ss);
    for (int j = 0; j < k; ++j) St[i * K + j] /= nrm;
  }
  // S = St^T (:456): S[p][q] = St[q][p]
  float trace_H = 0.0f;
  for (int i = 0; i < k; ++i) trace_H += H[i * K + i];            // :459
  float total = 0.0f;
  for (int q = 0; q < k; ++q) {                                   // one estimate per probe (:461-477)
    float ws = 0.0f, sr = 0.0f;
    for (int p = 0; p < k; ++p) {
      const float s = St[q * K + p];
      ws = fmaf(s, W[p * K + q], ws);                             // WS = sum_p W[p][q] S[p][q]
      if (p <= q) sr = fmaf(s, R[p * K + q], sr);                 // SR = sum_p S[p][q] R[p][q]  (R upper triangular)
    }
    for (int p = 0; p < k; ++p) X[p * K + q] = W[p * K + q] - ws * St[q * K + p];    // :463
    float shs = 0.0f, xhx = 0.0f, tx = 0.0f;
    for (int i = 0; i < k; ++i) {
      float hs = 0.0f, hx = 0.0f;
      for (int p = 0; p < k; ++p) {
        hs = fmaf(H[i * K + p], St[q * K + p], hs);
        hx = fmaf(H[i * K + p], X[p * K + q], hx);
      }
      shs = fmaf(St[q * K + i], hs, shs);                         // :465
      xhx = fmaf(X[i * K + q], hx, xhx);                          // :467
      tx = fmaf(T[i * K + q], X[i * K + q], tx);                  // :473
    }
    total += trace_H - shs + ws * sr - tx + xhx;                  // :475
  }
  return total / (float)k;                                        // :477
}
// XTrace (diffusion.py:402-481), index for index; k = m (m <= D is enforced by the caller as in :410)
template <int TR_MAXD, int TR_MAXK>
__host__ __device__ inline float xtrace_one(const ffb_trace_args& a, const TraceOp& op, int64_t b) {
  const int D = a.dim, k = a.rank;
  constexpr int K = TR_MAXK;
  float Y[TR_MAXD * K], Q[TR_MAXD * K], Z[TR_MAXD * K], tau[K], v[TR_MAXD], w[TR_MAXD];
  float H[K * K], W[K * K], T[K * K];
  for (int i = 0; i < k; ++i) {                                   // Y = A O (:433-435)
    const float* o = a.S + ((int64_t)i * a.batch + b) * D;
    for (int d = 0; d < D; ++d) v[d] = o[d];
    op.apply(v, w);
    for (int d = 0; d < D; ++d) Y[d * K + i] = w[d];
  }
  tr_qr<TR_MAXK>(Y, tau, D, k);                                            // R = upper triangle of Y (:438)
  tr_formq<TR_MAXK>(Y, tau, Q, D, k);
  for (int i = 0; i < k; ++i) {                                   // Z = A Q (:443-445)
    for (int d = 0; d < D; ++d) v[d] = Q[d * K + i];
    op.apply(v, w);
    for (int d = 0; d < D; ++d) Z[d * K + i] = w[d];
  }
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      const float* o = a.S + ((int64_t)j * a.batch + b) * D;      // probe j
      float h = 0.0f, wq = 0.0f, tz = 0.0f;
      for (int d = 0; d < D; ++d) {
        h = fmaf(Q[d * K + i], Z[d * K + j], h);                  // H = Q^T Z      (:447)
        wq = fmaf(Q[d * K + i], o[d], wq);                        // W = Q^T O      (:449)
        tz = fmaf(Z[d * K + i], o[d], tz);                        // T = Z^T O      (:451)
      }
      H[i * K + j] = h; W[i * K + j] = wq; T[i * K + j] = tz;
    }
  float R[K * K];
  for (int p = 0; p < k; ++p)
    for (int q = 0; q < k; ++q) R[p * K + q] = (p <= q) ? Y[p * K + q] : 0.0f;
  return xtrace_tail<K>(H, W, T, R, k);
}

template <int TR_MAXD, int TR_MAXK>
__host__ __device__ inline float trace_estimate_one(const ffb_trace_args& a, const float* A, int64_t b) {
  TraceOp op;
  op.A = A; op.D = a.dim;
  op.score = a.score != 0; op.use_sigma = a.use_sigma != 0; op.has_drift = a.has_drift != 0;
  op.a = a.a; op.c = a.c; op.sigma = a.sigma;
  const float v = (a.kind == FFB_TRACE_HUTCHPP) ? hutchpp_one<TR_MAXD, TR_MAXK>(a, op, b) : xtrace_one<TR_MAXD, TR_MAXK>(a, op, b);
  return v * a.sign;
}

// deterministic block reduction of one double per thread (thread 0 returns the sum)
__device__ __forceinline__ double st_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += red[w];
  return s;
}

}  // namespace ffb

// One thread per sample.  STAGE = false (default): every thread streams its own contiguous D*D Jacobian straight from
// global memory (each 32-byte sector is fetched once per pass and then served by L1), no shared memory, so the
// occupancy is set by registers alone.  STAGE = true (first version, measured 6.4 ms per 10^6 rows against 3.0 ms): the block's Jacobians go
// through shared memory first (coalesced load, row stride D*D + 1: conflict-free reads) -- 1 KB of shared memory per
// thread at D = 16 caps an SM at 192 threads and every latency of the serial per-sample algebra is exposed.
template <int DMAX, int KMAX, bool STAGE>
__global__ void k_trace_estimate(const __grid_constant__ ffb_trace_args a, const int64_t ntiles) {
  using namespace ffb;
  extern __shared__ float sA[];
  __shared__ double red[32];
  const int D = a.dim, DD = D * D, stride = DD + 1;
  double q2 = 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * blockDim.x;
    const int nv = (int)min((int64_t)blockDim.x, a.batch - row0);
    if (STAGE) {
      const float* __restrict__ src = a.jac + row0 * DD;
      __syncthreads();
      for (int idx = threadIdx.x; idx < nv * DD; idx += blockDim.x) {
        const int s = idx / DD, e = idx - s * DD;
        sA[s * stride + e] = src[idx];
      }
      __syncthreads();
    }
    if ((int)threadIdx.x < nv) {
      const int64_t b = row0 + threadIdx.x;
      const float* A = STAGE ? sA + threadIdx.x * stride : a.jac + b * DD;
      const float dv = trace_estimate_one<DMAX, KMAX>(a, A, b);
      a.dlp[b] = dv;
      if (a.norms == 1) {
        const float q = dv / a.atol;
        q2 += (double)q * q;
      } else if (a.norms == 2) {
        const float q = (dv - a.dlpbase[b]) / a.atol;
        q2 += (double)q * q;
      }
    }
  }
  if (a.norms) {
    const double s = st_block_sum(q2, red);
    if (threadIdx.x == 0) a.partials[(int64_t)blockIdx.x * FFB_NPART + (a.norms == 1 ? P_LP_F : P_LP_DF)] = s;
  }
}

// =============================================================================================
// Cooperative estimator kernel (the default): LANES = 16 or 32 lanes per sample.
// Lane j keeps ROW j of A = J^T in registers (the group reads one contiguous D*D block: coalesced, once), vectors
// are distributed one element per lane, A v = D shuffles + D FMAs per lane, inner products are butterfly
// reductions (every lane gets the same bits), the Householder QR runs on distributed columns and the k x k XTrace
// tail redundantly on every lane.  Same mathematics as trace_estimate_one with tree- instead of serially-ordered
// sums; tests compare both with the oracle.  The thread-per-sample kernel re-read each Jacobian once per pass from
// DRAM (3.7 KB / row against 1.2 KB algorithmic, ncu) behind dependent 32-byte loads: 3.0 ms per 10^6 rows.
// =============================================================================================
namespace ffb {

template <int LANES>
__device__ __forceinline__ float co_sum(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LANES);
  return v;
}

template <int LANES>
struct CoOp {
  float row[LANES];       // A[lane][n]
  bool score, use_sigma, has_drift;
  float a, c, sigma;
  __device__ __forceinline__ float apply(float v) const {
    float acc = 0.0f;
#pragma unroll
    for (int n = 0; n < LANES; ++n) acc = fmaf(row[n], __shfl_sync(0xffffffffu, v, n, LANES), acc);
    if (!score) return acc;
    const float s = use_sigma ? acc / sigma : acc;
    return (has_drift ? a * v : 0.0f) - c * s;
  }
};

// Householder QR of the distributed columns Y[0..k) (lane d holds row d), LAPACK conventions as tr_qr; then the thin Q
template <int LANES, int K>
__device__ __forceinline__ void co_qr(float (&Y)[K], float (&Q)[K], int k, int lane) {
  float tau[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    tau[j] = 0.0f;
    if (j < k) {
      const float alpha = __shfl_sync(0xffffffffu, Y[j], j, LANES);
      const float ss = co_sum<LANES>(lane > j ? Y[j] * Y[j] : 0.0f);
      const bool on = ss != 0.0f;
      const float beta = -copysignf(sqrtf(fmaf(alpha, alpha, ss)), alpha);
      const float t = on ? (beta - alpha) / beta : 0.0f;
      const float scal = on ? 1.0f / (alpha - beta) : 1.0f;
      if (lane > j) Y[j] *= scal;
      if (lane == j && on) Y[j] = beta;
      tau[j] = t;
#pragma unroll
      for (int c = j + 1; c < K; ++c) {
        if (c < k) {
          const float w = t * co_sum<LANES>(lane > j ? Y[j] * Y[c] : (lane == j ? Y[c] : 0.0f));
          if (lane == j) Y[c] -= w;
          if (lane > j) Y[c] = fmaf(-w, Y[j], Y[c]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < K; ++i) Q[i] = (lane == i) ? 1.0f : 0.0f;
#pragma unroll
  for (int j = K - 1; j >= 0; --j) {
    if (j < k) {
#pragma unroll
      for (int c = j; c < K; ++c) {
        if (c < k) {
          const float w = tau[j] * co_sum<LANES>(lane > j ? Y[j] * Q[c] : (lane == j ? Q[c] : 0.0f));
          if (lane == j) Q[c] -= w;
          if (lane > j) Q[c] = fmaf(-w, Y[j], Q[c]);
        }
      }
    }
  }
}

}  // namespace ffb

template <int LANES, int K>
__global__ void __launch_bounds__(128) k_trace_coop(const __grid_constant__ ffb_trace_args a, const int64_t ngroups) {
  using namespace ffb;
  __shared__ double red[32];
  constexpr int GPB = 128 / LANES;                  // samples per block and pass
  const int D = a.dim, k = a.rank, lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
  const bool row_on = lane < D;
  double q2 = 0.0;
  for (int64_t g0 = (int64_t)blockIdx.x * GPB; g0 < ngroups; g0 += (int64_t)gridDim.x * GPB) {
    const int64_t bb = g0 + grp;
    const bool live = bb < a.batch;
    const int64_t b = live ? bb : a.batch - 1;      // idle groups shadow the last sample: every lane runs every shuffle
    CoOp<LANES> op;
    op.score = a.score != 0; op.use_sigma = a.use_sigma != 0; op.has_drift = a.has_drift != 0;
    op.a = a.a; op.c = a.c; op.sigma = a.sigma;
    const float* __restrict__ ar = a.jac + (b * D + (row_on ? lane : 0)) * D;
    if ((D & 3) == 0) {
#pragma unroll
      for (int n = 0; n < LANES; n += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_on && n < D) v = *reinterpret_cast<const float4*>(ar + n);
        op.row[n] = v.x; op.row[n + 1] = v.y; op.row[n + 2] = v.z; op.row[n + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int n = 0; n < LANES; ++n) op.row[n] = (row_on && n < D) ? ar[n] : 0.0f;
    }
    float P[K], Y[K], Q[K];                          // probes (S or O), A P, thin Q: lane d holds row d
#pragma unroll
    for (int i = 0; i < K; ++i) {
      P[i] = (i < k && row_on) ? a.S[((int64_t)i * a.batch + b) * D + lane] : 0.0f;
      Y[i] = 0.0f;
      if (i < k) Y[i] = op.apply(P[i]);
    }
    co_qr<LANES, K>(Y, Q, k, lane);
    float est;
    if (a.kind == FFB_TRACE_HUTCHPP) {
      float trace_lr = 0.0f, trace_res = 0.0f;
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (i < k) trace_lr += co_sum<LANES>(Q[i] * op.apply(Q[i]));
      for (int l = 0; l < a.nvec; ++l) {
        const float gv = row_on ? a.G[((int64_t)l * a.batch + b) * D + lane] : 0.0f;
        float u = gv;
#pragma unroll
        for (int i = 0; i < K; ++i)
          if (i < k) u = fmaf(-Q[i], co_sum<LANES>(Q[i] * gv), u);
        trace_res += co_sum<LANES>(u * op.apply(u));
      }
      est = trace_lr + trace_res / (float)a.nvec;
    } else {
      float Z[K], H[K * K], W[K * K], T[K * K], R[K * K];
#pragma unroll
      for (int i = 0; i < K; ++i) Z[i] = (i < k) ? op.apply(Q[i]) : 0.0f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const bool in = i < k && j < k;            // uniform
          H[i * K + j] = in ? co_sum<LANES>(Q[i] * Z[j]) : 0.0f;
          W[i * K + j] = in ? co_sum<LANES>(Q[i] * P[j]) : 0.0f;
          T[i * K + j] = in ? co_sum<LANES>(Z[i] * P[j]) : 0.0f;
          R[i * K + j] = (in && i <= j) ? __shfl_sync(0xffffffffu, Y[j], i, LANES) : 0.0f;
        }
      }
      est = xtrace_tail<K>(H, W, T, R, k);
    }
    if (live && lane == 0) {
      const float dv = est * a.sign;
      a.dlp[b] = dv;
      if (a.norms == 1) {
        const float q = dv / a.atol;
        q2 += (double)q * q;
      } else if (a.norms == 2) {
        const float q = (dv - a.dlpbase[b]) / a.atol;
        q2 += (double)q * q;
      }
    }
  }
  if (a.norms) {
    const double s = st_block_sum(q2, red);
    if (threadIdx.x == 0) a.partials[(int64_t)blockIdx.x * FFB_NPART + (a.norms == 1 ? P_LP_F : P_LP_DF)] = s;
  }
}

__global__ void k_rk_combine(const __grid_constant__ ffb_rk_combine_args a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = __fmul_rn(a.k[0][i], a.coef[0]);
#pragma unroll
    for (int j = 1; j < 7; ++j)
      if (j < a.n_terms) acc = fmaf(a.k[j][i], a.coef[j], acc);
    a.out[i] = __fadd_rn(a.y0[i], acc);
  }
}

// the tail of k_dopri5 (ffb_kernels.cu) on global buffers: error sums, log-det column, dense output
__global__ void k_rk_finish(const __grid_constant__ ffb_rk_finish_args a) {
  using namespace ffb;
  __shared__ double red[32];
  const int64_t nx = a.batch * a.dim;
  const int64_t step = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ex = 0.0, el = 0.0, nonfinite = 0.0;
  const int nk = (a.n_k >= 2 && a.n_k <= 7) ? a.n_k : 7;       // stage derivatives of the method; f1 = the last one
  for (int64_t i = i0; i < nx; i += step) {
    const float y0 = a.y0[i], y1 = a.y1[i];
    float kv[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) kv[j] = (j < nk) ? a.k[j][i] : 0.0f;
    float f1 = kv[0];
#pragma unroll
    for (int j = 1; j < 7; ++j) f1 = (j == nk - 1) ? kv[j] : f1;
    if (!is_finite_f(y0)) nonfinite += 1.0;
    float err = __fmul_rn(kv[0], a.ce[0]);
#pragma unroll
    for (int j = 1; j < 7; ++j) err = (j < nk) ? fmaf(kv[j], a.ce[j], err) : err;
    const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(y0), fabsf(y1))));
    const float q = __fdiv_rn(err, tol);
    ex += (double)q * q;
    if (a.final) {
      float mid = __fmul_rn(kv[0], a.cm[0]);
#pragma unroll
      for (int j = 1; j < 7; ++j) mid = (j < nk) ? fmaf(kv[j], a.cm[j], mid) : mid;
      a.y_out[i] = dense_output(y0, y1, __fadd_rn(y0, mid), kv[0], f1, a.dt, a.x_interp);
    }
  }
  if (a.lp0) {
    for (int64_t s = i0; s < a.batch; s += step) {
      const float l0 = a.lp0[s];
      float kl[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) kl[j] = (j < nk) ? a.dlp[j][s] : 0.0f;
      float fl1 = kl[0];
#pragma unroll
      for (int j = 1; j < 7; ++j) fl1 = (j == nk - 1) ? kl[j] : fl1;
      if (!is_finite_f(l0)) nonfinite += 1.0;
      float acc = __fmul_rn(kl[0], a.cl[0]);
#pragma unroll
      for (int j = 1; j < 6; ++j) acc = (j < nk) ? fmaf(kl[j], a.cl[j], acc) : acc;
      const float l1 = __fadd_rn(l0, acc);
      float err = __fmul_rn(kl[0], a.ce[0]);
#pragma unroll
      for (int j = 1; j < 7; ++j) err = (j < nk) ? fmaf(kl[j], a.ce[j], err) : err;
      const float tol = __fadd_rn(a.atol, __fmul_rn(a.rtol, fmaxf(fabsf(l0), fabsf(l1))));
      const float q = __fdiv_rn(err, tol);
      el += (double)q * q;
      a.lp1[s] = l1;
      if (a.final) {
        float mid = __fmul_rn(kl[0], a.cm[0]);
#pragma unroll
        for (int j = 1; j < 7; ++j) mid = (j < nk) ? fmaf(kl[j], a.cm[j], mid) : mid;
        a.lp_out[s] = dense_output(l0, l1, __fadd_rn(l0, mid), kl[0], fl1, a.dt, a.x_interp);
      }
    }
  }
  const double sx = st_block_sum(ex, red), sl = st_block_sum(el, red), sn = st_block_sum(nonfinite, red);
  if (threadIdx.x == 0) {
    double* out = a.partials + (int64_t)blockIdx.x * FFB_NPART;
    out[P_X_ERR] = sx; out[P_LP_ERR] = sl; out[P_NONFINITE] = sn;
  }
}

// =============================================================================================
// C ABI
// =============================================================================================
static int trace_args_ok(const ffb_trace_args* a, const char* who) {
  if (!a || !a->jac || !a->S || !a->dlp) return ffb_fail(FFB_ERR_ARG, std::string(who) + ": jac, S and dlp are required");
  if (a->kind != FFB_TRACE_HUTCHPP && a->kind != FFB_TRACE_XTRACE) return ffb_fail(FFB_ERR_ARG, std::string(who) + ": unknown kind");
  if (a->dim < 1 || a->dim > FFB_TRACE_MAX_DIM) return ffb_fail(FFB_ERR_ARG, std::string(who) + ": dim must be 1..FFB_TRACE_MAX_DIM");
  if (a->rank < 1 || a->rank > FFB_TRACE_MAX_RANK || a->rank > a->dim)
    return ffb_fail(FFB_ERR_ARG, std::string(who) + ": rank must be 1..min(dim, FFB_TRACE_MAX_RANK)");
  if (a->kind == FFB_TRACE_HUTCHPP && (a->nvec < 1 || !a->G)) return ffb_fail(FFB_ERR_ARG, std::string(who) + ": Hutch++ needs G and nvec >= 1");
  return FFB_OK;
}

static int staged_grid(int64_t work_items, int threads) {
  const int64_t blocks = (work_items + threads - 1) / threads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, std::min(FFB_STAGED_BLOCKS, 8 * ffb_num_sms())));
}

// FFB_TRACE_KERNEL=thread : the thread-per-sample kernel (same statements as the CPU twin), kept for A/B.
// Default: the cooperative kernel.  (The STAGE = true variant of the thread kernel is no longer instantiated.)
static int trace_kernel_choice() {
  static const int c = [] {
    const char* e = getenv("FFB_TRACE_KERNEL");
    if (!e) return 0;
    return !strcmp(e, "thread") ? 1 : 0;
  }();
  return c;
}
template <int DMAX, int KMAX>
static int launch_trace(const ffb_trace_args* a, cudaStream_t stream) {
  const int choice = trace_kernel_choice();
  if constexpr (DMAX <= 32) {
   if (choice == 0) {
    constexpr int GPB = 128 / DMAX;
    const int64_t blocks = (a->batch + GPB - 1) / GPB;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, FFB_STAGED_BLOCKS));
    k_trace_coop<DMAX, KMAX><<<grid, 128, 0, stream>>>(*a, a->batch);
    ffb_count_launches(1);
    CUDA_TRY(cudaGetLastError());
    return FFB_OK;
   }
  }
  {
    // thread per sample: the A/B twin of the cooperative kernel for D <= 32, and THE kernel for 32 < D <= 124 (a row of
    // the Jacobian no longer fits the lanes of a warp; the work arrays of the per-sample algebra live in local memory)
    const int threads = 128;
    const int64_t ntiles = (a->batch + threads - 1) / threads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, FFB_STAGED_BLOCKS));
    k_trace_estimate<DMAX, KMAX, false><<<grid, threads, 0, stream>>>(*a, ntiles);
  }
  ffb_count_launches(1);
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}
template <int DMAX, int KMAX>
static void host_trace(const ffb_trace_args* a) {
  const int DD = a->dim * a->dim;
  for (int64_t b = 0; b < a->batch; ++b) a->dlp[b] = ffb::trace_estimate_one<DMAX, KMAX>(*a, a->jac + b * DD, b);
}
// work-array sizes: DMAX in {16, 32}, KMAX in {1, 2, 4, 8}; DMAX in {64, 128}, KMAX in {2, 8} (fewer instantiations: build time)
#define TRACE_DISPATCH_K(CALL, DM)                                                    \
  do { if (kk == 1) CALL(DM, 1); else if (kk == 2) CALL(DM, 2); else if (kk == 4) CALL(DM, 4); else CALL(DM, 8); } while (0)
#define TRACE_DISPATCH(CALL)                                                          \
  do {                                                                                \
    const int kk = a->rank <= 1 ? 1 : (a->rank <= 2 ? 2 : (a->rank <= 4 ? 4 : 8));    \
    if (a->dim <= 16) TRACE_DISPATCH_K(CALL, 16);                                     \
    else if (a->dim <= 32) TRACE_DISPATCH_K(CALL, 32);                                \
    else if (a->dim <= 64) { if (kk <= 2) CALL(64, 2); else CALL(64, 8); }            \
    else { if (kk <= 2) CALL(128, 2); else CALL(128, 8); }                            \
  } while (0)

extern "C" int ffb_trace_estimate(const ffb_trace_args* a, void* stream) {
  if (int rc = trace_args_ok(a, "ffb_trace_estimate")) return rc;
  if (a->norms && (!a->partials || (a->norms == 2 && !a->dlpbase)))
    return ffb_fail(FFB_ERR_ARG, "ffb_trace_estimate: norms need partials (and dlpbase for norms = 2)");
  if (a->batch <= 0) return FFB_OK;
#define TRACE_LAUNCH(DM, KK) return launch_trace<DM, KK>(a, reinterpret_cast<cudaStream_t>(stream))
  TRACE_DISPATCH(TRACE_LAUNCH);
#undef TRACE_LAUNCH
  return FFB_OK;
}

extern "C" int ffb_trace_estimate_host(const ffb_trace_args* a) {
  if (int rc = trace_args_ok(a, "ffb_trace_estimate_host")) return rc;
#define TRACE_HOST(DM, KK) host_trace<DM, KK>(a)
  TRACE_DISPATCH(TRACE_HOST);
#undef TRACE_HOST
  return FFB_OK;
}

extern "C" int ffb_rk_combine(const ffb_rk_combine_args* a, void* stream) {
  if (!a || !a->y0 || !a->out || a->n_terms < 1 || a->n_terms > 7) return ffb_fail(FFB_ERR_ARG, "ffb_rk_combine: bad arguments");
  ffb_rk_combine_args c = *a;
  for (int j = 0; j < 7; ++j) {
    if (j < c.n_terms && !c.k[j]) return ffb_fail(FFB_ERR_ARG, "ffb_rk_combine: missing stage derivative");
    if (j >= c.n_terms) { c.k[j] = c.k[0]; c.coef[j] = 0.0f; }
  }
  if (c.n <= 0) return FFB_OK;
  k_rk_combine<<<staged_grid(c.n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(c);
  ffb_count_launches(1);
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

extern "C" int ffb_rk_finish(const ffb_rk_finish_args* a, void* stream) {
  if (!a || !a->y0 || !a->y1 || !a->partials) return ffb_fail(FFB_ERR_ARG, "ffb_rk_finish: y0, y1 and partials are required");
  if (a->n_k != 0 && (a->n_k < 2 || a->n_k > 7)) return ffb_fail(FFB_ERR_ARG, "ffb_rk_finish: n_k must be 0 (= 7) or 2..7");
  for (int j = 0; j < (a->n_k ? a->n_k : 7); ++j)
    if (!a->k[j] || (a->lp0 && !a->dlp[j])) return ffb_fail(FFB_ERR_ARG, "ffb_rk_finish: missing stage derivative");
  if (a->lp0 && !a->lp1) return ffb_fail(FFB_ERR_ARG, "ffb_rk_finish: lp1 is required with lp0");
  if (a->final && (!a->y_out || (a->lp0 && !a->lp_out))) return ffb_fail(FFB_ERR_ARG, "ffb_rk_finish: final needs y_out (and lp_out)");
  if (a->batch <= 0) return FFB_OK;
  k_rk_finish<<<staged_grid(a->batch * a->dim, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
  ffb_count_launches(1);
  CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

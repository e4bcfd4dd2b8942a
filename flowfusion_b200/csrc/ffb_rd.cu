// ffb_rd.cu -- translation unit of the dual-tile tensor-core engine (ffb_engine_rd.cuh, ffb_kernels_rd.cuh):
// kernel instantiations and their launchers.  The C ABI entry points live in ffb_kernels.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <string>

#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"
#include "ffb_engine_tc.cuh"
#include "ffb_control.cuh"
#include "ffb_kernels_rd.cuh"
#include "ffb_rd.h"

namespace ffb {

#define RD_CUDA_TRY(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ffb_fail(FFB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

static int rd_smem_optin() {
  static int v = 0;
  if (!v) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); }
  return v;
}
static bool rd_gen(const FieldDev& fd) {
  for (int c = 0; c < fd.n_calls; ++c) if (fd.net[c].act != FFB_ACT_SILU) return true;
  return false;
}
// smallest MEM mode (most state in shared memory) whose block fits; -1 when the stage input and the conditional do not
// fit beside the two A_lo images (such fields stay on the single-tile engine)
static int rd_plan(const FieldDev& fd, int nslot, int nbeff, size_t* smem) {
  for (int mem = 0; mem < 2; ++mem) {
    const size_t s = smem_layout_rd(fd.state_dim, fd.cond_dim, rd_field_ka(fd), rd_field_maxl(fd), mem, nslot, fd.n_calls, nbeff,
                                    nullptr, nullptr);
    if (s <= (size_t)rd_smem_optin()) { *smem = s; return mem; }
  }
  return -1;
}

size_t rd_scratch_bytes(int state_dim, int cond_dim) {
  return (size_t)ffb_num_sms() * RD_NGROUP * rd_scratch_floats(state_dim, cond_dim) * sizeof(float);
}

// CTAs of `kern` that can be co-resident (one per SM; with clusters: whole clusters only), cached per kernel
template <typename Kern>
static int rd_max_ctas(Kern kern, size_t smem) {
  static Kern known[64];
  static int value[64];
  static int n_known = 0;
  for (int i = 0; i < n_known; ++i) if (known[i] == kern) return value[i];
  int v = ffb_num_sms();
#if RD_CLUSTER > 1
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ffb_num_sms() / RD_CLUSTER * RD_CLUSTER), 1, 1);
  cfg.blockDim = dim3(RD_NTHR, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RD_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess && nc > 0) v = std::min(v, nc * RD_CLUSTER);
  else { cudaGetLastError(); v = v / RD_CLUSTER * RD_CLUSTER; }
#else
  (void)smem;
#endif
  if (n_known < 64) { known[n_known] = kern; value[n_known] = v; ++n_known; }
  return v;
}

template <typename Kern, typename Args>
static int rd_launch(Kern kern, size_t smem, const FieldDev& fd, const Args& a, int64_t batch, cudaStream_t stream) {
  RD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (batch + TM - 1) / TM;
  if (ntiles <= 0) return FFB_OK;
  // one CTA per two tiles, whole clusters, no more than fit the device at once (the kernels are persistent)
  const int64_t want = ((ntiles + 1) / 2 + RD_CLUSTER - 1) / RD_CLUSTER * RD_CLUSTER;
  const int grid = (int)std::min<int64_t>(want, rd_max_ctas(kern, smem));
  kern<<<grid, RD_NTHR, smem, stream>>>(fd, a, ntiles);
  ffb_count_launches(1);
  RD_CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

#define RD_BY_MEM(mem, EXPR)                                  \
  switch (mem) {                                              \
    case 0: { constexpr int MEM_ = 0; return EXPR; }          \
    default: { constexpr int MEM_ = 1; return EXPR; }         \
  }

bool rd_dopri5_fits(const FieldDev& fd) {
  size_t smem = 0;
  return rd_plan(fd, RD_SCR_SLOTS, 6, &smem) >= 0;
}

int rd_launch_dopri5(const FieldDev& fd, const ffb_dopri5_args& a, cudaStream_t st) {
  size_t smem = 0;
  const int mem = rd_plan(fd, RD_SCR_SLOTS, 6, &smem);
  if (mem < 0) return ffb_fail(FFB_ERR_ARG, "ffb_dopri5_attempt: the field does not fit the dual-tile engine's shared memory");
  const bool dyn = a.ctl != nullptr;
  if (rd_gen(fd)) {
    if (dyn) { RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, true, true>, smem, fd, a, a.batch, st))) }
    RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, true, false>, smem, fd, a, a.batch, st)))
  }
  if (dyn) { RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, false, true>, smem, fd, a, a.batch, st))) }
  RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, false, false>, smem, fd, a, a.batch, st)))
}

}  // namespace ffb

// debug: timeline trace buffer of the dual-tile kernels (5 roles x 2048 x 2 int64; library built with -DFFB_TRACE) or NULL
extern "C" int ffb_debug_trace_rd(long long* buf) {
  cudaMemcpyToSymbol(ffb::g_trace, &buf, sizeof(buf));
  return 0;
}

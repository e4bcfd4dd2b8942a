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
// smallest MEM mode (most state in shared memory) whose block fits; -1 when none does
static int rd_plan(const FieldDev& fd, int nslot, int nbeff, size_t* smem) {
  for (int mem = 0; mem < 3; ++mem) {
    const size_t s = smem_layout_rd(fd.state_dim, fd.cond_dim, rd_field_ka(fd), rd_field_maxl(fd), mem, nslot, fd.n_calls, nbeff,
                                    nullptr, nullptr);
    if (s <= (size_t)rd_smem_optin()) { *smem = s; return mem; }
  }
  return -1;
}

size_t rd_scratch_bytes(int state_dim, int cond_dim) {
  return (size_t)ffb_num_sms() * RD_NGROUP * rd_scratch_floats(state_dim, cond_dim) * sizeof(float);
}

template <typename Kern, typename Args>
static int rd_launch(Kern kern, size_t smem, const FieldDev& fd, const Args& a, int64_t batch, cudaStream_t stream) {
  RD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (batch + TM - 1) / TM;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>((ntiles + 1) / 2, ffb_num_sms());
  kern<<<grid, RD_NTHR, smem, stream>>>(fd, a, ntiles);
  ffb_count_launches(1);
  RD_CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

#define RD_BY_MEM(mem, EXPR)                                  \
  switch (mem) {                                              \
    case 0: { constexpr int MEM_ = 0; return EXPR; }          \
    case 1: { constexpr int MEM_ = 1; return EXPR; }          \
    default: { constexpr int MEM_ = 2; return EXPR; }         \
  }

int rd_launch_eval(const FieldDev& fd, const ffb_eval_args& a, cudaStream_t st) {
  size_t smem = 0;
  const int mem = rd_plan(fd, 3, 1, &smem);
  if (mem < 0) return ffb_fail(FFB_ERR_ARG, "ffb_field_eval: the field does not fit the dual-tile engine's shared memory");
  if (rd_gen(fd)) { RD_BY_MEM(mem, (rd_launch(k_field_eval_rd<MEM_, true>, smem, fd, a, a.batch, st))) }
  RD_BY_MEM(mem, (rd_launch(k_field_eval_rd<MEM_, false>, smem, fd, a, a.batch, st)))
}

int rd_launch_dopri5(const FieldDev& fd, const ffb_dopri5_args& a, cudaStream_t st) {
  size_t smem = 0;
  const int mem = rd_plan(fd, NSLOT, 6, &smem);
  if (mem < 0) return ffb_fail(FFB_ERR_ARG, "ffb_dopri5_attempt: the field does not fit the dual-tile engine's shared memory");
  const bool dyn = a.ctl != nullptr;
  if (rd_gen(fd)) {
    if (dyn) { RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, true, true>, smem, fd, a, a.batch, st))) }
    RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, true, false>, smem, fd, a, a.batch, st)))
  }
  if (dyn) { RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, false, true>, smem, fd, a, a.batch, st))) }
  RD_BY_MEM(mem, (rd_launch(k_dopri5_rd<MEM_, false, false>, smem, fd, a, a.batch, st)))
}

int rd_launch_fixed(const FieldDev& fd, const ffb_fixed_args& a, cudaStream_t st) {
  size_t smem = 0;
  const int mem = rd_plan(fd, rr_fixed_slots(a.method), 16, &smem);
  if (mem < 0) return ffb_fail(FFB_ERR_ARG, "ffb_integrate_fixed: the field does not fit the dual-tile engine's shared memory");
  if (rd_gen(fd)) { RD_BY_MEM(mem, (rd_launch(k_fixed_rd<MEM_, true>, smem, fd, a, a.batch, st))) }
  RD_BY_MEM(mem, (rd_launch(k_fixed_rd<MEM_, false>, smem, fd, a, a.batch, st)))
}

}  // namespace ffb

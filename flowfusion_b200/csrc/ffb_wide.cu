// ffb_wide.cu -- translation unit of the wide engine (ffb_engine_wide.cuh): the engine-generic kernels instantiated on
// EngineWide, and their launchers.  The C ABI entry points live in ffb_kernels.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <string>

#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"
#include "ffb_engine_wide.cuh"
#include "ffb_kernels_generic.cuh"
#include "ffb_wide.h"

namespace ffb {

#define WD_CUDA_TRY(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ffb_fail(FFB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

static int wd_smem_optin() {
  static int v = 0;
  if (!v) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); }
  return v;
}
static int wd_tangents(const FieldDev& fd) {
  return fd.div_mode == FFB_DIV_EXACT ? fd.net[0].x_dim : (fd.div_mode == FFB_DIV_HUTCH ? 1 : 0);
}
// shared-memory block of the field for H halves per pass; picks fd.slots_smem
static size_t wd_pick_smem(FieldDev* fd, int H) {
  const int T = wd_tangents(*fd), hutch = fd->div_mode == FFB_DIV_HUTCH;
  const size_t with_slots = smem_layout_wide(fd->state_dim, fd->cond_dim, T, hutch, 1, fd->wide_maxk, nullptr, H);
  fd->slots_smem = (with_slots <= (size_t)wd_smem_optin()) ? 1 : 0;
  return fd->slots_smem ? with_slots : smem_layout_wide(fd->state_dim, fd->cond_dim, T, hutch, 0, fd->wide_maxk, nullptr, H);
}
// two 32-row halves per pass (8 rows per thread) when the doubled activation buffers still fit, else one
// (FFB_WIDE_HALVES=1 forces one: A/B runs)
static int wd_halves(const FieldDev& fd) {
  static const int forced = [] { const char* e = getenv("FFB_WIDE_HALVES"); return e ? atoi(e) : 0; }();
  if (forced == 1) return 1;
  for (int c = 0; c < fd.n_calls; ++c) if (fd.net[c].act != FFB_ACT_SILU) return 1;      // the two-half kernels are SiLU-only
  FieldDev tmp = fd;
  return wd_pick_smem(&tmp, 2) <= (size_t)wd_smem_optin() ? 2 : 1;
}

template <typename Kern, typename Args>
static int wd_launch(Kern kern, size_t smem, const char* name, const FieldDev& fd, const Args& a, int64_t batch, cudaStream_t stream) {
  if ((int)smem > wd_smem_optin())
    return ffb_fail(FFB_ERR_ARG, std::string(name) + ": the wide engine needs " + std::to_string(smem) +
                    " B of shared memory for this field (widest layer x 32 rows x 2 + state), the device allows " +
                    std::to_string(wd_smem_optin()));
  WD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int S = TM / (1 + wd_tangents(fd));
  const int64_t ntiles = (batch + S - 1) / S;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, ffb_num_sms());
  kern<<<grid, EngineWide::NTHR, smem, stream>>>(fd, a, ntiles);
  ffb_count_launches(1);
  WD_CUDA_TRY(cudaGetLastError());
  return FFB_OK;
}

// The two-half kernels live in their own translation unit (ffb_wide2.cu) so that the two compile in parallel.
#define WD_DISPATCH_ONE(KERNEL, NAME, BATCH)                                                                    \
  const size_t smem = wd_pick_smem(&fd, 1);                                                                     \
  if (fd.slots_smem) return wd_launch(KERNEL<EngineWideT<1, true>, true>, smem, NAME, fd, a, BATCH, st);        \
  return wd_launch(KERNEL<EngineWideT<1, true>, false>, smem, NAME, fd, a, BATCH, st)

// (the single evaluation keeps one half per pass: it is a small share of a solve, and every instantiation costs build time)
int wide_launch_field_eval(FieldDev fd, const ffb_eval_args& a, cudaStream_t st) { WD_DISPATCH_ONE(k_field_eval, "ffb_field_eval", a.batch); }
int wide_launch_dopri5(FieldDev fd, const ffb_dopri5_args& a, cudaStream_t st) {
  if (wd_halves(fd) == 2) { const size_t smem = wd_pick_smem(&fd, 2); return wide2_launch_dopri5(fd, a, smem, st); }
  WD_DISPATCH_ONE(k_dopri5, "ffb_dopri5_attempt", a.batch);
}
int wide_launch_fixed(FieldDev fd, const ffb_fixed_args& a, cudaStream_t st) {
  if (wd_halves(fd) == 2) { const size_t smem = wd_pick_smem(&fd, 2); return wide2_launch_fixed(fd, a, smem, st); }
  WD_DISPATCH_ONE(k_fixed, "ffb_integrate_fixed", a.batch);
}

}  // namespace ffb

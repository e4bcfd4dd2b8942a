// ffb_wide2.cu -- the wide engine's kernels with two 32-row halves per pass (EngineWideT<2, false>: 8 rows per thread, SiLU
// networks whose doubled activation buffers fit shared memory, i.e. widths up to 256), in their own translation unit so that
// they compile in parallel with ffb_wide.cu.  The choice between the variants is made in ffb_wide.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <string>

#include "ffb200.h"
#include "ffb_common.cuh"
#include "ffb_engine.cuh"
#include "ffb_engine_wide.cuh"
#include "ffb_kernels_generic.cuh"
#include "ffb_wide.h"

namespace ffb {

template <typename Kern, typename Args>
static int wd2_launch(Kern kern, size_t smem, const char* name, const FieldDev& fd, const Args& a, int64_t batch, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return ffb_fail(FFB_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(e));
  const int T = fd.div_mode == FFB_DIV_EXACT ? fd.net[0].x_dim : (fd.div_mode == FFB_DIV_HUTCH ? 1 : 0);
  const int S = TM / (1 + T);
  const int64_t ntiles = (batch + S - 1) / S;
  if (ntiles <= 0) return FFB_OK;
  const int grid = (int)std::min<int64_t>(ntiles, ffb_num_sms());
  kern<<<grid, EngineWide::NTHR, smem, stream>>>(fd, a, ntiles);
  ffb_count_launches(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return ffb_fail(FFB_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(e));
  return FFB_OK;
}

int wide2_launch_dopri5(const FieldDev& fd, const ffb_dopri5_args& a, size_t smem, cudaStream_t st) {
  if (fd.slots_smem) return wd2_launch(k_dopri5<EngineWideT<2, false>, true>, smem, "ffb_dopri5_attempt", fd, a, a.batch, st);
  return wd2_launch(k_dopri5<EngineWideT<2, false>, false>, smem, "ffb_dopri5_attempt", fd, a, a.batch, st);
}
int wide2_launch_fixed(const FieldDev& fd, const ffb_fixed_args& a, size_t smem, cudaStream_t st) {
  if (fd.slots_smem) return wd2_launch(k_fixed<EngineWideT<2, false>, true>, smem, "ffb_integrate_fixed", fd, a, a.batch, st);
  return wd2_launch(k_fixed<EngineWideT<2, false>, false>, smem, "ffb_integrate_fixed", fd, a, a.batch, st);
}

}  // namespace ffb

// fft_reg_k1_f64.cu -- instantiations of stage_reg_kernel (fft_reg_kernel.h) for one kind and precision
// (one translation unit each so that they compile in parallel)
#include <algorithm>

#include "fft_reg_kernel.h"

namespace pfb {
cudaError_t launch_reg_k1_f64(StageParams &sp, cudaStream_t stream) { return launch_reg_kind<double, 1>(sp, stream); }
}  // namespace pfb

// fft_pow2.cu -- register-resident power-of-two stage kernel (placeholder: not enabled yet).
#include "kernels.h"
namespace pfb {
template <typename T> bool pow2_supported(const Stage &, int) { return false; }
template <typename T> int pow2_pick_tile(const Stage &, int) { return 0; }
template <typename T> cudaError_t launch_stage_pow2(StageParams &, cudaStream_t) { return cudaErrorNotSupported; }
template bool pow2_supported<float>(const Stage &, int);
template bool pow2_supported<double>(const Stage &, int);
template int pow2_pick_tile<float>(const Stage &, int);
template int pow2_pick_tile<double>(const Stage &, int);
template cudaError_t launch_stage_pow2<float>(StageParams &, cudaStream_t);
template cudaError_t launch_stage_pow2<double>(StageParams &, cudaStream_t);
}

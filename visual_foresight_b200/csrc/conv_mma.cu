// conv_mma.cu — tcgen05 implicit-GEMM convolution (under construction: reports unsupported so the
// engine keeps every layer on the fp32 FFMA kernels).
#include "conv_mma.cuh"

namespace vf {
bool mma_conv_supported(int, int, int, int, int) { return false; }
int mma_conv_prepare_weights(const float*, int, int, int, MmaConvWeights*, std::vector<void*>*, std::string* err) {
  if (err) *err = "tcgen05 conv not built";
  return -1;
}
int mma_conv_launch(const MmaConvWeights&, const MmaConvCall&, int, cudaStream_t) { return -1; }
}  // namespace vf

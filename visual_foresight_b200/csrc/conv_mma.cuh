// conv_mma.cuh — tcgen05 (5th-gen tensor core) implicit-GEMM convolution: host-side interface.
#pragma once
#include <string>
#include <vector>

#include "vf_common.cuh"

namespace vf {

// fp16 hi/lo split copies of one conv layer's spatial weights, laid out for the UMMA B operand
struct MmaConvWeights {
  bool ready = false;
  int k = 0, cin = 0, cout = 0;
  int kw = 0;    // filter columns multiplied by the MMAs: k, or 1 when the dx taps are folded into the input channels
  int kcl = 0;   // nominal kernel size (border classes of the action/state bias)
  __half* w_hi = nullptr;   // [k*k][cout][cin]  (K-major rows), scaled by 2^scale_log2
  __half* w_lo = nullptr;
  int scale_log2 = 0;
};

struct MmaConvCall {
  View src, out;     // src: split-half storage (vf_common.cuh); channels [0, src.C)
  View src1 = make_view(nullptr, 0, 0, 0, 0);   // optional second source: channels [src.C, src.C + src1.C), src.C % 32 == 0
  const float* sabias;
  const float* bias;
  int H, W;
  int passes;   // 3 = hi*hi + lo*hi + hi*lo (fp32-grade), 1 = hi*hi
  int act = 0;  // ACT_NONE / ACT_SIGMOID applied in the epilogue
  double* stats_partial = nullptr;   // in: scratch for fused instance-norm partial sums (null = not requested)
  int* stats_slots = nullptr;        // out: partial slots per (sample, channel) written (0 = this launch did not fuse them)
  float* stats_fin = nullptr;        // in (optional): (mean, rstd) pairs written by the last-arriving warp of every sample
  int* stats_cnt = nullptr;          // in: zeroed arrival counters [B][VF_STAT_CNT_STRIDE] (left zeroed)
  float stats_eps = 1e-6f;
  bool* stats_finalized = nullptr;   // out: the launch also finalised the statistics into stats_fin
};

bool mma_conv_supported(int k, int cin, int cout, int H, int W);
// partial slots per (sample, channel) of the fused instance-norm statistics for this layer shape (0 = not fused)
int mma_conv_stats_slots(int k, int kw, int cin, int cout, int H, int W);
// host-only: the tiling picked for a layer as 24 integers (see vf_debug_conv_plan in include/vfengine.h); false = unsupported
bool mma_conv_describe(int k, int kw, int cin, int cout, int H, int W, int B, int passes, int out[24]);
// returns 0 on success; device allocations are appended to *allocs (owned by the engine handle)
// w_sp: [k * kw][cin][cout]
int mma_conv_prepare_weights(const float* w_sp, int k, int kw, int kcl, int cin, int cout, MmaConvWeights* out,
                             std::vector<void*>* allocs, std::string* err);
int mma_conv_launch(const MmaConvWeights& w, const MmaConvCall& c, int B, cudaStream_t s);

}  // namespace vf

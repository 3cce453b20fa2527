// conv_simt.cu — fp32 FFMA convolution kernels (SAME zero padding, stride 1, NHWC).
//
// Role: (1) the layers that are too thin for the tensor-core path (first encoder conv with 6 input
// channels, the 3- and 7-channel head convs), (2) the fp32 checker the tcgen05 implicit-GEMM kernel
// is unit-tested against.  The tiled action/state channels of the reference graph (tile_concat,
// spec P2) never appear as input channels here: they are a per-sample bias that depends only on
// the border class of the output pixel (sabias), added in the epilogue.
#include "vf_common.cuh"

namespace vf {

long long g_launch_counter = 0;
bool g_use_pdl = false;   // set by vf_create from VF_PDL=1 (off by default: see engine.cu)

namespace {

__device__ __forceinline__ float load_src(const ConvArgs& a, int b, int y, int x, int c) {
  if (c < a.src0.C) return vld1(a.src0, voff(a.src0, b, (long long)(y * a.W + x)) + c);
  c -= a.src0.C;
  return vld1(a.src1, voff(a.src1, b, (long long)(y * a.W + x)) + c);
}

// Tile kernel: block = TH x TW output pixels x TN output channels of one sample; 256 threads, each
// 4 consecutive pixels (along x) x 4 consecutive output channels.  K loop over chunks of CI input
// channels; per chunk the (TH+k-1)x(TW+k-1) halo patch and the k*k*CI*TN weights are staged in smem.
template <int KS, int TN, int TH, int TW>
__global__ void __launch_bounds__(256) k_conv_tile(ConvArgs a) {
  constexpr int CI = 4;
  constexpr int PAD = KS / 2;
  constexpr int PH = TH + KS - 1, PW = TW + KS - 1;
  constexpr int NL = TN / 4;             // n-lanes
  constexpr int CG = TW / 4;             // column groups per row
  static_assert(NL * TH * CG == 256, "thread mapping");
  __shared__ float s_in[CI][PH][PW + 1];
  __shared__ __align__(16) float s_w[KS * KS][CI][TN];

  const int tid = threadIdx.x;
  const int nl = tid % NL;
  const int pl = tid / NL;
  const int r = pl / CG, cg = pl % CG;
  const int tiles_x = (a.W + TW - 1) / TW;
  const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
  const int n0 = blockIdx.y * TN;
  const int b = blockIdx.z;

  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;

  for (int c0 = 0; c0 < a.Cin; c0 += CI) {
    __syncthreads();
    for (int idx = tid; idx < PH * PW * CI; idx += 256) {
      const int ci = idx % CI;
      const int px = (idx / CI) % PW;
      const int py = idx / (CI * PW);
      const int y = ty0 + py - PAD, x = tx0 + px - PAD, c = c0 + ci;
      float v = 0.f;
      if (y >= 0 && y < a.H && x >= 0 && x < a.W && c < a.Cin) v = load_src(a, b, y, x, c);
      s_in[ci][py][px] = v;
    }
    for (int idx = tid; idx < KS * KS * CI * TN; idx += 256) {
      const int n = idx % TN;
      const int ci = (idx / TN) % CI;
      const int tap = idx / (TN * CI);
      const int c = c0 + ci;
      float v = 0.f;
      if (c < a.Cin && n0 + n < a.Cout) v = __ldg(a.w + ((long long)tap * a.Cin + c) * a.Cout + n0 + n);
      s_w[tap][ci][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
        float in[4 + KS - 1];
#pragma unroll
        for (int i = 0; i < 4 + KS - 1; ++i) in[i] = s_in[ci][r + dy][cg * 4 + i];
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float4 w = *reinterpret_cast<const float4*>(&s_w[dy * KS + dx][ci][nl * 4]);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            acc[p][0] = fmaf(in[p + dx], w.x, acc[p][0]);
            acc[p][1] = fmaf(in[p + dx], w.y, acc[p][1]);
            acc[p][2] = fmaf(in[p + dx], w.z, acc[p][2]);
            acc[p][3] = fmaf(in[p + dx], w.w, acc[p][3]);
          }
        }
      }
    }
  }

  const int y = ty0 + r;
  const int n = n0 + nl * 4;
  if (y >= a.H || n >= a.Cout) return;
  const int cy = border_class(y, a.H, PAD);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int x = tx0 + cg * 4 + p;
    if (x >= a.W) continue;
    float add[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.sabias) {
      const int cls = cy * KS + border_class(x, a.W, PAD);
      const float* sb = a.sabias + ((long long)b * KS * KS + cls) * a.Cout + n;
#pragma unroll
      for (int q = 0; q < 4; ++q) add[q] = (n + q < a.Cout) ? sb[q] : 0.f;
    } else if (a.bias) {
#pragma unroll
      for (int q = 0; q < 4; ++q) add[q] = (n + q < a.Cout) ? a.bias[n + q] : 0.f;
    }
    const long long o = voff(a.out, b, (long long)(y * a.W + x)) + n;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (n + q < a.Cout) vst1(a.out, o + q, acc[p][q] + add[q]);
  }
}

// Thin-output kernel (Cout <= 8): one thread per output pixel, all output channels in registers,
// weights in shared memory.  Used for scratch.conv1 (32->3, sigmoid) and masks.conv1 (53->7).
template <int KS>
__global__ void __launch_bounds__(128) k_conv_small(ConvArgs a) {
  constexpr int PAD = KS / 2;
  extern __shared__ float s_w[];   // [k*k][Cin][Cout]
  const int b = blockIdx.y;
  const int nw = KS * KS * a.Cin * a.Cout;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = __ldg(a.w + i);
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= a.H * a.W) return;
  const int y = pix / a.W, x = pix % a.W;
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  for (int dy = 0; dy < KS; ++dy) {
    const int yy = y + dy - PAD;
    if (yy < 0 || yy >= a.H) continue;
    for (int dx = 0; dx < KS; ++dx) {
      const int xx = x + dx - PAD;
      if (xx < 0 || xx >= a.W) continue;
      const float* wt = s_w + (dy * KS + dx) * a.Cin * a.Cout;
      for (int c = 0; c < a.Cin; ++c) {
        const float v = load_src(a, b, yy, xx, c);
        const float* wc = wt + c * a.Cout;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q < a.Cout) acc[q] = fmaf(v, wc[q], acc[q]);
      }
    }
  }
  const long long o = voff(a.out, b, pix);
  for (int q = 0; q < a.Cout; ++q) {
    float v = acc[q] + (a.bias ? a.bias[q] : 0.f);
    if (a.act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    vst1(a.out, o + q, v);
  }
}

}  // namespace

void launch_conv_simt(const ConvArgs& a, int B, cudaStream_t s) {
  ++g_launch_counter;
  if (a.Cout <= 8) {
    dim3 grid((a.H * a.W + 127) / 128, B);
    size_t smem = (size_t)a.k * a.k * a.Cin * a.Cout * sizeof(float);
    if (a.k == 3) k_conv_small<3><<<grid, 128, smem, s>>>(a);
    else k_conv_small<5><<<grid, 128, smem, s>>>(a);
    return;
  }
  if (a.Cout % 64 == 0) {
    constexpr int TH = 8, TW = 8, TN = 64;
    dim3 grid(((a.H + TH - 1) / TH) * ((a.W + TW - 1) / TW), (a.Cout + TN - 1) / TN, B);
    if (a.k == 3) k_conv_tile<3, TN, TH, TW><<<grid, 256, 0, s>>>(a);
    else k_conv_tile<5, TN, TH, TW><<<grid, 256, 0, s>>>(a);
  } else {
    constexpr int TH = 8, TW = 16, TN = 32;
    dim3 grid(((a.H + TH - 1) / TH) * ((a.W + TW - 1) / TW), (a.Cout + TN - 1) / TN, B);
    if (a.k == 3) k_conv_tile<3, TN, TH, TW><<<grid, 256, 0, s>>>(a);
    else k_conv_tile<5, TN, TH, TW><<<grid, 256, 0, s>>>(a);
  }
}

}  // namespace vf

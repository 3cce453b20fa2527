// cdna_cost.cu — CDNA kernel head, CDNA application (5x5 per-sample stencil on a SYMMETRIC-padded
// image), softmax mask composite (image + designated-pixel distribution), distribution
// renormalisation and the planning costs.  Every kernel here is HBM/L2-bandwidth bound: one thread
// per pixel, channels innermost, warp reductions in a fixed order (deterministic scores).
#include "vf_common.cuh"

namespace vf {
namespace {

__device__ __forceinline__ const float* vptr(const View& v, int b, long long pix) {
  return v.p + (long long)b * v.sample_stride + pix * v.pix_stride + v.ch_off;
}
// TF 'SYMMETRIC' padding index: -1 -> 0, -2 -> 1, n -> n-1, n+1 -> n-2
__device__ __forceinline__ int mirror(int i, int n) { return i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i); }

// CDNA kernel head (spec P5): dense(flatten(h)) -> +identity -> relu-shift -> L1 normalise.
// Block = CK_S samples x all outputs: every weight element fetched from L2 feeds CK_S FMAs.  1024 threads =
// 128 output lanes x CK_G K-groups; the feature chunk of the CK_S samples is staged in shared memory (broadcast reads).
constexpr int CK_S = 4, CK_G = 8, CK_T = 1024;
__global__ void __launch_bounds__(1024) k_cdna_kernels(View feat, int npix, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int ksize, int nt, int B, float* kern) {
  const int b0 = blockIdx.x * CK_S;
  const int nout = ksize * ksize * nt;          // <= 128
  const int j = threadIdx.x & 127, g = threadIdx.x >> 7;
  const int K = npix * feat.C;
  __shared__ float sf[CK_S][CK_T];
  __shared__ float part[CK_G][CK_S][128];
  float acc[CK_S];
#pragma unroll
  for (int sI = 0; sI < CK_S; ++sI) acc[sI] = 0.f;
  for (int k0 = 0; k0 < K; k0 += CK_T) {
    for (int i = threadIdx.x; i < CK_S * CK_T; i += 1024) {
      const int sI = i / CK_T, kk = i - sI * CK_T, k = k0 + kk, b = b0 + sI;
      sf[sI][kk] = (b < B && k < K) ? __ldg(vptr(feat, b, k / feat.C) + (k % feat.C)) : 0.f;
    }
    __syncthreads();
    if (j < nout) {
      constexpr int KG = CK_T / CK_G;
      const int ke = min(KG, max(0, K - k0 - g * KG));
      const float* wp = w + (long long)(k0 + g * KG) * nout + j;
      for (int kk = 0; kk < ke; ++kk) {
        const float wv = __ldg(wp + (long long)kk * nout);
#pragma unroll
        for (int sI = 0; sI < CK_S; ++sI) acc[sI] = fmaf(sf[sI][g * KG + kk], wv, acc[sI]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int sI = 0; sI < CK_S; ++sI) part[g][sI][j] = acc[sI];
  __syncthreads();
  // raw kernel taps: dense output reshaped (k, k, nt): j = (u*k + v)*nt + n ; + identity at the centre tap
  for (int i = threadIdx.x; i < CK_S * 128; i += 1024) {
    const int sI = i >> 7, jj = i & 127;
    if (jj < nout) {
      float v = bias[jj];
#pragma unroll
      for (int gg = 0; gg < CK_G; ++gg) v += part[gg][sI][jj];      // fixed order
      if (jj / nt == (ksize / 2) * ksize + ksize / 2) v += 1.0f;
      part[0][sI][jj] = fmaxf(v - 1e-12f, 0.f) + 1e-12f;
    }
  }
  __syncthreads();
  if (threadIdx.x < CK_S * nt) {
    const int sI = threadIdx.x / nt, n = threadIdx.x % nt, b = b0 + sI;
    if (b < B) {
      float sum = 0.f;
      for (int t = 0; t < ksize * ksize; ++t) sum += part[0][sI][t * nt + n];
      for (int t = 0; t < ksize * ksize; ++t) kern[((long long)b * nt + n) * ksize * ksize + t] = part[0][sI][t * nt + n] / sum;
    }
  }
}

__global__ void k_cdna_apply(View image, View first, const float* __restrict__ kern, int ksize, int nt, int H, int W,
                             View mask_in, int base) {
  extern __shared__ float sk[];   // [nt][k*k]
  const int b = blockIdx.y;
  const int kk = ksize * ksize;
  for (int i = threadIdx.x; i < nt * kk; i += blockDim.x) sk[i] = kern[(long long)b * nt * kk + i];
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix % W, pad = ksize / 2;
  float acc[8][3];
  for (int n = 0; n < nt; ++n) acc[n][0] = acc[n][1] = acc[n][2] = 0.f;
  for (int u = 0; u < ksize; ++u) {
    const int yy = mirror(y + u - pad, H);
    for (int v = 0; v < ksize; ++v) {
      const int xx = mirror(x + v - pad, W);
      const float* ip = vptr(image, b, (long long)yy * W + xx);
      const float r = __ldg(ip), g = __ldg(ip + 1), bl = __ldg(ip + 2);
      for (int n = 0; n < nt; ++n) {
        const float kv = sk[n * kk + u * ksize + v];
        acc[n][0] = fmaf(r, kv, acc[n][0]);
        acc[n][1] = fmaf(g, kv, acc[n][1]);
        acc[n][2] = fmaf(bl, kv, acc[n][2]);
      }
    }
  }
  float* o = mask_in.p + (long long)b * mask_in.sample_stride + (long long)pix * mask_in.pix_stride + mask_in.ch_off + base;
  for (int n = 0; n < nt; ++n)
    for (int c = 0; c < 3; ++c) o[3 * n + c] = acc[n][c];
  const float* ip = vptr(image, b, pix);
  const float* fp = vptr(first, b, pix);
  for (int c = 0; c < 3; ++c) {
    o[3 * nt + c] = __ldg(ip + c);
    o[3 * nt + 3 + c] = __ldg(fp + c);
  }
}

constexpr int COMP_THREADS = 256;

// one thread per pixel.  masks = softmax(logits); gen_image = sum_n m_n * layer_n;
// gen_distrib = sum_{n<nt} m_n * T_n(prev_d) + m_nt*prev_d + m_{nt+1}*first_d + m_{nt+2}*prev_d   (spec P8)
__global__ void __launch_bounds__(COMP_THREADS) k_composite(CompositeArgs a) {
  extern __shared__ float sk[];
  __shared__ float red[COMP_THREADS / 32];
  const int b = blockIdx.y;
  const int kk = a.ksize * a.ksize, nm = a.nt + 3;
  for (int i = threadIdx.x; i < a.nt * kk; i += blockDim.x) sk[i] = a.kern[(long long)b * a.nt * kk + i];
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = pix < a.H * a.W;
  float m[16];
  float dsum[4] = {0.f, 0.f, 0.f, 0.f};
  if (ok) {
    const int y = pix / a.W, x = pix % a.W, pad = a.ksize / 2;
    const float* lg = vptr(a.logits, b, pix);
    float mx = -3.4e38f;
    for (int n = 0; n < nm; ++n) { m[n] = __ldg(lg + n); mx = fmaxf(mx, m[n]); }
    float se = 0.f;
    for (int n = 0; n < nm; ++n) { m[n] = expf(m[n] - mx); se += m[n]; }
    for (int n = 0; n < nm; ++n) m[n] = m[n] / se;
    const float* ly = vptr(a.layers, b, pix);
    float* gi = a.gen_image.p + (long long)b * a.gen_image.sample_stride + (long long)pix * a.gen_image.pix_stride + a.gen_image.ch_off;
    for (int c = 0; c < 3; ++c) {
      float v = 0.f;
      for (int n = 0; n < nm; ++n) v += m[n] * __ldg(ly + 3 * n + c);
      gi[c] = v;
    }
    float* gd = a.gen_distrib.p + (long long)b * a.gen_distrib.sample_stride + (long long)pix * a.gen_distrib.pix_stride + a.gen_distrib.ch_off;
    for (int p = 0; p < a.nd; ++p) {
      float t[8];
      for (int n = 0; n < a.nt; ++n) t[n] = 0.f;
      for (int u = 0; u < a.ksize; ++u) {
        const int yy = mirror(y + u - pad, a.H);
        for (int v = 0; v < a.ksize; ++v) {
          const int xx = mirror(x + v - pad, a.W);
          const float d = __ldg(vptr(a.prev_d, b, (long long)yy * a.W + xx) + p);
          for (int n = 0; n < a.nt; ++n) t[n] = fmaf(d, sk[n * kk + u * a.ksize + v], t[n]);
        }
      }
      const float pd = __ldg(vptr(a.prev_d, b, pix) + p), fd = __ldg(vptr(a.first_d, b, pix) + p);
      float v = 0.f;
      for (int n = 0; n < a.nt; ++n) v += m[n] * t[n];
      v += m[a.nt] * pd;
      v += m[a.nt + 1] * fd;
      v += m[a.nt + 2] * pd;
      gd[p] = v;
      dsum[p] = v;
    }
  }
  // block sums of the raw distribution (fixed shuffle tree + fixed warp order)
  for (int p = 0; p < a.nd; ++p) {
    float v = dsum[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < COMP_THREADS / 32; ++i) s += red[i];
      a.partial[((long long)b * a.nd + p) * gridDim.x + blockIdx.x] = s;
    }
    __syncthreads();
  }
}

__global__ void k_distrib_normalize(View d, const float* __restrict__ partial, int nblk, int H, int W, int nd) {
  const int b = blockIdx.y;
  __shared__ float inv[4];
  if (threadIdx.x < nd) {
    float s = 0.f;
    for (int i = 0; i < nblk; ++i) s += partial[((long long)b * nd + threadIdx.x) * nblk + i];
    inv[threadIdx.x] = s;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W * nd) return;
  const int p = i % nd, pix = i / nd;
  float* q = d.p + (long long)b * d.sample_stride + (long long)pix * d.pix_stride + d.ch_off + p;
  *q = *q / inv[p];
}

// block per (m, t, cam, p) plane.  Mirrors _expected_distance (pixel_cost_controller.py:168-187):
// the distance grid is float64, the product p*d is formed in float64 and rounded to float32
// (numpy in-place f32 *= f64), sums are float32.
__global__ void __launch_bounds__(256) k_pixel_cost(const float* __restrict__ distrib, int P, int ncam, int H, int W,
                                                    int nd, const double* __restrict__ goal, float* cost) {
  const int plane = blockIdx.x;                 // ((m*P + t)*ncam + cam)*nd + p
  const int p = plane % nd;
  const int cam = (plane / nd) % ncam;
  const long long mt = plane / (nd * ncam);
  const float* base = distrib + ((mt * ncam + cam) * (long long)H * W) * nd + p;
  const double gy = goal[(cam * nd + p) * 2], gx = goal[(cam * nd + p) * 2 + 1];
  float s = 0.f, sd = 0.f;
  for (int i = threadIdx.x; i < H * W; i += 256) {
    const float v = __ldg(base + (long long)i * nd);
    const double dy = gy - (double)(i / W), dx = gx - (double)(i % W);
    const double dist = sqrt(dy * dy + dx * dx);
    s += v;
    sd += (float)((double)v * dist);
  }
  __shared__ float r0[8], r1[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
  }
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s; r1[threadIdx.x >> 5] = sd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += r0[i]; c += r1[i]; }
    const int m = (int)(mt / P), t = (int)(mt % P);
    cost[((long long)m * P + t) * (ncam * nd) + cam * nd + p] = c / a;
  }
}

__global__ void k_score_final(const float* __restrict__ cost, int M, int P, int ntask, const double* __restrict__ tw,
                              double finalweight, double* scores) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const double tsum = (double)(P - 1) + finalweight;
  double acc = 0.0;
  for (int k = 0; k < ntask; ++k) {
    float s = 0.f;                           // float32 time sum like np.sum over a float32 array
    for (int t = 0; t < P; ++t) {
      const float mult = (t == P - 1) ? (float)finalweight : 1.f;
      s += cost[((long long)m * P + t) * ntask + k] * mult;
    }
    acc += tw[k] * ((double)s / tsum);
  }
  scores[m] = acc;
}

__global__ void __launch_bounds__(256) k_goal_image_cost(const float* __restrict__ frames, int P, int ncam, int H, int W,
                                                         const float* __restrict__ goal, double* scores) {
  const int m = blockIdx.x;
  const long long n = (long long)H * W * 3;
  const float* f = frames + (((long long)m * P + (P - 1)) * ncam + 0) * n;
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float d = f[i] - goal[i];
    s = fmaf(d, d, s);
  }
  __shared__ float r[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += r[i];
    scores[m] = (double)(a / (float)n);
  }
}

}  // namespace

void launch_cdna_kernels(View feat, int npix, const float* w, const float* bias, int ksize, int nt, int B, float* kern,
                         cudaStream_t s) {
  ++g_launch_counter;
  k_cdna_kernels<<<(B + CK_S - 1) / CK_S, 1024, 0, s>>>(feat, npix, w, bias, ksize, nt, B, kern);
}
void launch_cdna_apply(View image, View first, const float* kern, int ksize, int nt, int B, int H, int W, View mask_in,
                       int base, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid((H * W + 127) / 128, B);
  k_cdna_apply<<<grid, 128, nt * ksize * ksize * sizeof(float), s>>>(image, first, kern, ksize, nt, H, W, mask_in, base);
}
int composite_blocks(int H, int W) { return (H * W + COMP_THREADS - 1) / COMP_THREADS; }
void launch_composite(const CompositeArgs& a, int B, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid(composite_blocks(a.H, a.W), B);
  k_composite<<<grid, COMP_THREADS, a.nt * a.ksize * a.ksize * sizeof(float), s>>>(a);
}
void launch_distrib_normalize(View d, const float* partial, int nblk, int B, int H, int W, int nd, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid((H * W * nd + 255) / 256, B);
  k_distrib_normalize<<<grid, 256, 0, s>>>(d, partial, nblk, H, W, nd);
}
void launch_pixel_cost(const float* distrib, int M, int P, int ncam, int H, int W, int nd, const double* goal,
                       float* cost, cudaStream_t s) {
  ++g_launch_counter;
  k_pixel_cost<<<M * P * ncam * nd, 256, 0, s>>>(distrib, P, ncam, H, W, nd, goal, cost);
}
void launch_score_final(const float* cost, int M, int P, int ntask, const double* tw, double finalweight, double* scores,
                        cudaStream_t s) {
  ++g_launch_counter;
  k_score_final<<<(M + 127) / 128, 128, 0, s>>>(cost, M, P, ntask, tw, finalweight, scores);
}
void launch_goal_image_cost(const float* frames, int M, int P, int ncam, int H, int W, const float* goal, double* scores,
                            cudaStream_t s) {
  ++g_launch_counter;
  k_goal_image_cost<<<M, 256, 0, s>>>(frames, P, ncam, H, W, goal, scores);
}

}  // namespace vf

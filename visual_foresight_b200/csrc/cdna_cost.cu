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

// ---- CDNA kernel head (spec P5): dense(flatten(h)) -> +identity -> relu-shift -> L1 normalise -----------------------
// Split-K partial products: grid (K slices of CK_KC, sample groups of CK_S) fills the 148 SMs (c2: 16 x 25 blocks); every
// weight element fetched from L2 feeds CK_S FMAs.  part[ks][b][128] is combined in a fixed order by cdna_finalize(), which
// the consumers (CDNA apply) run in their prologue: no separate finalize launch.
constexpr int CK_S = 8, CK_KC = 256;
__global__ void __launch_bounds__(256) k_cdna_partial(View feat, int npix, const float* __restrict__ w, int nout, int B,
                                                      float* __restrict__ part) {
  pdl_wait();
  pdl_trigger();
  const int ks = blockIdx.x, k0 = ks * CK_KC, b0 = blockIdx.y * CK_S;
  const int K = npix * feat.C;
  __shared__ __align__(16) float sf[CK_S][CK_KC];
  __shared__ float red[CK_S][128];
  for (int i = threadIdx.x; i < CK_S * CK_KC; i += 256) {
    const int sI = i / CK_KC, kk = i - sI * CK_KC, k = k0 + kk, b = b0 + sI;
    sf[sI][kk] = (b < B && k < K) ? vld1(feat, voff(feat, b, k / feat.C) + (k % feat.C)) : 0.f;
  }
  __syncthreads();
  const int j = threadIdx.x & 127, g = threadIdx.x >> 7;          // 128 output lanes x 2 K halves
  float acc[CK_S];
#pragma unroll
  for (int sI = 0; sI < CK_S; ++sI) acc[sI] = 0.f;
  if (j < nout) {
    const int kb = g * (CK_KC / 2);
    const int ke = min(CK_KC / 2, max(0, K - k0 - kb));
    const float* wp = w + (long long)(k0 + kb) * nout + j;
    int kk = 0;
    for (; kk + 4 <= ke; kk += 4) {          // 4 weights in registers, one 16-byte broadcast read per sample
      const float w0 = __ldg(wp + (long long)kk * nout), w1 = __ldg(wp + (long long)(kk + 1) * nout);
      const float w2 = __ldg(wp + (long long)(kk + 2) * nout), w3 = __ldg(wp + (long long)(kk + 3) * nout);
#pragma unroll
      for (int sI = 0; sI < CK_S; ++sI) {
        const float4 f = *reinterpret_cast<const float4*>(&sf[sI][kb + kk]);
        acc[sI] = fmaf(f.x, w0, acc[sI]); acc[sI] = fmaf(f.y, w1, acc[sI]);
        acc[sI] = fmaf(f.z, w2, acc[sI]); acc[sI] = fmaf(f.w, w3, acc[sI]);
      }
    }
    for (; kk < ke; ++kk) {
      const float wv = __ldg(wp + (long long)kk * nout);
#pragma unroll
      for (int sI = 0; sI < CK_S; ++sI) acc[sI] = fmaf(sf[sI][kb + kk], wv, acc[sI]);
    }
  }
  if (g == 1) {
#pragma unroll
    for (int sI = 0; sI < CK_S; ++sI) red[sI][j] = acc[sI];
  }
  __syncthreads();
  if (g == 0 && j < nout) {
#pragma unroll
    for (int sI = 0; sI < CK_S; ++sI)
      if (b0 + sI < B) part[((long long)ks * B + b0 + sI) * 128 + j] = acc[sI] + red[sI][j];
  }
}

// Block-wide (>= 128 threads): sk[t*nt + n] = normalised CDNA kernel tap t of transformed image n for sample b.
// Dense output j = (u*k + v)*nt + n.  tmp: 128 floats, sums: 8 floats of shared memory.
__device__ __forceinline__ void cdna_finalize(const float* __restrict__ part, int nks, int B, int b,
                                              const float* __restrict__ bias, int ksize, int nt, float* sk, float* tmp, float* sums) {
  const int kk = ksize * ksize, nout = kk * nt, j = threadIdx.x;
  if (j < nout) {
    float v = bias[j];
    for (int k0 = 0; k0 < nks; k0 += 8) {                    // 8 independent loads in flight, summed in a fixed order
      float pv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) pv[u] = k0 + u < nks ? __ldg(part + ((long long)(k0 + u) * B + b) * 128 + j) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) v += pv[u];
    }
    if (j / nt == (ksize / 2) * ksize + ksize / 2) v += 1.0f;
    tmp[j] = fmaxf(v - 1e-12f, 0.f) + 1e-12f;
  }
  __syncthreads();
  if (j < nt) {
    float sum = 0.f;
    for (int t = 0; t < kk; ++t) sum += tmp[t * nt + j];
    sums[j] = sum;
  }
  __syncthreads();
  if (j < nout) sk[j] = tmp[j] / sums[j % nt];
  __syncthreads();
}

// kern[b][n][tap]: one block per sample finalises the split-K partials once (the appliers read the 100 finished taps)
__global__ void __launch_bounds__(128) k_cdna_finalize(const float* __restrict__ part, int nks, int B,
                                                       const float* __restrict__ bias, int ksize, int nt, float* __restrict__ kern) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sk[128], tmp[128], sums[8];
  const int b = blockIdx.x, kk = ksize * ksize;
  cdna_finalize(part, nks, B, b, bias, ksize, nt, sk, tmp, sums);
  if (threadIdx.x < kk * nt) {
    const int t = threadIdx.x / nt, n = threadIdx.x % nt;
    kern[((long long)b * nt + n) * kk + t] = sk[threadIdx.x];
  }
}

// CDNA application (spec P6), nt = 4 transformed images, 5x5 kernels: block = (band of TR rows, sample).  The band
// (+2 halo rows, SYMMETRIC-mirrored) is staged in shared memory as float4 pixels, the four kernels of a tap are one
// float4: 2 LDS.128 + 12 FMA per tap.  Writes layers[.., 0..17] = T_0..T_3 (rgb each), prev image, first image.
__global__ void __launch_bounds__(256) k_cdna_apply4(View image, View first, const float* __restrict__ kern, int B, int H, int W,
                                                     int TR, View layers) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float4 sm4[];
  float4* sk4 = sm4;                       // [25] taps x 4 kernels
  float4* img = sm4 + 25 + 34;             // [(TR+4)][W]
  const int b = blockIdx.y, y0 = blockIdx.x * TR;
  if (threadIdx.x < 100) {                 // kern[b][n][t] -> sk[t][n]
    const int t = threadIdx.x >> 2, n = threadIdx.x & 3;
    reinterpret_cast<float*>(sk4)[threadIdx.x] = __ldg(kern + ((long long)b * 4 + n) * 25 + t);
  }
  for (int i = threadIdx.x; i < (TR + 4) * W; i += 256) {
    const int r = i / W, x = i - r * W;
    const int yy = mirror(y0 - 2 + r, H);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < H) {                // rows far below the image (last band) stay zero: never used by valid pixels
      const float* ip = vptr(image, b, (long long)yy * W + x);
      v = make_float4(__ldg(ip), __ldg(ip + 1), __ldg(ip + 2), 0.f);
    }
    img[i] = v;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < TR * W; p += 256) {
    const int r = p / W, x = p - r * W, y = y0 + r;
    if (y >= H) break;
    float a[4][3];
#pragma unroll
    for (int n = 0; n < 4; ++n) a[n][0] = a[n][1] = a[n][2] = 0.f;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        const int xx = mirror(x + v - 2, W);
        const float4 px = img[(r + u) * W + xx];
        const float4 kv = sk4[u * 5 + v];
        a[0][0] = fmaf(px.x, kv.x, a[0][0]); a[0][1] = fmaf(px.y, kv.x, a[0][1]); a[0][2] = fmaf(px.z, kv.x, a[0][2]);
        a[1][0] = fmaf(px.x, kv.y, a[1][0]); a[1][1] = fmaf(px.y, kv.y, a[1][1]); a[1][2] = fmaf(px.z, kv.y, a[1][2]);
        a[2][0] = fmaf(px.x, kv.z, a[2][0]); a[2][1] = fmaf(px.y, kv.z, a[2][1]); a[2][2] = fmaf(px.z, kv.z, a[2][2]);
        a[3][0] = fmaf(px.x, kv.w, a[3][0]); a[3][1] = fmaf(px.y, kv.w, a[3][1]); a[3][2] = fmaf(px.z, kv.w, a[3][2]);
      }
    }
    const float4 pv = img[(r + 2) * W + x];
    const float* fp = vptr(first, b, (long long)y * W + x);
    const float f0 = __ldg(fp), f1 = __ldg(fp + 1), f2 = __ldg(fp + 2);
    const long long o = voff(layers, b, (long long)y * W + x);
    vst4(layers, o, make_float4(a[0][0], a[0][1], a[0][2], a[1][0]));
    vst4(layers, o + 4, make_float4(a[1][1], a[1][2], a[2][0], a[2][1]));
    vst4(layers, o + 8, make_float4(a[2][2], a[3][0], a[3][1], a[3][2]));
    vst4(layers, o + 12, make_float4(pv.x, pv.y, pv.z, f0));
    vst1(layers, o + 16, f1);
    vst1(layers, o + 17, f2);
  }
}

// Same contract, W % 4 == 0: one thread = a strip of 4 horizontally adjacent pixels.  A filter row needs 8 staged pixels for
// the 4 outputs (instead of 4 x 5) and a tap's kernel vector is loaded once for the strip: 16 shared-memory loads per
// pixel instead of 50, same accumulation order per output (bit-identical to k_cdna_apply4).  Every pixel of `layers` is
// written as three 8-channel vectors (T_0..T_3, prev, first, zeros): channels 18..20 are rewritten by the scratch-image conv
// later in the step, 21..23 are padding.
__global__ void __launch_bounds__(256) k_cdna_apply4s(View image, View first, const float* __restrict__ kern, int B, int H, int W,
                                                      int TR, View layers) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float4 sm4[];
  float4* sk4 = sm4;                       // [25] taps x 4 kernels
  float4* img = sm4 + 25 + 34;             // [(TR+4)][W]
  const int b = blockIdx.y, y0 = blockIdx.x * TR;
  if (threadIdx.x < 100) {                 // kern[b][n][t] -> sk[t][n]
    const int t = threadIdx.x >> 2, n = threadIdx.x & 3;
    reinterpret_cast<float*>(sk4)[threadIdx.x] = __ldg(kern + ((long long)b * 4 + n) * 25 + t);
  }
  for (int i = threadIdx.x; i < (TR + 4) * W; i += 256) {
    const int r = i / W, x = i - r * W;
    const int yy = mirror(y0 - 2 + r, H);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < H) {
      const float* ip = vptr(image, b, (long long)yy * W + x);
      v = make_float4(__ldg(ip), __ldg(ip + 1), __ldg(ip + 2), 0.f);
    }
    img[i] = v;
  }
  __syncthreads();
  const int W4 = W >> 2;
  for (int sidx = threadIdx.x; sidx < TR * W4; sidx += 256) {
    const int r = sidx / W4, x0 = (sidx - r * W4) << 2, y = y0 + r;
    if (y >= H) break;
    int col[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) col[i] = mirror(x0 - 2 + i, W);
    float a[4][4][3];                      // [pixel][kernel][channel]
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int n = 0; n < 4; ++n) a[p][n][0] = a[p][n][1] = a[p][n][2] = 0.f;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      float4 px[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) px[i] = img[(r + u) * W + col[i]];
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        const float4 kv = sk4[u * 5 + v];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float4 q = px[p + v];
          a[p][0][0] = fmaf(q.x, kv.x, a[p][0][0]); a[p][0][1] = fmaf(q.y, kv.x, a[p][0][1]); a[p][0][2] = fmaf(q.z, kv.x, a[p][0][2]);
          a[p][1][0] = fmaf(q.x, kv.y, a[p][1][0]); a[p][1][1] = fmaf(q.y, kv.y, a[p][1][1]); a[p][1][2] = fmaf(q.z, kv.y, a[p][1][2]);
          a[p][2][0] = fmaf(q.x, kv.z, a[p][2][0]); a[p][2][1] = fmaf(q.y, kv.z, a[p][2][1]); a[p][2][2] = fmaf(q.z, kv.z, a[p][2][2]);
          a[p][3][0] = fmaf(q.x, kv.w, a[p][3][0]); a[p][3][1] = fmaf(q.y, kv.w, a[p][3][1]); a[p][3][2] = fmaf(q.z, kv.w, a[p][3][2]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float4 pv = img[(r + 2) * W + x0 + p];
      const float* fp = vptr(first, b, (long long)y * W + x0 + p);
      const float f0 = __ldg(fp), f1 = __ldg(fp + 1), f2 = __ldg(fp + 2);
      const long long o = voff(layers, b, (long long)y * W + x0 + p);
      float8 w0, w1, w2;
      w0.a = make_float4(a[p][0][0], a[p][0][1], a[p][0][2], a[p][1][0]);
      w0.b = make_float4(a[p][1][1], a[p][1][2], a[p][2][0], a[p][2][1]);
      w1.a = make_float4(a[p][2][2], a[p][3][0], a[p][3][1], a[p][3][2]);
      w1.b = make_float4(pv.x, pv.y, pv.z, f0);
      w2.a = make_float4(f1, f2, 0.f, 0.f);
      w2.b = make_float4(0.f, 0.f, 0.f, 0.f);
      vst8(layers, o, w0);
      vst8(layers, o + 8, w1);
      vst8(layers, o + 16, w2);
    }
  }
}

// generic (any nt <= 8, odd ksize): one thread per pixel, global loads
__global__ void __launch_bounds__(128) k_cdna_apply(View image, View first, const float* __restrict__ kern, int ksize, int nt,
                                                    int B, int H, int W, View layers) {
  __shared__ float sk[128];
  const int b = blockIdx.y;
  const int kk = ksize * ksize;
  if (threadIdx.x < kk * nt) {
    const int t = threadIdx.x / nt, n = threadIdx.x % nt;
    sk[threadIdx.x] = __ldg(kern + ((long long)b * nt + n) * kk + t);
  }
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
        const float kv = sk[(u * ksize + v) * nt + n];
        acc[n][0] = fmaf(r, kv, acc[n][0]);
        acc[n][1] = fmaf(g, kv, acc[n][1]);
        acc[n][2] = fmaf(bl, kv, acc[n][2]);
      }
    }
  }
  const long long o = voff(layers, b, pix);
  for (int n = 0; n < nt; ++n)
    for (int c = 0; c < 3; ++c) vst1(layers, o + 3 * n + c, acc[n][c]);
  const float* ip = vptr(image, b, pix);
  const float* fp = vptr(first, b, pix);
  for (int c = 0; c < 3; ++c) {
    vst1(layers, o + 3 * nt + c, __ldg(ip + c));
    vst1(layers, o + 3 * nt + 3 + c, __ldg(fp + c));
  }
}

constexpr int COMP_THREADS = 256;

// masks = softmax(logits); gen_image = sum_n m_n * layer_n;
// gen_distrib = sum_{n<nt} m_n * T_n(prev_d) + m_nt*prev_d + m_{nt+1}*first_d + m_{nt+2}*prev_d   (spec P8)
// Block = (band of TR rows, sample); the previous distribution's band (+halo, SYMMETRIC-mirrored) is staged in shared
// memory; kernels as [tap][nt].  Per-thread sums of the raw distribution are combined by a fixed shuffle tree + fixed
// warp order -> partial[b][p][band].
// NTT > 0: compile-time transformed-image count (the mask / kernel arrays live in registers); NTT == 0: generic
template <int NTT>
__global__ void __launch_bounds__(COMP_THREADS) k_composite(CompositeArgs a, int TR) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float smc[];
  __shared__ float red[COMP_THREADS / 32];
  const int b = blockIdx.y, y0 = blockIdx.x * TR;
  const int nt = NTT > 0 ? NTT : a.nt;
  const int kk = a.ksize * a.ksize, nm = nt + 3, pad = a.ksize / 2, W = a.W, H = a.H;
  const int BR = TR + 2 * pad;
  float* sk = smc;                         // [kk][nt]
  float* band = smc + ((kk * nt + 3) & ~3);   // [nd][BR][W]
  for (int i = threadIdx.x; i < nt * kk; i += COMP_THREADS) {
    const int t = i / nt, n = i - t * nt;
    sk[i] = a.kern[((long long)b * nt + n) * kk + t];
  }
  for (int i = threadIdx.x; i < a.nd * BR * W; i += COMP_THREADS) {
    const int p = i / (BR * W), rem = i - p * BR * W, r = rem / W, x = rem - r * W;
    const int yy = mirror(y0 - pad + r, H);
    band[i] = (yy >= 0 && yy < H) ? __ldg(vptr(a.prev_d, b, (long long)yy * W + x) + p) : 0.f;
  }
  __syncthreads();
  float dsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int q = threadIdx.x; q < TR * W; q += COMP_THREADS) {
    const int r = q / W, x = q - r * W, y = y0 + r;
    if (y >= H) break;
    const long long pix = (long long)y * W + x;
    float m[16];
    const float* lg = vptr(a.logits, b, pix);
    float mx = -3.4e38f;
    _Pragma("unroll") for (int n = 0; n < nm; ++n) { m[n] = __ldg(lg + n); mx = fmaxf(mx, m[n]); }
    float se = 0.f;
    _Pragma("unroll") for (int n = 0; n < nm; ++n) { m[n] = expf(m[n] - mx); se += m[n]; }
    _Pragma("unroll") for (int n = 0; n < nm; ++n) m[n] = m[n] / se;
    const long long lo = voff(a.layers, b, pix);
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (nm == 7) {                           // 21 channels = 6 vector loads (the buffer is padded to 24 channels)
      float l[24];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float4 v4 = vld4(a.layers, lo + 4 * j);
        l[4 * j] = v4.x; l[4 * j + 1] = v4.y; l[4 * j + 2] = v4.z; l[4 * j + 3] = v4.w;
      }
#pragma unroll
      for (int n = 0; n < 7; ++n) { g0 += m[n] * l[3 * n]; g1 += m[n] * l[3 * n + 1]; g2 += m[n] * l[3 * n + 2]; }
    } else {
      _Pragma("unroll") for (int n = 0; n < nm; ++n) {
        g0 += m[n] * vld1(a.layers, lo + 3 * n);
        g1 += m[n] * vld1(a.layers, lo + 3 * n + 1);
        g2 += m[n] * vld1(a.layers, lo + 3 * n + 2);
      }
    }
    float* gi = a.gen_image.p + voff(a.gen_image, b, pix);
    gi[0] = g0; gi[1] = g1; gi[2] = g2;
    float* gd = a.gen_distrib.p + voff(a.gen_distrib, b, pix);
    for (int p = 0; p < a.nd; ++p) {
      const float* bp = band + p * BR * W;
      float t[8];
      _Pragma("unroll") for (int n = 0; n < nt; ++n) t[n] = 0.f;
      const int ks = NTT > 0 ? 5 : a.ksize;        // the specialised instance is the 5x5 CDNA kernel
      _Pragma("unroll") for (int u = 0; u < ks; ++u)
        _Pragma("unroll") for (int v = 0; v < ks; ++v) {
          const float d = bp[(r + u) * W + mirror(x + v - pad, W)];
          const float* kp = sk + (u * a.ksize + v) * nt;
          _Pragma("unroll") for (int n = 0; n < nt; ++n) t[n] = fmaf(d, kp[n], t[n]);
        }
      const float pd = bp[(r + pad) * W + x], fd = __ldg(vptr(a.first_d, b, pix) + p);
      float v = 0.f;
      _Pragma("unroll") for (int n = 0; n < nt; ++n) v += m[n] * t[n];
      v += m[nt] * pd;
      v += m[nt + 1] * fd;
      v += m[nt + 2] * pd;
      gd[p] = v;
      dsum[p] += v;
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

// nt = 4, 5x5 kernels, 7 masks, W % 4 == 0, logits dense [.,7], gen_image dense [.,3]: one thread = a strip of 4 adjacent
// pixels.  28 logits and 12 output colours are 7 + 3 vector accesses, a filter row of the distribution stencil needs 8
// staged values for the 4 outputs and a tap's four kernel weights are one 16-byte load (the one-pixel kernel issues 125
// scalar shared-memory loads per pixel, this one ~17).  Accumulation order per output pixel is unchanged.
__global__ void __launch_bounds__(COMP_THREADS) k_composite4s(CompositeArgs a, int TR) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float smc[];
  __shared__ float red[COMP_THREADS / 32];
  const int b = blockIdx.y, y0 = blockIdx.x * TR;
  const int W = a.W, H = a.H, BR = TR + 4;
  float4* sk4 = reinterpret_cast<float4*>(smc);            // [25] taps x 4 kernels
  float* band = smc + 100;                                 // [nd][BR][W]
  if (threadIdx.x < 100) {
    const int t = threadIdx.x >> 2, n = threadIdx.x & 3;
    smc[threadIdx.x] = a.kern[((long long)b * 4 + n) * 25 + t];
  }
  for (int i = threadIdx.x; i < a.nd * BR * W; i += COMP_THREADS) {
    const int p = i / (BR * W), rem = i - p * BR * W, r = rem / W, x = rem - r * W;
    const int yy = mirror(y0 - 2 + r, H);
    band[i] = (yy >= 0 && yy < H) ? __ldg(vptr(a.prev_d, b, (long long)yy * W + x) + p) : 0.f;
  }
  __syncthreads();
  float dsum[4] = {0.f, 0.f, 0.f, 0.f};
  const int W4 = W >> 2;
  for (int sidx = threadIdx.x; sidx < TR * W4; sidx += COMP_THREADS) {
    const int r = sidx / W4, x0 = (sidx - r * W4) << 2, y = y0 + r;
    if (y >= H) break;
    const long long pix = (long long)y * W + x0;
    // masks of the 4 pixels
    float m[4][7];
    {
      float lg[28];
      const float4* lp = reinterpret_cast<const float4*>(vptr(a.logits, b, pix));
#pragma unroll
      for (int j = 0; j < 7; ++j) { const float4 v4 = __ldg(lp + j); lg[4 * j] = v4.x; lg[4 * j + 1] = v4.y; lg[4 * j + 2] = v4.z; lg[4 * j + 3] = v4.w; }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float mx = -3.4e38f;
#pragma unroll
        for (int n = 0; n < 7; ++n) { m[p][n] = lg[7 * p + n]; mx = fmaxf(mx, m[p][n]); }
        float se = 0.f;
#pragma unroll
        for (int n = 0; n < 7; ++n) { m[p][n] = expf(m[p][n] - mx); se += m[p][n]; }
#pragma unroll
        for (int n = 0; n < 7; ++n) m[p][n] = m[p][n] / se;
      }
    }
    // colour composite
    float g[12];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const long long lo = voff(a.layers, b, pix + p);
      float l[24];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float8 v8 = vld8(a.layers, lo + 8 * j);
        l[8 * j] = v8.a.x; l[8 * j + 1] = v8.a.y; l[8 * j + 2] = v8.a.z; l[8 * j + 3] = v8.a.w;
        l[8 * j + 4] = v8.b.x; l[8 * j + 5] = v8.b.y; l[8 * j + 6] = v8.b.z; l[8 * j + 7] = v8.b.w;
      }
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
      for (int n = 0; n < 7; ++n) { g0 += m[p][n] * l[3 * n]; g1 += m[p][n] * l[3 * n + 1]; g2 += m[p][n] * l[3 * n + 2]; }
      g[3 * p] = g0; g[3 * p + 1] = g1; g[3 * p + 2] = g2;
    }
    float4* gi = reinterpret_cast<float4*>(a.gen_image.p + voff(a.gen_image, b, pix));
    gi[0] = make_float4(g[0], g[1], g[2], g[3]);
    gi[1] = make_float4(g[4], g[5], g[6], g[7]);
    gi[2] = make_float4(g[8], g[9], g[10], g[11]);
    // distribution: 5x5 stencil of the previous distribution with the sample's 4 kernels, then the mask mix
    int col[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) col[i] = mirror(x0 - 2 + i, W);
    float* gd = a.gen_distrib.p + voff(a.gen_distrib, b, pix);
    for (int p = 0; p < a.nd; ++p) {
      const float* bp = band + p * BR * W;
      float t[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) t[q][0] = t[q][1] = t[q][2] = t[q][3] = 0.f;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        float d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = bp[(r + u) * W + col[i]];
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          const float4 kv = sk4[u * 5 + v];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            t[q][0] = fmaf(d[q + v], kv.x, t[q][0]); t[q][1] = fmaf(d[q + v], kv.y, t[q][1]);
            t[q][2] = fmaf(d[q + v], kv.z, t[q][2]); t[q][3] = fmaf(d[q + v], kv.w, t[q][3]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float pd = bp[(r + 2) * W + x0 + q], fd = __ldg(vptr(a.first_d, b, pix + q) + p);
        float v = 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n) v += m[q][n] * t[q][n];
        v += m[q][4] * pd;
        v += m[q][5] * fd;
        v += m[q][6] * pd;
        gd[(long long)q * a.gen_distrib.pix_stride + p] = v;
        dsum[p] += v;
      }
    }
  }
  for (int p = 0; p < a.nd; ++p) {
    float v = dsum[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float sacc = 0.f;
      for (int i = 0; i < COMP_THREADS / 32; ++i) sacc += red[i];
      a.partial[((long long)b * a.nd + p) * gridDim.x + blockIdx.x] = sacc;
    }
    __syncthreads();
  }
}

__global__ void k_distrib_normalize(View d, const float* __restrict__ partial, int nblk, int H, int W, int nd) {
  pdl_wait();
  pdl_trigger();
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

inline int band_rows(int H, int W) {                 // rows per block of the banded CDNA-apply kernel (1-2 pixels per thread)
  int tr = 8;
  while (tr > 2 && tr * W > 512) tr /= 2;
  return tr;
}
inline int comp_rows(int H, int W) {                 // rows per block of the composite kernels (a 4-pixel strip per thread and pass)
  int tr = 16;
  while (tr > 1 && (tr * W / 4 > COMP_THREADS || tr > H)) tr /= 2;
  return tr;
}
inline int strip_rows(int H, int W) {                // rows per block of the strip-mined CDNA apply
  int tr = 16;
  while (tr > 2 && (tr * W / 4 > 256 || tr > H)) tr /= 2;
  return tr;
}
bool strips_enabled() {
  static const bool on = !(getenv("VF_STRIPS") && atoi(getenv("VF_STRIPS")) == 0);     // A/B switch: 0 = one pixel per thread
  return on;
}
size_t cdna_partial_floats(int K, int B) { return (size_t)((K + CK_KC - 1) / CK_KC) * B * 128; }
void launch_cdna_kernels(View feat, int npix, const float* w, int ksize, int nt, int B, float* part, cudaStream_t s) {
  ++g_launch_counter;
  const int K = npix * feat.C;
  dim3 grid((K + CK_KC - 1) / CK_KC, (B + CK_S - 1) / CK_S);
  launch_k(k_cdna_partial, dim3(grid), dim3(256), 0, s, feat, npix, w, ksize * ksize * nt, B, part);
}
void launch_cdna_apply(View image, View first, const float* part, int K, const float* bias, float* kern, int ksize, int nt,
                       int B, int H, int W, View layers, cudaStream_t s) {
  g_launch_counter += 2;
  const int nks = (K + CK_KC - 1) / CK_KC;
  launch_k(k_cdna_finalize, dim3(B), dim3(128), 0, s, part, nks, B, bias, ksize, nt, kern);
  const bool al8 = layers.lo_off ? ((layers.pix_stride | layers.ch_off) % 8 == 0 && layers.sample_stride % 8 == 0 && layers.lo_off % 8 == 0)
                                 : ((layers.pix_stride | layers.ch_off) % 4 == 0 && layers.sample_stride % 4 == 0);
  if (nt == 4 && ksize == 5 && W >= 8 && W % 4 == 0 && layers.C >= 24 && al8 && strips_enabled()) {
    const int TR = strip_rows(H, W);
    dim3 grid((H + TR - 1) / TR, B);
    const size_t smem = (size_t)(25 + 34 + (TR + 4) * W) * sizeof(float4);
    launch_k(k_cdna_apply4s, dim3(grid), dim3(256), smem, s, image, first, (const float*)kern, B, H, W, TR, layers);
  } else if (nt == 4 && ksize == 5 && W >= 8) {
    const int TR = band_rows(H, W);
    dim3 grid((H + TR - 1) / TR, B);
    const size_t smem = (size_t)(25 + 34 + (TR + 4) * W) * sizeof(float4);
    launch_k(k_cdna_apply4, dim3(grid), dim3(256), smem, s, image, first, (const float*)kern, B, H, W, TR, layers);
  } else {
    dim3 grid((H * W + 127) / 128, B);
    k_cdna_apply<<<grid, 128, 0, s>>>(image, first, (const float*)kern, ksize, nt, B, H, W, layers);
  }
}
int composite_blocks(int H, int W) { const int TR = comp_rows(H, W); return (H + TR - 1) / TR; }
void launch_composite(const CompositeArgs& a, int B, cudaStream_t s) {
  ++g_launch_counter;
  const int TR = comp_rows(a.H, a.W);
  dim3 grid((a.H + TR - 1) / TR, B);
  const size_t smem = (size_t)(((a.nt * a.ksize * a.ksize + 3) & ~3) + a.nd * (TR + a.ksize - 1) * a.W) * sizeof(float);
  const bool dense = a.logits.pix_stride == 7 && a.logits.ch_off == 0 && a.logits.sample_stride % 4 == 0 && !a.logits.lo_off &&
                     a.gen_image.pix_stride == 3 && a.gen_image.ch_off == 0 && a.gen_image.sample_stride % 4 == 0 &&
                     ((uintptr_t)a.logits.p % 16 == 0) && ((uintptr_t)a.gen_image.p % 16 == 0) &&
                     (a.layers.lo_off ? ((a.layers.pix_stride | a.layers.ch_off) % 8 == 0 && a.layers.sample_stride % 8 == 0 && a.layers.lo_off % 8 == 0)
                                      : ((a.layers.pix_stride | a.layers.ch_off) % 4 == 0 && a.layers.sample_stride % 4 == 0)) && a.layers.C >= 24;
  if (a.nt == 4 && a.ksize == 5 && a.W % 4 == 0 && a.nd <= 4 && dense && strips_enabled()) {
    const size_t smem4 = (size_t)(100 + a.nd * (TR + 4) * a.W) * sizeof(float);
    launch_k(k_composite4s, dim3(grid), dim3(COMP_THREADS), smem4, s, a, TR);
  } else if (a.nt == 4 && a.ksize == 5) launch_k(k_composite<4>, dim3(grid), dim3(COMP_THREADS), smem, s, a, TR);
  else launch_k(k_composite<0>, dim3(grid), dim3(COMP_THREADS), smem, s, a, TR);
}
void launch_distrib_normalize(View d, const float* partial, int nblk, int B, int H, int W, int nd, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid((H * W * nd + 255) / 256, B);
  launch_k(k_distrib_normalize, dim3(grid), dim3(256), 0, s, d, partial, nblk, H, W, nd);
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

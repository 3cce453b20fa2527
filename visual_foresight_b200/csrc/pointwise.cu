// pointwise.cu — instance-norm statistics, normalise+activation, conv-LSTM pointwise, bilinear x2
// upsample, and the action/state vector with its per-layer border-class bias.  All HBM/L2-bound:
// threads run along the channel dimension (NHWC innermost) so every warp access is contiguous.
#include <stdlib.h>

#include <algorithm>

#include <cooperative_groups.h>

#include "vf_common.cuh"

namespace vf {
namespace {

__device__ __forceinline__ const float* vptr(const View& v, int b, long long pix) {
  return v.p + (long long)b * v.sample_stride + pix * v.pix_stride + v.ch_off;
}

// Instance-norm statistics, one pass: grid (B, ceil(C/32), S pixel splits), block 256 = 8 channel quads x 32 pixel lanes.
// Every thread accumulates sum and sum-of-squares of 4 channels in float64 (no cancellation in E[x^2]-mean^2 at fp32 data
// precision) over its pixels, 4 pixels (16-byte loads) in flight per thread; the 32 pixel lanes are combined in a fixed
// order, and the S partials of a (sample, channel) are combined in a fixed order by the consumer (stat_of):
// bit-reproducible, independent of sample count and GPU count.
constexpr int STATS_MAX_SPLIT = 8;
__device__ __forceinline__ float4 pooled4(const View& x, int b, int W, int pool, int p, int c) {
  if (!pool) return vld4(x, voff(x, b, p) + c);
  const int y = p / W, xx = p - y * W, Wi = W * 2;
  const long long o = voff(x, b, (long long)(2 * y) * Wi + 2 * xx) + c;
  const float4 a0 = vld4(x, o), a1 = vld4(x, o + x.pix_stride);
  const float4 a2 = vld4(x, o + (long long)Wi * x.pix_stride), a3 = vld4(x, o + (long long)(Wi + 1) * x.pix_stride);
  return make_float4(((a0.x + a1.x) + (a2.x + a3.x)) * 0.25f, ((a0.y + a1.y) + (a2.y + a3.y)) * 0.25f,
                     ((a0.z + a1.z) + (a2.z + a3.z)) * 0.25f, ((a0.w + a1.w) + (a2.w + a3.w)) * 0.25f);
}
// "last arriver finalises" (see conv_mma.cu): partial slots written by other blocks are read from L2
struct FinArgs { float* fin; int* cnt; float eps; };
__device__ __forceinline__ void finalize_plane(const double* part, int S, int npix, float eps, float* fin) {
  double ts = 0.0, tq = 0.0;
  for (int k = 0; k < S; ++k) { ts += __ldcg(part + 2 * k); tq += __ldcg(part + 2 * k + 1); }
  const double mean = ts / npix;
  double var = tq / npix - mean * mean;
  if (var < 0.0) var = 0.0;
  fin[0] = (float)mean;
  fin[1] = (float)(1.0 / sqrt(var + (double)eps));
}
__global__ void __launch_bounds__(256) k_plane_stats(View x, int H, int W, int pool, double* partial, FinArgs fa) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  const int cq = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.y * 32 + cq * 4;
  const int S = gridDim.z, sp = blockIdx.z;
  const int npix = H * W;
  const int p0 = (int)((long long)npix * sp / S), p1 = (int)((long long)npix * (sp + 1) / S);
  __shared__ double red[32][8][9];
  double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < x.C) {
    int p = p0 + pl;
    for (; p + 96 < p1; p += 128) {                 // 4 independent loads in flight
      const float4 v0 = pooled4(x, b, W, pool, p, c), v1 = pooled4(x, b, W, pool, p + 32, c);
      const float4 v2 = pooled4(x, b, W, pool, p + 64, c), v3 = pooled4(x, b, W, pool, p + 96, c);
      const float4 vv[4] = {v0, v1, v2, v3};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double d0 = vv[u].x, d1 = vv[u].y, d2 = vv[u].z, d3 = vv[u].w;
        s[0] += d0; q[0] = fma(d0, d0, q[0]); s[1] += d1; q[1] = fma(d1, d1, q[1]);
        s[2] += d2; q[2] = fma(d2, d2, q[2]); s[3] += d3; q[3] = fma(d3, d3, q[3]);
      }
    }
    for (; p < p1; p += 32) {
      const float4 v = pooled4(x, b, W, pool, p, c);
      const double d0 = v.x, d1 = v.y, d2 = v.z, d3 = v.w;
      s[0] += d0; q[0] = fma(d0, d0, q[0]); s[1] += d1; q[1] = fma(d1, d1, q[1]);
      s[2] += d2; q[2] = fma(d2, d2, q[2]); s[3] += d3; q[3] = fma(d3, d3, q[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[pl][cq][j] = s[j]; red[pl][cq][4 + j] = q[j]; }
  __syncthreads();
  if (threadIdx.x < 32 && blockIdx.y * 32 + threadIdx.x < x.C) {
    const int cc = threadIdx.x;
    double ts = 0.0, tq = 0.0;
    for (int i = 0; i < 32; ++i) { ts += red[i][cc >> 2][cc & 3]; tq += red[i][cc >> 2][4 + (cc & 3)]; }
    double* o = partial + (((long long)b * x.C + blockIdx.y * 32 + cc) * S + sp) * 2;
    o[0] = ts;
    o[1] = tq;
  }
  if (fa.fin && threadIdx.x < 32) {                  // warp 0 wrote this block's slots: the last of the S blocks finalises 32 channels
    __threadfence();
    __syncwarp();
    int last = 0;
    if (threadIdx.x == 0) {
      int* cnt = fa.cnt + (long long)b * VF_STAT_CNT_STRIDE + blockIdx.y;
      last = atomicAdd(cnt, 1) == S - 1;
      if (last) *cnt = 0;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    const int c = blockIdx.y * 32 + threadIdx.x;
    if (last && c < x.C) {
      __threadfence();
      finalize_plane(partial + ((long long)b * x.C + c) * S * 2, S, npix, fa.eps, fa.fin + ((long long)b * x.C + c) * 2);
    }
  }
}

__global__ void k_stats_finalize(const double* __restrict__ partial, int n, int S, int npix, float eps, float* stats) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double ts = 0.0, tq = 0.0;
  for (int k = 0; k < S; ++k) { ts += partial[((long long)i * S + k) * 2]; tq += partial[((long long)i * S + k) * 2 + 1]; }
  const double mean = ts / npix;
  double var = tq / npix - mean * mean;            // biased variance
  if (var < 0.0) var = 0.0;
  stats[(long long)i * 2] = (float)mean;
  stats[(long long)i * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// (mean, rstd) of plane `idx` from its S float64 partial sums: the same arithmetic as k_stats_finalize, run by the
// consumer in its prologue so that no separate finalize launch is needed
__device__ __forceinline__ float2 stat_of(const StatsRef& r, long long idx) {
  if (r.fin) return __ldg(reinterpret_cast<const float2*>(r.fin) + idx);
  const double2* p = reinterpret_cast<const double2*>(r.partial) + idx * r.S;
  double ts = 0.0, tq = 0.0;
  for (int k0 = 0; k0 < r.S; k0 += 8) {           // 8 independent 16-byte loads in flight, summed in slot order (== k_stats_finalize)
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = k0 + u < r.S ? __ldg(p + k0 + u) : make_double2(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 8; ++u) { ts += v[u].x; tq += v[u].y; }
  }
  const double mean = ts / r.npix;
  double var = tq / r.npix - mean * mean;
  if (var < 0.0) var = 0.0;
  return make_float2((float)mean, (float)(1.0 / sqrt(var + (double)r.eps)));
}

// one thread = 4 consecutive channels of one output pixel; grid.y = sample
// y2 (optional, y2.C > 0): channels >= y.C go to the second view (the merged scratch/mask head conv feeds two convolutions,
// each of which should read a dense 32-channel buffer rather than every other 64 bytes of a shared one)
__global__ void __launch_bounds__(256) k_norm_act(View x, int H, int W, int pool, StatsRef sr,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta, int act, View y, View y2) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int C4 = x.C >> 2;
  const int total = H * W * C4;
  __shared__ __align__(16) float st[2 * 512];
  for (int c = threadIdx.x; c < x.C; c += blockDim.x) {
    const float2 v = stat_of(sr, (long long)b * x.C + c);
    st[2 * c] = v.x;
    st[2 * c + 1] = v.y;
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (i % C4) << 2;
    const int pix = i / C4;
    float4 v;
    if (!pool) {
      v = vld4(x, voff(x, b, pix) + c);
    } else {
      const int py = pix / W, px = pix - py * W, Wi = 2 * W;
      const float* p00 = vptr(x, b, (long long)(2 * py) * Wi + 2 * px) + c;
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(p00)), a1 = __ldg(reinterpret_cast<const float4*>(p00 + x.pix_stride));
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(p00 + (long long)Wi * x.pix_stride));
      const float4 a3 = __ldg(reinterpret_cast<const float4*>(p00 + (long long)(Wi + 1) * x.pix_stride));
      v.x = ((a0.x + a1.x) + (a2.x + a3.x)) * 0.25f; v.y = ((a0.y + a1.y) + (a2.y + a3.y)) * 0.25f;
      v.z = ((a0.z + a1.z) + (a2.z + a3.z)) * 0.25f; v.w = ((a0.w + a1.w) + (a2.w + a3.w)) * 0.25f;
    }
    const float4 s0 = *reinterpret_cast<const float4*>(st + 2 * c), s1 = *reinterpret_cast<const float4*>(st + 2 * c + 4);
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c)), b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 o;                                            // stats layout: (mean, rstd) pairs
    o.x = (v.x - s0.x) * s0.y * g4.x + b4.x; o.y = (v.y - s0.z) * s0.w * g4.y + b4.y;
    o.z = (v.z - s1.x) * s1.y * g4.z + b4.z; o.w = (v.w - s1.z) * s1.w * g4.w + b4.w;
    if (act == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (c < y.C) vst4(y, voff(y, b, pix) + c, o);
    else vst4(y2, voff(y2, b, pix) + (c - y.C), o);
  }
}

// 8 channels per thread: two 16-byte loads of the float32 conv output (x4 when pooling), one 16-byte store per fp16 plane
template <int POOL>
__global__ void __launch_bounds__(256) k_norm_act8(View x, int H, int W, StatsRef sr, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, int act, View y, View y2) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int C8 = x.C >> 3;
  const int total = H * W * C8;
  __shared__ __align__(16) float sm_[512], sr_[512], sg_[512], sb_[512];     // mean, rstd, gamma, beta per channel
  for (int c = threadIdx.x; c < x.C; c += blockDim.x) {
    const float2 v = stat_of(sr, (long long)b * x.C + c);
    sm_[c] = v.x; sr_[c] = v.y; sg_[c] = __ldg(gamma + c); sb_[c] = __ldg(beta + c);
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (i % C8) << 3;
    const int pix = i / C8;
    float4 v0, v1;
    if (!POOL) {
      const float* p = vptr(x, b, pix) + c;
      v0 = __ldg(reinterpret_cast<const float4*>(p));
      v1 = __ldg(reinterpret_cast<const float4*>(p + 4));
    } else {
      const int py = pix / W, px = pix - py * W, Wi = 2 * W;
      const float* p00 = vptr(x, b, (long long)(2 * py) * Wi + 2 * px) + c;
      const float* p10 = p00 + (long long)Wi * x.pix_stride;
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(p00)), a1 = __ldg(reinterpret_cast<const float4*>(p00 + x.pix_stride));
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(p10)), a3 = __ldg(reinterpret_cast<const float4*>(p10 + x.pix_stride));
      const float4 e0 = __ldg(reinterpret_cast<const float4*>(p00 + 4)), e1 = __ldg(reinterpret_cast<const float4*>(p00 + x.pix_stride + 4));
      const float4 e2 = __ldg(reinterpret_cast<const float4*>(p10 + 4)), e3 = __ldg(reinterpret_cast<const float4*>(p10 + x.pix_stride + 4));
      v0 = make_float4(((a0.x + a1.x) + (a2.x + a3.x)) * 0.25f, ((a0.y + a1.y) + (a2.y + a3.y)) * 0.25f,
                       ((a0.z + a1.z) + (a2.z + a3.z)) * 0.25f, ((a0.w + a1.w) + (a2.w + a3.w)) * 0.25f);
      v1 = make_float4(((e0.x + e1.x) + (e2.x + e3.x)) * 0.25f, ((e0.y + e1.y) + (e2.y + e3.y)) * 0.25f,
                       ((e0.z + e1.z) + (e2.z + e3.z)) * 0.25f, ((e0.w + e1.w) + (e2.w + e3.w)) * 0.25f);
    }
    const float4 m0 = *reinterpret_cast<const float4*>(sm_ + c), m1 = *reinterpret_cast<const float4*>(sm_ + c + 4);
    const float4 r0 = *reinterpret_cast<const float4*>(sr_ + c), r1 = *reinterpret_cast<const float4*>(sr_ + c + 4);
    const float4 g0 = *reinterpret_cast<const float4*>(sg_ + c), g1 = *reinterpret_cast<const float4*>(sg_ + c + 4);
    const float4 q0 = *reinterpret_cast<const float4*>(sb_ + c), q1 = *reinterpret_cast<const float4*>(sb_ + c + 4);
    float8 o;                                                    // the arithmetic of k_norm_act, term for term
    o.a = make_float4((v0.x - m0.x) * r0.x * g0.x + q0.x, (v0.y - m0.y) * r0.y * g0.y + q0.y,
                      (v0.z - m0.z) * r0.z * g0.z + q0.z, (v0.w - m0.w) * r0.w * g0.w + q0.w);
    o.b = make_float4((v1.x - m1.x) * r1.x * g1.x + q1.x, (v1.y - m1.y) * r1.y * g1.y + q1.y,
                      (v1.z - m1.z) * r1.z * g1.z + q1.z, (v1.w - m1.w) * r1.w * g1.w + q1.w);
    if (act == ACT_RELU) {
      o.a = make_float4(fmaxf(o.a.x, 0.f), fmaxf(o.a.y, 0.f), fmaxf(o.a.z, 0.f), fmaxf(o.a.w, 0.f));
      o.b = make_float4(fmaxf(o.b.x, 0.f), fmaxf(o.b.y, 0.f), fmaxf(o.b.z, 0.f), fmaxf(o.b.w, 0.f));
    }
    if (c < y.C) vst8(y, voff(y, b, pix) + c, o);
    else vst8(y2, voff(y2, b, pix) + (c - y.C), o);
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float gate_norm(const View& g, int b, long long pix, int ch, const float* gstats,
                                           const float* gg, const float* gb) {
  const float v = __ldg(vptr(g, b, pix) + ch);
  const float* st = gstats + ((long long)b * g.C + ch) * 2;
  return (v - st[0]) * st[1] * gg[ch] + gb[ch];
}

// gate order along channels: i, j, f, o   (spec P3).  grid (pixel blocks, sample); 256 threads = (256/F) pixel lanes x F
// channels (F in {32, 64, 128}: each thread keeps ONE channel).  The instance-norm statistics of the new cell state
// are accumulated on the fly (float64 per thread, lanes combined in a fixed order) -> partial[(b*F+f)*S + block].
__global__ void __launch_bounds__(256) k_lstm_gates(View gates, int HW, int F, StatsRef gsr,
                                                    const float* __restrict__ gg, const float* __restrict__ gb, float fb, float* c,
                                                    double* partial, FinArgs fa) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int lanes = 256 / F;
  const int f = threadIdx.x % F, pl = threadIdx.x / F;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  float* cb_ = c + (long long)b * HW * F;
  const float2 si = stat_of(gsr, (long long)b * gates.C + f), sj = stat_of(gsr, (long long)b * gates.C + F + f);
  const float2 sf = stat_of(gsr, (long long)b * gates.C + 2 * F + f);
  const float mi = si.x, ri = si.y, mj = sj.x, rj = sj.y, mf = sf.x, rf = sf.y;
  const float gi_g = gg[f], gi_b = gb[f], gj_g = gg[F + f], gj_b = gb[F + f], gf_g = gg[2 * F + f], gf_b = gb[2 * F + f];
  double s = 0.0, q = 0.0;
  int pix = p0 + pl;
  for (; pix + lanes < p1; pix += 2 * lanes) {              // two pixels in flight per thread
    const float* gp = vptr(gates, b, pix);
    const float* gq = vptr(gates, b, pix + lanes);
    const float a0 = __ldg(gp + f), a1 = __ldg(gp + F + f), a2 = __ldg(gp + 2 * F + f);
    const float e0 = __ldg(gq + f), e1 = __ldg(gq + F + f), e2 = __ldg(gq + 2 * F + f);
    const int i0 = pix * F + f, i1 = (pix + lanes) * F + f;
    const float c0 = cb_[i0], c1 = cb_[i1];
    const float cn0 = c0 * sigmoidf_((a2 - mf) * rf * gf_g + gf_b + fb) + sigmoidf_((a0 - mi) * ri * gi_g + gi_b) * tanhf((a1 - mj) * rj * gj_g + gj_b);
    const float cn1 = c1 * sigmoidf_((e2 - mf) * rf * gf_g + gf_b + fb) + sigmoidf_((e0 - mi) * ri * gi_g + gi_b) * tanhf((e1 - mj) * rj * gj_g + gj_b);
    cb_[i0] = cn0;
    cb_[i1] = cn1;
    s += (double)cn0; q = fma((double)cn0, (double)cn0, q);
    s += (double)cn1; q = fma((double)cn1, (double)cn1, q);
  }
  for (; pix < p1; pix += lanes) {
    const float* gp = vptr(gates, b, pix);
    const float gi = (__ldg(gp + f) - mi) * ri * gi_g + gi_b;
    const float gj = (__ldg(gp + F + f) - mj) * rj * gj_g + gj_b;
    const float gf = (__ldg(gp + 2 * F + f) - mf) * rf * gf_g + gf_b;
    const int i = pix * F + f;
    const float cn = cb_[i] * sigmoidf_(gf + fb) + sigmoidf_(gi) * tanhf(gj);
    cb_[i] = cn;
    s += (double)cn;
    q = fma((double)cn, (double)cn, q);
  }
  __shared__ double red[2][256];
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  if (pl == 0) {
    double ts = 0.0, tq = 0.0;
    for (int l = 0; l < lanes; ++l) { ts += red[0][l * F + f]; tq += red[1][l * F + f]; }
    double* o = partial + (((long long)b * F + f) * gridDim.x + blockIdx.x) * 2;
    o[0] = ts;
    o[1] = tq;
  }
  if (fa.fin) {                                      // the last of the sample's gridDim.x blocks finalises the cell-state statistics
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      int* cnt = fa.cnt + (long long)b * VF_STAT_CNT_STRIDE;
      const int last = atomicAdd(cnt, 1) == (int)gridDim.x - 1;
      if (last) *cnt = 0;
      s_last = last;
    }
    __syncthreads();
    if (s_last && threadIdx.x < F) {
      __threadfence();
      finalize_plane(partial + ((long long)b * F + threadIdx.x) * gridDim.x * 2, gridDim.x, HW, fa.eps, fa.fin + ((long long)b * F + threadIdx.x) * 2);
    }
  }
}

__global__ void __launch_bounds__(256) k_lstm_gates_generic(View gates, int HW, int F, const float* __restrict__ gstats,
                                                            const float* __restrict__ gg, const float* __restrict__ gb, float fb, float* c) {
  const int b = blockIdx.y;
  const int total = HW * F;
  float* cb_ = c + (long long)b * total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int f = i % F, pix = i / F;
    const float gi = gate_norm(gates, b, pix, f, gstats, gg, gb);
    const float gj = gate_norm(gates, b, pix, F + f, gstats, gg, gb);
    const float gf = gate_norm(gates, b, pix, 2 * F + f, gstats, gg, gb);
    cb_[i] = cb_[i] * sigmoidf_(gf + fb) + sigmoidf_(gi) * tanhf(gj);
  }
}

__global__ void __launch_bounds__(256) k_lstm_out(View gates, int HW, int F, StatsRef gsr,
                                                  const float* __restrict__ gg, const float* __restrict__ gb,
                                                  StatsRef csr, const float* __restrict__ cg,
                                                  const float* __restrict__ cb, float* c, View h, View h2, int W) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int total = HW * F;
  float* cb_ = c + (long long)b * total;
  // per-channel affine forms of both instance norms: y = x * a + d
  __shared__ float ca[512], cd[512], oa[512], od[512];
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float2 cs = stat_of(csr, (long long)b * F + f);
    const float2 os = stat_of(gsr, (long long)b * gates.C + 3 * F + f);
    ca[f] = cs.x; cd[f] = cs.y; oa[f] = os.x; od[f] = os.y;
  }
  __syncthreads();
  // one thread = 4 consecutive channels of one pixel (16-byte accesses)
  const int F4 = F >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (total >> 2); i += gridDim.x * blockDim.x) {
    const int f = (i % F4) << 2, pix = i / F4;
    float4* cp = reinterpret_cast<float4*>(cb_ + (long long)pix * F + f);
    const float4 cv = *cp;
    const float4 ov = __ldg(reinterpret_cast<const float4*>(vptr(gates, b, pix) + 3 * F + f));
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(cg + f)), b4 = __ldg(reinterpret_cast<const float4*>(cb + f));
    const float4 og = __ldg(reinterpret_cast<const float4*>(gg + 3 * F + f)), ob = __ldg(reinterpret_cast<const float4*>(gb + 3 * F + f));
    float4 cn, hv;
    cn.x = (cv.x - ca[f]) * cd[f] * g4.x + b4.x;             cn.y = (cv.y - ca[f + 1]) * cd[f + 1] * g4.y + b4.y;
    cn.z = (cv.z - ca[f + 2]) * cd[f + 2] * g4.z + b4.z;     cn.w = (cv.w - ca[f + 3]) * cd[f + 3] * g4.w + b4.w;
    *cp = cn;
    hv.x = tanhf(cn.x) * sigmoidf_((ov.x - oa[f]) * od[f] * og.x + ob.x);
    hv.y = tanhf(cn.y) * sigmoidf_((ov.y - oa[f + 1]) * od[f + 1] * og.y + ob.y);
    hv.z = tanhf(cn.z) * sigmoidf_((ov.z - oa[f + 2]) * od[f + 2] * og.z + ob.z);
    hv.w = tanhf(cn.w) * sigmoidf_((ov.w - oa[f + 3]) * od[f + 3] * og.w + ob.w);
    vst4(h, voff(h, b, pix) + f, hv);
    if (h2.p) {          // space-to-depth copy for the next encoder conv (2x2 pixel blocks -> channels [sub-position][F])
      const int y = pix / W, x = pix - y * W;
      vst4(h2, voff(h2, b, (long long)(y >> 1) * (W >> 1) + (x >> 1)) + ((y & 1) * 2 + (x & 1)) * F + f, hv);
    }
  }
}

// Both halves of the conv-LSTM pointwise (k_lstm_gates + k_lstm_out) in ONE kernel: a thread-block cluster of CL CTAs owns
// one sample, every thread keeps its NIT float4 slices of the new cell state and of the normalised output gate in registers,
// the instance-norm statistics of the new cell state are reduced inside the block (fixed order), exchanged between the CTAs
// of the cluster through distributed shared memory (fixed rank order) and applied without the cell state ever leaving the SM:
// one read of the gates, one read and one write of c, one write of h; no partial sums in HBM, no finalize launch.
// Requires finalised gate statistics (gsr.fin), 256 % (F/4) == 0 and HW * F / 4 == CL * NIT * 256.
template <int NIT>
__global__ void __launch_bounds__(256) k_lstm_fused(View gates, int HW, int F, const float* __restrict__ gfin,
                                                    const float* __restrict__ gg, const float* __restrict__ gb, float fb,
                                                    const float* __restrict__ cg, const float* __restrict__ cb, float eps, float* c,
                                                    View h, View h2, int W, int CL) {
  namespace cgp = cooperative_groups;
  pdl_wait();
  pdl_trigger();
  cgp::cluster_group cluster = cgp::this_cluster();
  const int rank = blockIdx.x;                       // grid.x == cluster size: the CTA's rank in its cluster
  const int b = blockIdx.y;
  const int F4 = F >> 2;
  const int f = (threadIdx.x % F4) << 2;             // this thread's 4 channels (the same for all its items: 256 % F4 == 0)
  const int per = NIT * 256;
  float* cb_ = c + (long long)b * HW * F;
  __shared__ double red[256][8];
  __shared__ double part[256][2];                    // this CTA's per-channel (sum, sum of squares), F <= 256 channels... F4 <= 64 threads x 4
  __shared__ float cst[256][2];
  // gate normalisation of this thread's channels: y = (x - mean) * rstd * gamma + beta
  float gm[4][4], gr[4][4], ga[4][4], gbt[4][4];    // [gate i,j,f,o][channel]
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(gfin) + (long long)b * gates.C + g * F + f + j);
      gm[g][j] = st.x; gr[g][j] = st.y;
      ga[g][j] = __ldg(gg + g * F + f + j); gbt[g][j] = __ldg(gb + g * F + f + j);
    }
  float4 cn[NIT], on[NIT];
  double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    const int idx = rank * per + k * 256 + threadIdx.x;
    const int pix = idx / F4;
    const float* gp = vptr(gates, b, pix) + f;
    const float4 vi = __ldg(reinterpret_cast<const float4*>(gp)), vj = __ldg(reinterpret_cast<const float4*>(gp + F));
    const float4 vf = __ldg(reinterpret_cast<const float4*>(gp + 2 * F)), vo = __ldg(reinterpret_cast<const float4*>(gp + 3 * F));
    const float4 c0 = *reinterpret_cast<const float4*>(cb_ + (long long)pix * F + f);
    const float xi[4] = {vi.x, vi.y, vi.z, vi.w}, xj[4] = {vj.x, vj.y, vj.z, vj.w}, xf[4] = {vf.x, vf.y, vf.z, vf.w},
                xo[4] = {vo.x, vo.y, vo.z, vo.w}, cc[4] = {c0.x, c0.y, c0.z, c0.w};
    float r[4], o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ni = (xi[j] - gm[0][j]) * gr[0][j] * ga[0][j] + gbt[0][j];
      const float nj = (xj[j] - gm[1][j]) * gr[1][j] * ga[1][j] + gbt[1][j];
      const float nf = (xf[j] - gm[2][j]) * gr[2][j] * ga[2][j] + gbt[2][j];
      o[j] = (xo[j] - gm[3][j]) * gr[3][j] * ga[3][j] + gbt[3][j];
      r[j] = cc[j] * sigmoidf_(nf + fb) + sigmoidf_(ni) * tanhf(nj);
      s[j] += (double)r[j];
      q[j] = fma((double)r[j], (double)r[j], q[j]);
    }
    cn[k] = make_float4(r[0], r[1], r[2], r[3]);
    on[k] = make_float4(o[0], o[1], o[2], o[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][4 + j] = q[j]; }
  __syncthreads();
  if (threadIdx.x < F4) {                            // fixed order over the 256 / F4 threads that share this channel quad
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = threadIdx.x; l < 256; l += F4)
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] += red[l][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) { part[f + j][0] = t[j]; part[f + j][1] = t[4 + j]; }
  }
  cluster.sync();                                    // every CTA's partials are visible cluster-wide
  if (threadIdx.x < F) {
    double ts = 0.0, tq = 0.0;
    for (int rr = 0; rr < CL; ++rr) {                // fixed rank order
      const double* rp = cluster.map_shared_rank(&part[0][0], rr);
      ts += rp[2 * threadIdx.x];
      tq += rp[2 * threadIdx.x + 1];
    }
    const double mean = ts / HW;
    double var = tq / HW - mean * mean;
    if (var < 0.0) var = 0.0;
    cst[threadIdx.x][0] = (float)mean;
    cst[threadIdx.x][1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  cluster.sync();                                    // all remote reads done (a CTA may exit from here on) + cst visible block-wide
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(cg + f)), b4 = __ldg(reinterpret_cast<const float4*>(cb + f));
  const float m0 = cst[f][0], r0 = cst[f][1], m1 = cst[f + 1][0], r1 = cst[f + 1][1];
  const float m2 = cst[f + 2][0], r2 = cst[f + 2][1], m3 = cst[f + 3][0], r3 = cst[f + 3][1];
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    const int idx = rank * per + k * 256 + threadIdx.x;
    const int pix = idx / F4;
    float4 cv, hv;
    cv.x = (cn[k].x - m0) * r0 * g4.x + b4.x; cv.y = (cn[k].y - m1) * r1 * g4.y + b4.y;
    cv.z = (cn[k].z - m2) * r2 * g4.z + b4.z; cv.w = (cn[k].w - m3) * r3 * g4.w + b4.w;
    *reinterpret_cast<float4*>(cb_ + (long long)pix * F + f) = cv;
    hv.x = tanhf(cv.x) * sigmoidf_(on[k].x); hv.y = tanhf(cv.y) * sigmoidf_(on[k].y);
    hv.z = tanhf(cv.z) * sigmoidf_(on[k].z); hv.w = tanhf(cv.w) * sigmoidf_(on[k].w);
    vst4(h, voff(h, b, pix) + f, hv);
    if (h2.p) {
      const int y = pix / W, x = pix - y * W;
      vst4(h2, voff(h2, b, (long long)(y >> 1) * (W >> 1) + (x >> 1)) + ((y & 1) * 2 + (x & 1)) * F + f, hv);
    }
  }
}

// PyTorch upsample_bilinear2d(align_corners=False, scale 2): src = max(0.5*(dst+0.5)-0.5, 0)
__device__ __forceinline__ void bil_idx(int d, int n, int& i0, int& i1, float& l0, float& l1) {
  float src = 0.5f * ((float)d + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + ((i0 < n - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__device__ __forceinline__ float4 bil4(float hl0, float hl1, float wl0, float wl1, float4 v00, float4 v01, float4 v10, float4 v11) {
  float4 o;
  o.x = hl0 * (wl0 * v00.x + wl1 * v01.x) + hl1 * (wl0 * v10.x + wl1 * v11.x);
  o.y = hl0 * (wl0 * v00.y + wl1 * v01.y) + hl1 * (wl0 * v10.y + wl1 * v11.y);
  o.z = hl0 * (wl0 * v00.z + wl1 * v01.z) + hl1 * (wl0 * v10.z + wl1 * v11.z);
  o.w = hl0 * (wl0 * v00.w + wl1 * v01.w) + hl1 * (wl0 * v10.w + wl1 * v11.w);
  return o;
}

// one thread = V (4 or 8) consecutive channels of one output pixel (both sources have C % V == 0); grid.y = sample
template <int V>
__global__ void __launch_bounds__(256) k_upsample2x(View s0, View s1, int H, int W, View out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int C = s0.C + s1.C, CV = C / V;
  const int Ho = 2 * H, Wo = 2 * W;
  const int total = Ho * Wo * CV;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (i % CV) * V;
    const int pix = i / CV;
    const int Y = pix / Wo, X = pix - Y * Wo;
    int y0, y1, x0, x1;
    float hl0, hl1, wl0, wl1;
    bil_idx(Y, H, y0, y1, hl0, hl1);
    bil_idx(X, W, x0, x1, wl0, wl1);
    const View& s = (c < s0.C) ? s0 : s1;
    const int cc = (c < s0.C) ? c : c - s0.C;
    if (V == 8) {
      const float8 v00 = vld8(s, voff(s, b, y0 * W + x0) + cc), v01 = vld8(s, voff(s, b, y0 * W + x1) + cc);
      const float8 v10 = vld8(s, voff(s, b, y1 * W + x0) + cc), v11 = vld8(s, voff(s, b, y1 * W + x1) + cc);
      float8 o;
      o.a = bil4(hl0, hl1, wl0, wl1, v00.a, v01.a, v10.a, v11.a);
      o.b = bil4(hl0, hl1, wl0, wl1, v00.b, v01.b, v10.b, v11.b);
      vst8(out, voff(out, b, pix) + c, o);
    } else {
      const float4 v00 = vld4(s, voff(s, b, y0 * W + x0) + cc), v01 = vld4(s, voff(s, b, y0 * W + x1) + cc);
      const float4 v10 = vld4(s, voff(s, b, y1 * W + x0) + cc), v11 = vld4(s, voff(s, b, y1 * W + x1) + cc);
      vst4(out, voff(out, b, pix) + c, bil4(hl0, hl1, wl0, wl1, v00, v01, v10, v11));
    }
  }
}

// Same arithmetic, one thread = the 2x2 OUTPUT block of input pixel (i, j) x 8 channels: the block only reads the 3x3 input
// neighbourhood (9 pixel loads instead of 16 for its 4 outputs) and the index arithmetic is paid once per block.  Horizontal
// interpolation first, then vertical — the operation order of bil4, so both kernels produce the same bits.
__device__ __forceinline__ float8 lerp8(float a, const float8& x, float b, const float8& y) {
  float8 o;
  o.a = make_float4(a * x.a.x + b * y.a.x, a * x.a.y + b * y.a.y, a * x.a.z + b * y.a.z, a * x.a.w + b * y.a.w);
  o.b = make_float4(a * x.b.x + b * y.b.x, a * x.b.y + b * y.b.y, a * x.b.z + b * y.b.z, a * x.b.w + b * y.b.w);
  return o;
}
__global__ void __launch_bounds__(256) k_upsample2x_blk(View s0, View s1, int H, int W, View out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int C = s0.C + s1.C, C8 = C >> 3;
  const int total = H * W * C8;
  const int Wo = 2 * W;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = (idx % C8) << 3;
    const int pix = idx / C8;
    const int i = pix / W, j = pix - i * W;
    const View& s = (c < s0.C) ? s0 : s1;
    const int cc = (c < s0.C) ? c : c - s0.C;
    const int ra = max(i - 1, 0), rc = min(i + 1, H - 1), ca = max(j - 1, 0), cd = min(j + 1, W - 1);
    // output rows 2i, 2i+1 (columns 2j, 2j+1): source indices and weights exactly as bil_idx computes them
    int ye0, ye1, yo0, yo1, xe0, xe1, xo0, xo1;
    float he0, he1, ho0, ho1, we0, we1, wo0, wo1;
    bil_idx(2 * i, H, ye0, ye1, he0, he1);
    bil_idx(2 * i + 1, H, yo0, yo1, ho0, ho1);
    bil_idx(2 * j, W, xe0, xe1, we0, we1);
    bil_idx(2 * j + 1, W, xo0, xo1, wo0, wo1);
    float8 hE[3], hO[3];                           // rows ra, i, rc interpolated to the even / odd output column
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int y = r == 0 ? ra : (r == 1 ? i : rc);
      const long long rowo = voff(s, b, (long long)y * W) + cc;
      const float8 vA = vld8(s, rowo + (long long)ca * s.pix_stride), vB = vld8(s, rowo + (long long)j * s.pix_stride);
      const float8 vC = vld8(s, rowo + (long long)cd * s.pix_stride);
      // even column: (xe0, xe1) is (j-1, j), or (0, 1) with weights (1, 0) at j == 0; odd column: (j, min(j+1, W-1))
      const float8& e0 = xe0 == j ? vB : vA;
      const float8& e1 = xe1 == j ? vB : (xe1 == cd ? vC : vA);
      hE[r] = lerp8(we0, e0, we1, e1);
      const float8& o1 = xo1 == j ? vB : vC;
      hO[r] = lerp8(wo0, vB, wo1, o1);
    }
    // vertical: even output row uses (ye0, ye1), odd row (yo0 = i, yo1)
    const int re0 = ye0 == i ? 1 : 0, re1 = ye1 == i ? 1 : (ye1 == rc ? 2 : 0), ro1 = yo1 == i ? 1 : 2;
    const long long o00 = voff(out, b, (long long)(2 * i) * Wo + 2 * j) + c;
    const long long orow = (long long)Wo * out.pix_stride;
    const float8 E0 = re0 == 1 ? hE[1] : hE[0], E1 = re1 == 1 ? hE[1] : (re1 == 2 ? hE[2] : hE[0]);
    const float8 F0 = re0 == 1 ? hO[1] : hO[0], F1 = re1 == 1 ? hO[1] : (re1 == 2 ? hO[2] : hO[0]);
    vst8(out, o00, lerp8(he0, E0, he1, E1));
    vst8(out, o00 + out.pix_stride, lerp8(he0, F0, he1, F1));
    const float8 G1 = ro1 == 1 ? hE[1] : hE[2], K1 = ro1 == 1 ? hO[1] : hO[2];
    vst8(out, o00 + orow, lerp8(ho0, hE[1], ho1, G1));
    vst8(out, o00 + orow + out.pix_stride, lerp8(ho0, hO[1], ho1, K1));
  }
}

// sa[m] = concat(action_tau, state_tau[, z_tau]);  gen_state = dense(concat(action, state))   (P1, P9)
__global__ void k_build_sa(SaArgs a, int M, int tau) {
  pdl_wait();
  pdl_trigger();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int A = a.adim + a.sdim + a.nz;
  float* sa = a.sa + (long long)m * A;
  float act[8], st[16];
  for (int i = 0; i < a.adim; ++i) {
    float v;
    if (tau < a.n_ctx_actions) v = a.ctx_actions[tau * a.adim + i];
    else v = a.actions[((long long)m * a.T + (tau - a.n_ctx_actions)) * a.adim + i];
    act[i] = v;
    sa[i] = v;
  }
  for (int i = 0; i < a.sdim; ++i) {
    const float v = (tau < a.C) ? a.ctx_states[tau * a.sdim + i] : a.state_cur[(long long)m * a.sdim + i];
    st[i] = v;
    sa[a.adim + i] = v;
  }
  if (a.w_z && a.zs && a.nz <= 16) {          // latent through a BasicLSTMCell(nz): forget bias 1, gate order i, j, f, o
    const int nz = a.nz;
    float in[32], g[64];
    float* zc = a.zstate + (long long)m * 2 * nz;
    float* zh = zc + nz;
    for (int i = 0; i < nz; ++i) { in[i] = a.zs[((long long)m * (a.P + a.C - 1) + tau) * nz + i]; in[nz + i] = zh[i]; }
    for (int j = 0; j < 4 * nz; ++j) {
      float acc = 0.f;
      for (int i = 0; i < 2 * nz; ++i) acc = fmaf(in[i], a.w_z[i * 4 * nz + j], acc);
      g[j] = acc + a.b_z[j];
    }
    for (int i = 0; i < nz; ++i) {
      const float cn = zc[i] * sigmoidf_(g[2 * nz + i] + 1.0f) + sigmoidf_(g[i]) * tanhf(g[nz + i]);
      const float hn = tanhf(cn) * sigmoidf_(g[3 * nz + i]);
      zc[i] = cn;
      zh[i] = hn;
      sa[a.adim + a.sdim + i] = hn;
    }
  } else {
    for (int i = 0; i < a.nz; ++i) sa[a.adim + a.sdim + i] = a.zs ? a.zs[((long long)m * (a.P + a.C - 1) + tau) * a.nz + i] : 0.f;
  }
  for (int j = 0; j < a.sdim; ++j) {
    float acc = 0.f;
    for (int i = 0; i < a.adim; ++i) acc = fmaf(act[i], a.w_state[i * a.sdim + j], acc);
    for (int i = 0; i < a.sdim; ++i) acc = fmaf(st[i], a.w_state[(a.adim + i) * a.sdim + j], acc);
    acc += a.b_state[j];
    a.state_cur[(long long)m * a.sdim + j] = acc;
    if (a.gen_states_all && tau >= a.C - 1) a.gen_states_all[((long long)m * a.P + (tau - (a.C - 1))) * a.sdim + j] = acc;
  }
}

// All S-1 steps of one rollout at once: the state recurrence (and the latent LSTM) depend on the actions only, never on the
// predicted frames, so every step's tiled vector is known before the first cell step.  Same arithmetic as k_build_sa, the
// recurrent values stay in registers.  sa_all[tau][m_stride rows][A].
__global__ void k_build_sa_all(SaArgs a, int M, int nsteps, long long step_stride) {
  pdl_wait();
  pdl_trigger();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int A = a.adim + a.sdim + a.nz;
  float st_cur[16], zc[16], zh[16];
  for (int i = 0; i < 16; ++i) { st_cur[i] = 0.f; zc[i] = 0.f; zh[i] = 0.f; }
  for (int tau = 0; tau < nsteps; ++tau) {
    float* sa = a.sa + tau * step_stride + (long long)m * A;
    float act[8], st[16];
    for (int i = 0; i < a.adim; ++i) {
      float v;
      if (tau < a.n_ctx_actions) v = a.ctx_actions[tau * a.adim + i];
      else v = a.actions[((long long)m * a.T + (tau - a.n_ctx_actions)) * a.adim + i];
      act[i] = v;
      sa[i] = v;
    }
    for (int i = 0; i < a.sdim; ++i) {
      const float v = (tau < a.C) ? a.ctx_states[tau * a.sdim + i] : st_cur[i];
      st[i] = v;
      sa[a.adim + i] = v;
    }
    if (a.w_z && a.zs && a.nz <= 16) {
      const int nz = a.nz;
      float in[32], g[64];
      for (int i = 0; i < nz; ++i) { in[i] = a.zs[((long long)m * (a.P + a.C - 1) + tau) * nz + i]; in[nz + i] = zh[i]; }
      for (int j = 0; j < 4 * nz; ++j) {
        float acc = 0.f;
        for (int i = 0; i < 2 * nz; ++i) acc = fmaf(in[i], a.w_z[i * 4 * nz + j], acc);
        g[j] = acc + a.b_z[j];
      }
      for (int i = 0; i < nz; ++i) {
        const float cn = zc[i] * sigmoidf_(g[2 * nz + i] + 1.0f) + sigmoidf_(g[i]) * tanhf(g[nz + i]);
        const float hn = tanhf(cn) * sigmoidf_(g[3 * nz + i]);
        zc[i] = cn;
        zh[i] = hn;
        sa[a.adim + a.sdim + i] = hn;
      }
    } else {
      for (int i = 0; i < a.nz; ++i) sa[a.adim + a.sdim + i] = a.zs ? a.zs[((long long)m * (a.P + a.C - 1) + tau) * a.nz + i] : 0.f;
    }
    for (int j = 0; j < a.sdim; ++j) {
      float acc = 0.f;
      for (int i = 0; i < a.adim; ++i) acc = fmaf(act[i], a.w_state[i * a.sdim + j], acc);
      for (int i = 0; i < a.sdim; ++i) acc = fmaf(st[i], a.w_state[(a.adim + i) * a.sdim + j], acc);
      acc += a.b_state[j];
      st_cur[j] = acc;
      if (a.gen_states_all && tau >= a.C - 1) a.gen_states_all[((long long)m * a.P + (tau - (a.C - 1))) * a.sdim + j] = acc;
    }
  }
}

__global__ void k_sabias(const float* __restrict__ sa, int A, const float* __restrict__ wcls,
                         const float* __restrict__ bias, int ncls, int Cout, int B, float* out) {
  const long long total = (long long)B * ncls * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Cout);
    const int cls = (int)((i / Cout) % ncls);
    const int b = (int)(i / ((long long)Cout * ncls));
    float acc = 0.f;
    for (int k = 0; k < A; ++k) acc = fmaf(sa[(long long)b * A + k], __ldg(wcls + ((long long)cls * A + k) * Cout + n), acc);
    out[i] = acc + (bias ? bias[n] : 0.f);
  }
}

// every layer's border-class bias of one cell step in ONE launch: grid (row blocks, layer, sample chunks).  A thread owns one
// (class, channel) row: its A weights stay in registers while it walks SB_CHUNK samples (sa staged in shared memory),
// stores coalesced along the row index.
constexpr int SB_CHUNK = 25;
__global__ void __launch_bounds__(256) k_sabias_batch(SabiasBatch a) {
  pdl_wait();
  pdl_trigger();
  const SabiasBatch::Layer L = a.L[blockIdx.y];
  const int per = L.ncls * L.Cout;
  const int nchunk = (a.B + SB_CHUNK - 1) / SB_CHUNK;
  const int step = blockIdx.z / nchunk;                       // hoisted form: every cell step of the rollout in one launch
  const int b0 = (blockIdx.z - step * nchunk) * SB_CHUNK, nb = min(SB_CHUNK, a.B - b0);
  __shared__ float ssa[SB_CHUNK][24];
  const float* sa_step = a.sa + step * a.sa_step;
  for (int i = threadIdx.x; i < nb * a.A; i += blockDim.x) ssa[i / a.A][i % a.A] = sa_step[(long long)b0 * a.A + i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const int cls = i / L.Cout, n = i - cls * L.Cout;
  const float* w = L.wcls + (cls * a.A) * L.Cout + n;
  float wr[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) wr[k] = k < a.A ? __ldg(w + k * L.Cout) : 0.f;
  const float bias = L.bias ? L.bias[n] : 0.f;
  float* o = L.out + step * L.out_step + (long long)b0 * per + i;
  for (int b = 0; b < nb; ++b) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 24; ++k)
      if (k < a.A) acc = fmaf(ssa[b][k], wr[k], acc);
    o[(long long)b * per] = acc + bias;
  }
}

// out[b, pix, 0..7] = (image rgb, first rgb, 0, 0): 8-channel (16-byte-unit) input of the first encoder conv
__global__ void k_pack_rgb2(View image, View first, int HW, View out) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const float* ip = vptr(image, b, pix);
  const float* fp = vptr(first, b, pix);
  const long long o = voff(out, b, pix);
  vst4(out, o, make_float4(__ldg(ip), __ldg(ip + 1), __ldg(ip + 2), __ldg(fp)));
  vst4(out, o + 4, make_float4(__ldg(fp + 1), __ldg(fp + 2), 0.f, 0.f));
}
__global__ void k_pack_fold(View image, View first, int H, int W, int kf, View out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W * kf) return;
  const int dx = i % kf, pix = i / kf;
  const int y = pix / W, x = pix - y * W, xs = x + dx - kf / 2;
  float8 v;
  v.a = make_float4(0.f, 0.f, 0.f, 0.f);
  v.b = v.a;
  if (xs >= 0 && xs < W) {
    const float* ip = vptr(image, b, (long long)y * W + xs);
    const float* fp = vptr(first, b, (long long)y * W + xs);
    v.a = make_float4(__ldg(ip), __ldg(ip + 1), __ldg(ip + 2), __ldg(fp));
    v.b = make_float4(__ldg(fp + 1), __ldg(fp + 2), 0.f, 0.f);
  }
  vst8(out, voff(out, b, pix) + dx * 8, v);
}
// Space-to-depth pack of (image, first) for the first encoder conv: out[b][Y][X][(sy*2+sx)*8 + c] = (image rgb, first rgb, 0, 0)
// at pixel (2Y+sy, 2X+sx).  A k x k SAME convolution followed by the 2x2 average pool is a 3x3 SAME convolution over these
// 2x2 blocks (k <= 5) with pre-averaged weights (engine.cu: prepare_conv), so the full-resolution conv output never exists.
__global__ void k_pack_s2d(View image, View first, int H, int W, View out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int H2 = H >> 1, W2 = W >> 1;
  if (i >= H2 * W2 * 4) return;
  const int sub = i & 3, blk = i >> 2;
  const int Y = blk / W2, X = blk - Y * W2;
  const long long pix = (long long)(2 * Y + (sub >> 1)) * W + (2 * X + (sub & 1));
  const float* ip = vptr(image, b, pix);
  const float* fp = vptr(first, b, pix);
  float8 v;
  v.a = make_float4(__ldg(ip), __ldg(ip + 1), __ldg(ip + 2), __ldg(fp));
  v.b = make_float4(__ldg(fp + 1), __ldg(fp + 2), 0.f, 0.f);
  vst8(out, voff(out, b, blk) + sub * 8, v);
}
// grid (chunks, buffers, row groups): every thread reads 16 bytes of row 0 once and stores them to its group's rows
__global__ void k_broadcast_rows(BroadcastBatch a, int M, int rows_per_group) {
  pdl_wait();
  pdl_trigger();
  const BroadcastBatch::Buf buf = a.b[blockIdx.y];
  const long long n16 = buf.row_bytes >> 4;
  const int first = buf.src ? 0 : 1;
  const int m0 = first + blockIdx.z * rows_per_group, m1 = min(M, m0 + rows_per_group);
  uint4* base = reinterpret_cast<uint4*>(buf.p);
  const uint4* src = buf.src ? reinterpret_cast<const uint4*>(buf.src) : base;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    for (int m = m0; m < m1; ++m) base[(long long)m * n16 + i] = v;
  }
}
__global__ void k_view_to_dense(View v, int HW, float* dst) {
  const int b = blockIdx.y;
  const long long total = (long long)HW * v.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[(long long)b * total + i] = vld1(v, voff(v, b, i / v.C) + (i % v.C));
}
__global__ void k_dense_to_view(const float* src, int HW, View v) {
  const int b = blockIdx.y;
  const long long total = (long long)HW * v.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    vst1(v, voff(v, b, i / v.C) + (i % v.C), src[(long long)b * total + i]);
}
__global__ void k_u8_to_f32(const uint8_t* in, float* out, long long n, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)in[i] / scale;
}
__global__ void k_fill(float* p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_onehot(float* d, int C, int ncam, int H, int W, int nd, const int* pix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * ncam * nd) return;
  const int p = i % nd, cam = (i / nd) % ncam, t = i / (nd * ncam);
  const int y = pix[(cam * nd + p) * 2], x = pix[(cam * nd + p) * 2 + 1];
  d[((((long long)t * ncam + cam) * H + y) * W + x) * nd + p] = 1.f;
}
__global__ void k_gather_rows(const float* src, long long row, const int* idx, int n, float* dst) {
  const long long total = (long long)n * row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / row);
    dst[i] = src[(long long)idx[r] * row + (i % row)];
  }
}

inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int launch_plane_stats(View x, int B, int H, int W, int pool, double* partial, cudaStream_t s, float* fin, int* cnt, float eps) {
  ++g_launch_counter;
  const int npix = H * W;
  int S = npix >= 2048 ? 8 : (npix >= 512 ? 4 : (npix >= 128 ? 2 : 1));
  while (S < STATS_MAX_SPLIT && (long long)B * ((x.C + 31) / 32) * S < 296 && npix / (2 * S) >= 16) S *= 2;   // fill the 148 SMs
  dim3 grid(B, (x.C + 31) / 32, S);                 // requires C % 4 == 0 and 16-byte aligned pixel rows (all conv outputs)
  FinArgs fa;
  fa.fin = ((x.C + 31) / 32 <= VF_STAT_CNT_STRIDE) ? fin : nullptr; fa.cnt = cnt; fa.eps = eps;
  if (!cnt) fa.fin = nullptr;
  launch_k(k_plane_stats, dim3(grid), dim3(256), 0, s, x, H, W, pool, partial, fa);
  return S;
}
size_t plane_stats_partial_doubles(int B, int C) { return (size_t)B * C * 16 * 2; }   // up to 16 slots per (sample, channel)
void launch_norm_act(View x, int B, int H, int W, int pool, StatsRef stats, const float* gamma,
                     const float* beta, int act, View y, cudaStream_t s, View y2) {
  ++g_launch_counter;
  // few, long-lived blocks per sample: every block stages the sample's statistics in shared memory first
  static const bool v8 = getenv("VF_NORM_ACT8") && atoi(getenv("VF_NORM_ACT8")) == 1;   // A/B switch
  auto al8 = [](const View& v) { return v.C == 0 || ((v.C | v.ch_off | v.pix_stride) % 8 == 0 && v.sample_stride % 8 == 0 && v.lo_off % 8 == 0); };
  if (v8 && x.C % 8 == 0 && x.C <= 512 && !x.lo_off && (x.pix_stride | x.ch_off) % 4 == 0 && x.sample_stride % 4 == 0 && al8(y) && al8(y2)) {
    dim3 grid8(grid_for((long long)H * W * (x.C >> 3), 256, B >= 64 ? 16 : 64), B);
    if (pool) launch_k(k_norm_act8<1>, dim3(grid8), dim3(256), 0, s, x, H, W, stats, gamma, beta, act, y, y2);
    else launch_k(k_norm_act8<0>, dim3(grid8), dim3(256), 0, s, x, H, W, stats, gamma, beta, act, y, y2);
    return;
  }
  dim3 grid(grid_for((long long)H * W * (x.C >> 2), 256, B >= 64 ? 16 : 64), B);
  launch_k(k_norm_act, dim3(grid), dim3(256), 0, s, x, H, W, pool, stats, gamma, beta, act, y, y2);
}
int launch_lstm_gates(View gates, int B, int HW, int F, StatsRef gstats, const float* gg, const float* gb,
                      float fb, float* c, double* partial, cudaStream_t s, float* fin, int* cnt, float eps) {
  ++g_launch_counter;
  int S = HW >= 1024 ? 8 : (HW >= 256 ? 4 : (HW >= 64 ? 2 : 1));      // pixel blocks per sample = stats partial slots
  dim3 grid(S, B);
  FinArgs fa;
  fa.fin = cnt ? fin : nullptr; fa.cnt = cnt; fa.eps = eps;
  launch_k(k_lstm_gates, dim3(grid), dim3(256), 0, s, gates, HW, F, gstats, gg, gb, fb, c, partial, fa);
  return S;
}
void launch_lstm_gates_generic(View gates, int B, int HW, int F, const float* gstats, const float* gg, const float* gb,
                               float fb, float* c, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid(grid_for((long long)HW * F, 256, 64), B);
  k_lstm_gates_generic<<<grid, 256, 0, s>>>(gates, HW, F, gstats, gg, gb, fb, c);
}
void launch_stats_finalize(const double* partial, int n, int S, int npix, float eps, float* stats, cudaStream_t s) {
  ++g_launch_counter;
  launch_k(k_stats_finalize, dim3((n + 255) / 256), dim3(256), 0, s, partial, n, S, npix, eps, stats);
}
void launch_lstm_out(View gates, int B, int HW, int F, StatsRef gstats, const float* gg, const float* gb,
                     StatsRef cstats, const float* cg, const float* cb, float* c, View h, cudaStream_t s, View h2, int W) {
  ++g_launch_counter;
  dim3 grid(grid_for((long long)HW * F / 4, 256, B >= 64 ? 16 : 64), B);
  launch_k(k_lstm_out, dim3(grid), dim3(256), 0, s, gates, HW, F, gstats, gg, gb, cstats, cg, cb, c, h, h2, W);
}
// returns false when the shape has no fused instance (the caller runs k_lstm_gates / k_lstm_out)
bool launch_lstm_fused(View gates, int B, int HW, int F, const float* gfin, const float* gg, const float* gb, float fb,
                       const float* cg, const float* cb, float eps, float* c, View h, cudaStream_t s, View h2, int W) {
  if ((F & 3) || F > 256 || 256 % (F >> 2) || !gfin) return false;
  const long long total4 = (long long)HW * (F >> 2);
  int CL = 0, nit = 0;
  const int cls[4] = {4, 2, 8, 1};
  for (int i = 0; i < 4 && !CL; ++i) {
    const long long per = total4 / cls[i];
    if (total4 % cls[i] || per % 256) continue;
    const int n = (int)(per / 256);
    if (n == 1 || n == 2 || n == 3 || n == 4 || n == 6 || n == 8) { CL = cls[i]; nit = n; }
  }
  if (!CL) return false;
  ++g_launch_counter;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CL, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
#define VF_LF(N) cudaLaunchKernelEx(&cfg, k_lstm_fused<N>, gates, HW, F, gfin, gg, gb, fb, cg, cb, eps, c, h, h2, W, CL)
  cudaError_t e;
  switch (nit) {
    case 1: e = VF_LF(1); break; case 2: e = VF_LF(2); break; case 3: e = VF_LF(3); break;
    case 4: e = VF_LF(4); break; case 6: e = VF_LF(6); break; default: e = VF_LF(8); break;
  }
#undef VF_LF
  return e == cudaSuccess;
}
void launch_upsample2x(View s0, View s1, int B, int H, int W, View out, cudaStream_t s) {
  ++g_launch_counter;
  auto al8 = [](const View& v) { return v.C == 0 || ((v.C | v.ch_off | v.pix_stride) % 8 == 0 && v.sample_stride % 8 == 0 && v.lo_off % 8 == 0); };
  static const bool blk = getenv("VF_UPSAMPLE_BLK") && atoi(getenv("VF_UPSAMPLE_BLK")) == 1;   // A/B switch; measured no gain -> off
  if (al8(s0) && al8(s1) && al8(out) && blk && H >= 2 && W >= 2) {
    dim3 grid(grid_for((long long)H * W * ((s0.C + s1.C) >> 3), 256, 64), B);
    launch_k(k_upsample2x_blk, dim3(grid), dim3(256), 0, s, s0, s1, H, W, out);
  } else if (al8(s0) && al8(s1) && al8(out)) {
    dim3 grid(grid_for((long long)4 * H * W * ((s0.C + s1.C) >> 3), 256, 64), B);
    launch_k(k_upsample2x<8>, dim3(grid), dim3(256), 0, s, s0, s1, H, W, out);
  } else {
    dim3 grid(grid_for((long long)4 * H * W * ((s0.C + s1.C) >> 2), 256, 64), B);
    launch_k(k_upsample2x<4>, dim3(grid), dim3(256), 0, s, s0, s1, H, W, out);
  }
}
void launch_build_sa(const SaArgs& a, int M, int tau, cudaStream_t s) {
  ++g_launch_counter;
  launch_k(k_build_sa, dim3((M + 127) / 128), dim3(128), 0, s, a, M, tau);
}
void launch_build_sa_all(const SaArgs& a, int M, int nsteps, long long step_stride, cudaStream_t s) {
  ++g_launch_counter;
  launch_k(k_build_sa_all, dim3((M + 127) / 128), dim3(128), 0, s, a, M, nsteps, step_stride);
}
void launch_sabias_batch(const SabiasBatch& a, cudaStream_t s) {
  if (a.n == 0) return;
  ++g_launch_counter;
  if (a.A > 24) return;                                      // adim + sdim + nz <= 24 (vf_create)
  int mx = 1;
  for (int i = 0; i < a.n; ++i) mx = std::max(mx, a.L[i].ncls * a.L[i].Cout);
  dim3 grid((mx + 255) / 256, a.n, ((a.B + SB_CHUNK - 1) / SB_CHUNK) * (a.nsteps > 0 ? a.nsteps : 1));
  launch_k(k_sabias_batch, dim3(grid), dim3(256), 0, s, a);
}
void launch_sabias(const float* sa, int A, const float* wcls, const float* bias, int ncls, int Cout, int B,
                   float* out, cudaStream_t s) {
  ++g_launch_counter;
  k_sabias<<<grid_for((long long)B * ncls * Cout), 256, 0, s>>>(sa, A, wcls, bias, ncls, Cout, B, out);
}
void launch_view_to_dense(View v, int B, int HW, float* dst, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid(grid_for((long long)HW * v.C, 256, 64), B);
  k_view_to_dense<<<grid, 256, 0, s>>>(v, HW, dst);
}
void launch_dense_to_view(const float* src, int B, int HW, View v, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid(grid_for((long long)HW * v.C, 256, 64), B);
  k_dense_to_view<<<grid, 256, 0, s>>>(src, HW, v);
}
void launch_pack_fold(View image, View first, int B, int H, int W, int kf, View out, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid((H * W * kf + 255) / 256, B);
  launch_k(k_pack_fold, dim3(grid), dim3(256), 0, s, image, first, H, W, kf, out);
}
void launch_pack_s2d(View image, View first, int B, int H, int W, View out, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid(((H / 2) * (W / 2) * 4 + 255) / 256, B);
  launch_k(k_pack_s2d, dim3(grid), dim3(256), 0, s, image, first, H, W, out);
}
void launch_broadcast_rows(const BroadcastBatch& a, int M, cudaStream_t s) {
  if (a.n == 0 || M < 1) return;
  bool any_src = false;
  long long mx = 16;
  for (int i = 0; i < a.n; ++i) { mx = std::max(mx, a.b[i].row_bytes); any_src = any_src || a.b[i].src; }
  if (M <= 1 && !any_src) return;
  ++g_launch_counter;
  const int rows_per_group = 8;
  dim3 grid((unsigned)std::min<long long>((mx / 16 + 255) / 256, 64), a.n, (M + rows_per_group - 1) / rows_per_group);
  launch_k(k_broadcast_rows, dim3(grid), dim3(256), 0, s, a, M, rows_per_group);
}
void launch_pack_rgb2(View image, View first, int B, int HW, View out, cudaStream_t s) {
  ++g_launch_counter;
  dim3 grid((HW + 255) / 256, B);
  k_pack_rgb2<<<grid, 256, 0, s>>>(image, first, HW, out);
}
void launch_u8_to_f32(const uint8_t* in, float* out, long long n, float scale, cudaStream_t s) {
  ++g_launch_counter;
  k_u8_to_f32<<<grid_for(n), 256, 0, s>>>(in, out, n, scale);
}
void launch_fill(float* p, long long n, float v, cudaStream_t s) {
  ++g_launch_counter;
  k_fill<<<grid_for(n), 256, 0, s>>>(p, n, v);
}
void launch_onehot(float* d, int C, int ncam, int H, int W, int nd, const int* pix, cudaStream_t s) {
  ++g_launch_counter;
  k_onehot<<<(C * ncam * nd + 63) / 64, 64, 0, s>>>(d, C, ncam, H, W, nd, pix);
}
void launch_gather_rows(const float* src, long long row, const int* idx, int n, float* dst, cudaStream_t s) {
  ++g_launch_counter;
  k_gather_rows<<<grid_for((long long)n * row), 256, 0, s>>>(src, row, idx, n, dst);
}

}  // namespace vf

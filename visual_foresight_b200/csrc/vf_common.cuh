// vf_common.cuh — shared types and launcher declarations of the vfengine CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

namespace vf {

// NHWC activation view: element (b, y, x, c) lives at
//   p[b * sample_stride + (y * W + x) * pix_stride + ch_off + c]
// sample_stride == 0 broadcasts one image to every sample (context frames are shared by all
// action samples — the reference tf.tile()s them per tower, setup_predictor.py:40-44).
//
// Storage format.  lo_off == 0: float32 elements.  lo_off > 0: "split-half" storage — the buffer holds two fp16
// planes with the same element indexing, hi = rn16(x) at [i] and lo = rn16(x - hi) at [i + lo_off] (p reinterpreted
// as __half*; same bytes as float32).  Every tensor that feeds a convolution is kept split on the tensor-core
// path: the producer kernel performs the fp32 -> (hi, lo) split once, and the tcgen05 convolution fetches both planes
// with TMA straight into its swizzled shared-memory operand buffers (no conversion pass, no register staging).
struct View {
  float* p;
  long long sample_stride;
  int pix_stride;
  int ch_off;
  int C;
  long long lo_off;
};

__host__ __device__ inline View make_view(float* p, long long ss, int ps, int co, int C, long long lo_off = 0) {
  View v; v.p = p; v.sample_stride = ss; v.pix_stride = ps; v.ch_off = co; v.C = C; v.lo_off = lo_off; return v;
}

#ifdef __CUDACC__
// Programmatic dependent launch: every rollout kernel starts with pdl_wait() (blocks until the kernels it depends on have
// completed and flushed) and pdl_trigger() (lets the next kernel's launch + prologue overlap this kernel).  Without the launch
// attribute both are no-ops, so the kernels behave identically under plain stream ordering.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// pdl_trigger_conv(): the persistent one-wave convolution kernels (their dependents can only take SMs a finished CTA has left).
// Experiments (profiles/build_variant.sh): -DVF_PDL_LATE = no early trigger anywhere (dependents are released when the blocks
// exit); -DVF_PDL_CONV_ONLY = only the convolution kernels trigger early (a waiting 200 KB convolution CTA must not take an SM
// from the later waves of a multi-wave pointwise kernel).
#if defined(VF_PDL_LATE)
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_trigger_conv() {}
#elif defined(VF_PDL_CONV_ONLY)
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_trigger_conv() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger_conv() { pdl_trigger(); }
#endif
extern bool g_use_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// element offset of (b, pix, channel 0 of the view)
__device__ __forceinline__ long long voff(const View& v, int b, long long pix) {
  return (long long)b * v.sample_stride + pix * v.pix_stride + v.ch_off;
}
__device__ __forceinline__ void split_half(float x, __half& h, __half& l) {
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}
__device__ __forceinline__ float vld1(const View& v, long long off) {
  if (v.lo_off) {
    const __half* hp = reinterpret_cast<const __half*>(v.p);
    return __half2float(__ldg(hp + off)) + __half2float(__ldg(hp + off + v.lo_off));
  }
  return __ldg(v.p + off);
}
// 4 consecutive channels, off % 4 == 0 (and lo_off % 4 == 0)
__device__ __forceinline__ float4 vld4(const View& v, long long off) {
  if (v.lo_off) {
    const __half* hp = reinterpret_cast<const __half*>(v.p);
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(hp + off));
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(hp + off + v.lo_off));
    const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
    return make_float4(a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y);
  }
  return __ldg(reinterpret_cast<const float4*>(v.p + off));
}
// 8 consecutive channels, off % 8 == 0 (one 16-byte access per fp16 plane)
struct float8 { float4 a, b; };
__device__ __forceinline__ float8 vld8(const View& v, long long off) {
  float8 r;
  if (v.lo_off) {
    const __half* hp = reinterpret_cast<const __half*>(v.p);
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(hp + off));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(hp + off + v.lo_off));
    const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&h.z)), h3 = __half22float2(*reinterpret_cast<const __half2*>(&h.w));
    const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
    const float2 l2 = __half22float2(*reinterpret_cast<const __half2*>(&l.z)), l3 = __half22float2(*reinterpret_cast<const __half2*>(&l.w));
    r.a = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
    r.b = make_float4(h2.x + l2.x, h2.y + l2.y, h3.x + l3.x, h3.y + l3.y);
  } else {
    r.a = __ldg(reinterpret_cast<const float4*>(v.p + off));
    r.b = __ldg(reinterpret_cast<const float4*>(v.p + off + 4));
  }
  return r;
}
__device__ __forceinline__ void vst8(const View& v, long long off, const float8& x) {
  if (v.lo_off) {
    __half* hp = reinterpret_cast<__half*>(v.p);
    __align__(16) __half h[8], l[8];
    split_half(x.a.x, h[0], l[0]); split_half(x.a.y, h[1], l[1]); split_half(x.a.z, h[2], l[2]); split_half(x.a.w, h[3], l[3]);
    split_half(x.b.x, h[4], l[4]); split_half(x.b.y, h[5], l[5]); split_half(x.b.z, h[6], l[6]); split_half(x.b.w, h[7], l[7]);
    *reinterpret_cast<uint4*>(hp + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(hp + off + v.lo_off) = *reinterpret_cast<const uint4*>(l);
  } else {
    *reinterpret_cast<float4*>(v.p + off) = x.a;
    *reinterpret_cast<float4*>(v.p + off + 4) = x.b;
  }
}
__device__ __forceinline__ void vst1(const View& v, long long off, float x) {
  if (v.lo_off) {
    __half* hp = reinterpret_cast<__half*>(v.p);
    __half h, l;
    split_half(x, h, l);
    hp[off] = h;
    hp[off + v.lo_off] = l;
  } else {
    v.p[off] = x;
  }
}
__device__ __forceinline__ void vst4(const View& v, long long off, float4 x) {
  if (v.lo_off) {
    __half* hp = reinterpret_cast<__half*>(v.p);
    __half h[4], l[4];
    split_half(x.x, h[0], l[0]); split_half(x.y, h[1], l[1]); split_half(x.z, h[2], l[2]); split_half(x.w, h[3], l[3]);
    *reinterpret_cast<uint2*>(hp + off) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(hp + off + v.lo_off) = *reinterpret_cast<const uint2*>(l);
  } else {
    *reinterpret_cast<float4*>(v.p + off) = x;
  }
}
#endif

// border class of pixel coordinate y for a SAME zero-padded k-tap filter (k odd): which taps fall
// inside the image.  Classes 0..pad-1 = first rows, pad = interior, pad+1..k-1 = last rows.
__host__ __device__ inline int border_class(int y, int H, int pad) {
  if (y < pad) return y;
  if (y >= H - pad) return 2 * pad - (H - 1 - y);
  return pad;
}

struct ConvArgs {
  View src0, src1;        // input channels = src0.C + src1.C (src1.C may be 0)
  const float* w;         // [k*k][Cin][Cout]
  const float* bias;      // [Cout] or null
  const float* sabias;    // [B][k*k classes][Cout] (includes bias) or null
  View out;               // raw output
  int H, W, Cin, Cout, k;
  int act;                // 0 none, 1 sigmoid (small kernel only)
};

enum { ACT_NONE = 0, ACT_SIGMOID = 1, ACT_RELU = 2 };

// ---- convolution (SIMT fp32) ---------------------------------------------------------------
void launch_conv_simt(const ConvArgs& a, int B, cudaStream_t s);

// ---- normalisation / pointwise --------------------------------------------------------------
// Instance-norm statistics are kept as float64 partial sums partial[((b*C + c)*S + k)*2 + {sum, sumsq}] (k < S slots,
// written by k_plane_stats, by the tensor-core conv epilogue or by k_lstm_gates); every consumer finalises
// (mean, rstd) of the planes it needs in its prologue, in a fixed order (bit-reproducible).
constexpr int VF_STAT_CNT_STRIDE = 32;   // arrival counters per sample of the "last arriver finalises" statistics (channel groups)
struct StatsRef {
  const float* fin;       // finalised (mean, rstd) pairs (k_stats_finalize), or null: finalise from the partials on the fly
  const double* partial;
  int S;        // partial slots per (sample, channel)
  int npix;     // pixels per plane
  float eps;
};
inline StatsRef stats_ref(const double* partial, int S, int npix, float eps, const float* fin = nullptr) {
  StatsRef r; r.fin = fin; r.partial = partial; r.S = S; r.npix = npix; r.eps = eps; return r;
}
// partial sums of the (optionally 2x2 avg-pooled) planes of x; returns the slot count S
// fin / cnt (optional): the last block of a sample also writes the finalised (mean, rstd) pairs into fin (cnt: zeroed arrival
// counters [B][VF_STAT_CNT_STRIDE], left zeroed) — no k_stats_finalize launch needed
int launch_plane_stats(View x, int B, int H, int W, int pool, double* partial, cudaStream_t s, float* fin = nullptr, int* cnt = nullptr,
                       float eps = 1e-6f);
size_t plane_stats_partial_doubles(int B, int C);
// y = act((pool(x) - mean) * rstd * gamma + beta)
// y holds channels [0, y.C); the optional second view y2 the remaining x.C - y.C channels
void launch_norm_act(View x, int B, int H, int W, int pool, StatsRef stats, const float* gamma,
                     const float* beta, int act, View y, cudaStream_t s, View y2 = make_view(nullptr, 0, 0, 0, 0));
// conv-LSTM pointwise, part 1: c <- c*sigmoid(f+fb) + sigmoid(i)*tanh(j) with gates instance-normalised
// also accumulates the instance-norm partial sums of the new cell state; returns the partial slots per (sample, channel)
int launch_lstm_gates(View gates, int B, int HW, int F, StatsRef gstats, const float* ggamma,
                      const float* gbeta, float forget_bias, float* c, double* partial, cudaStream_t s, float* fin = nullptr,
                      int* cnt = nullptr, float eps = 1e-6f);
void launch_lstm_gates_generic(View gates, int B, int HW, int F, const float* gstats, const float* ggamma,
                               const float* gbeta, float forget_bias, float* c, cudaStream_t s);
void launch_stats_finalize(const double* partial, int n, int S, int npix, float eps, float* stats, cudaStream_t s);
// part 2: c <- IN(c);  h <- tanh(c) * sigmoid(IN(o))
void launch_lstm_out(View gates, int B, int HW, int F, StatsRef gstats, const float* ggamma,
                     const float* gbeta, StatsRef cstats, const float* cgamma, const float* cbeta,
                     float* c, View h, cudaStream_t s, View h2 = make_view(nullptr, 0, 0, 0, 0), int W = 0);   // h2: optional
                     // space-to-depth copy of h ([B][H/2*W/2][4F]: 2x2 pixel blocks in channels) for a pool-fused encoder conv
// parts 1 + 2 in one cluster kernel (one cluster of CTAs per sample, cell statistics exchanged through distributed shared
// memory); gfin: finalised (mean, rstd) pairs of the gates.  false = no instance for this shape (use the two kernels above)
bool launch_lstm_fused(View gates, int B, int HW, int F, const float* gfin, const float* ggamma, const float* gbeta, float forget_bias,
                       const float* cgamma, const float* cbeta, float eps, float* c, View h, cudaStream_t s,
                       View h2 = make_view(nullptr, 0, 0, 0, 0), int W = 0);
// out[b, 2H, 2W, C0+C1] = bilinear_x2(concat(src0, src1))   (half-pixel centres, edge clamp)
void launch_upsample2x(View src0, View src1, int B, int H, int W, View out, cudaStream_t s);

// ---- action/state vector and its per-layer border-class bias ----------------------------------
struct SaArgs {
  const float* actions;    // [M][T][adim] local samples
  int T, adim, sdim, nz, n_ctx_actions, C;
  const float* ctx_actions;  // [n_ctx_actions][adim]
  const float* ctx_states;   // [C][sdim]
  const float* zs;           // [M][S-1][nz] or null
  const float* w_z;          // [(2*nz)][4*nz] dense LSTM over the latent (use_rnn_z) or null: z is tiled directly
  const float* b_z;          // [4*nz]
  float* zstate;             // [M][2*nz] (c, h), zero at the start of a rollout
  const float* w_state;      // [(adim+sdim)][sdim]
  const float* b_state;      // [sdim]
  float* state_cur;          // [M][sdim]  (gen_state of the previous step, in/out)
  float* sa;                 // [M][A] out
  float* gen_states_all;     // [M][P][sdim] or null
  int P;
};
void launch_build_sa(const SaArgs& a, int M, int tau, cudaStream_t s);
struct SabiasBatch {
  struct Layer { const float* wcls; const float* bias; float* out; int ncls, Cout; long long out_step; };   // out_step: elements between steps
  Layer L[24];
  int n, A, B;
  const float* sa;
  int nsteps = 1;            // > 1: all cell steps of a rollout in one launch (sa / out advance by sa_step / out_step per step)
  long long sa_step = 0;
};
// every step's tiled action/state vector at once: sa_all[tau * step_stride + m * A] (the recurrences do not depend on frames)
void launch_build_sa_all(const SaArgs& a, int M, int nsteps, long long step_stride, cudaStream_t s);
void launch_sabias_batch(const SabiasBatch& a, cudaStream_t s);
// rows[m] = rows[0] for m in [1, M): replicates the recurrent state of sample 0 after the shared-prefix cell steps (engine.cu)
struct BroadcastBatch {
  struct Buf { void* p; long long row_bytes; const void* src; };   // row_bytes % 16 == 0, rows contiguous; src == nullptr: rows[m] = rows[0]
                                                                     // for m >= 1, else rows[m] = src row for every m >= 0
  Buf b[24];
  int n;
};
void launch_broadcast_rows(const BroadcastBatch& a, int M, cudaStream_t s);
// sabias[b][cls][n] = bias[n] + sum_a sa[b][a] * wcls[cls][a][n]
void launch_sabias(const float* sa, int A, const float* wcls, const float* bias, int ncls, int Cout,
                   int B, float* out, cudaStream_t s);

// ---- CDNA ------------------------------------------------------------------------------------
// CDNA head (spec P5): split-K partial products part[ks][b][128] of dense(feat); the consumers combine them, add the
// identity tap, relu-shift and L1-normalise in their prologue (cdna_finalize)
size_t cdna_partial_floats(int K, int B);
void launch_cdna_kernels(View feat, int npix, const float* w, int ksize, int nt, int B, float* part, cudaStream_t s);
// layers[.., 3n + c] = T_n(image)[c]; then prev image, first image   (spec P6); also writes kern[b][n][k*k]
void launch_cdna_apply(View image, View first, const float* part, int K, const float* bias, float* kern, int ksize, int nt,
                       int B, int H, int W, View layers, cudaStream_t s);
struct CompositeArgs {
  View logits;            // [B,H,W,n_masks]
  View layers;            // [B,H,W,>= 3*n_masks]: T_0..T_nt-1, prev, first, scratch (rgb each)
  View prev_d, first_d;   // distributions [.,H,W,nd]
  const float* kern;      // [B][nt][k*k]
  View gen_image;         // out [B,H,W,3]
  View gen_distrib;       // out raw [B,H,W,nd]
  float* partial;         // [B][nd][nblk] plane sums of the raw distribution
  int nt, ksize, nd, H, W;
};
int composite_blocks(int H, int W);
void launch_composite(const CompositeArgs& a, int B, cudaStream_t s);
// distrib /= sum_hw(distrib)   (renormalize_pixdistrib)
void launch_distrib_normalize(View d, const float* partial, int nblk, int B, int H, int W, int nd, cudaStream_t s);

// ---- costs -------------------------------------------------------------------------------------
// cost[m][t][task] = sum(p*d)/sum(p) on distrib (M,P,ncam,H,W,nd);  goal (ncam,nd,2) doubles
void launch_pixel_cost(const float* distrib, int M, int P, int ncam, int H, int W, int nd,
                       const double* goal, float* cost, cudaStream_t s);
// scores[m] = sum_task w_task * (sum_t cost*t_mult) / sum(t_mult)     (float64 like numpy>=2)
void launch_score_final(const float* cost, int M, int P, int ntask, const double* task_w, double finalweight,
                        double* scores, cudaStream_t s);
// scores[m] = mean((frames[m, P-1, cam0] - goal)^2)
void launch_goal_image_cost(const float* frames, int M, int P, int ncam, int H, int W, const float* goal,
                            double* scores, cudaStream_t s);

// ---- CEM ---------------------------------------------------------------------------------------
struct SampleArgs {
  int D, nactions, adim, repeat, K;   // adim = SAMPLED dims per step, D = nactions * adim; K = columns of the factor (0 -> diagonal, iteration 0)
  int adim_out;                       // dims per step of the action rows (adim + appended constants)
  double append[8];                   // the adim_out - adim constant dims (append_action)
  unsigned discrete_mask;             // bit a: floor + clip to [0,4] (discrete_ind)
  int kind;                           // 0 Gaussian (mean + std0*z | mean + factor z), 1 correlated noise (AR(1) around mean)
  double beta0, beta1, bias[8];       // correlated noise
  const double* mean;      // [D]
  const double* factor;    // [D][K]  (iteration > 0)
  const double* std0;      // [D]     (iteration 0)
  const float* noise;      // external standard normals for this iteration or null -> Philox
  int noise_stride;        // doubles per sample in the noise row
  const int* indices;      // global sample indices to generate (null -> offset + i)
  int offset;
  double clip_lo[8], clip_hi[8];
  uint64_t seed; uint32_t plan_index; uint32_t iteration;
  double* out_nr;          // [n][D] non-repeated actions (float64, post clip)
  float* out_actions;      // [n][T][adim] repeated, f32 (network input) or null
  double* out_actions64;   // [n][T][adim] or null
};
void launch_sample_actions(const SampleArgs& a, int n, cudaStream_t s);
// zs[n][steps][nz] standard normals, Philox keyed by the global rollout index goff + r (j word: 0x80000000 | (step*nz + i))
void launch_sample_latents(float* zs, int n, int steps, int nz, int goff, uint64_t seed, uint32_t plan, uint32_t iter,
                           cudaStream_t st);
// out[m] = mean_k s[m*K+k] + lambda * var_k
void launch_reduce_futures(const double* s, int M, int K, double lambda, double* out, cudaStream_t st);
// stable ascending top-k (ties -> lower index, NaN last) == np.argsort(kind='stable')[:k]
void launch_topk(const double* scores, int n, int k, int* out_idx, double* work_keys, int* work_idx, cudaStream_t s);
int topk_padded(int n);
// mean[D], factor[D][K] = Xc^T / sqrt(K-1), cov[D][D] (unbiased)
void launch_refit(const double* elites_nr, int K, int D, double* mean, double* factor, double* cov, cudaStream_t s);
// correlated-noise sampler: mean[d] = sum_k S_k x[k][d] / (sum_k S_k + 1e-4), S_k = exp(kappa * (r_k - max r)), r_k = -scores[idx[k]]
void launch_refit_correlated(const double* elites_nr, const double* scores, const int* idx, int K, int D, double kappa, double* mean,
                             cudaStream_t s);

// score exchange over peer memory (engine.cu: vf_cem_exchange): store my segment of a score row into every peer's window,
// publish my arrival counter, wait for the peers' counters
constexpr int VF_MAX_WORLD = 16;
struct ExchangeArgs {
  double* scores[VF_MAX_WORLD];      // score matrix of every rank's window (this rank's own included)
  unsigned* flags[VF_MAX_WORLD];     // arrival counters of every rank's window: flags[r][q] = last epoch rank q published to r
  unsigned* status;                  // this rank's window status word (set to 1 on timeout)
  int world, rank;
  long long row_off;                 // iteration * global_samples
  int offset, local;
  unsigned epoch;
  unsigned long long timeout_ns;
};
void launch_score_exchange(const ExchangeArgs& a, cudaStream_t s);

void launch_pack_rgb2(View image, View first, int B, int HW, View out, cudaStream_t s);
// out[b, (y,x), dx*8 + 0..7] = (image rgb, first rgb, 0, 0) at (y, x + dx - kf/2), zeros outside the image: the kf dx taps of
// the first encoder conv folded into channels (its tensor-core form is then a kf x 1 convolution over 8*kf channels)
void launch_pack_fold(View image, View first, int B, int H, int W, int kf, View out, cudaStream_t s);
void launch_pack_s2d(View image, View first, int B, int H, int W, View out, cudaStream_t s);
// dst[b][pix][c] (dense float32) = view element (either storage format)
void launch_view_to_dense(View v, int B, int HW, float* dst, cudaStream_t s);
void launch_dense_to_view(const float* src, int B, int HW, View v, cudaStream_t s);

// ---- misc --------------------------------------------------------------------------------------
void launch_u8_to_f32(const uint8_t* in, float* out, long long n, float scale, cudaStream_t s);
void launch_fill(float* p, long long n, float v, cudaStream_t s);
void launch_onehot(float* distrib, int C, int ncam, int H, int W, int nd, const int* pix, cudaStream_t s);
void launch_gather_rows(const float* src, long long row, const int* idx, int n, float* dst, cudaStream_t s);

extern long long g_launch_counter;   // incremented by every launcher

}  // namespace vf

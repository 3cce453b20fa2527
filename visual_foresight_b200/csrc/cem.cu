// cem.cu — CEM sampler / elite selection / Gaussian refit kernels (float64, tiny, latency-bound).
//
// Sampling law.  The reference draws x ~ N(mu, Sigma) with np.random.multivariate_normal
// (gaussian_sampler.py:82).  Iteration 0 has a diagonal Sigma (construct_initial_sigma,
// controller_utils.py:47-84) so x = mu + sigma .* z.  After a refit on K elites,
// Sigma = Xc^T Xc / (K-1) (np.cov, gaussian_sampler.py:101) has rank <= K-1, and
// x = mu + (Xc^T / sqrt(K-1)) z with z in R^K has exactly that law — no factorisation needed.
// Noise is Philox4x32-10 keyed by (seed; global sample index, draw index, iteration, plan index), so a
// sample's actions do not depend on how samples are sharded over GPUs; any rank can regenerate
// any elite's action row, which is what makes the refit communication-free.
#include "vf_common.cuh"

namespace vf {
namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// standard normal #j of (sample, iteration, plan): Box-Muller on two 32-bit uniforms, float64
__device__ double philox_normal(uint64_t seed, uint32_t plan, uint32_t iter, uint32_t sample, uint32_t j) {
  uint32_t c[4] = {sample, j >> 1, iter, plan};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u1 = ((double)c[0] + 0.5) * (1.0 / 4294967296.0);
  const double u2 = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
  const double r = sqrt(-2.0 * log(u1));
  const double th = 6.283185307179586476925286766559 * u2;
  return (j & 1) ? r * sin(th) : r * cos(th);
}

// latent z[r][step][i] ~ N(0,1) for rollout sample r (global index goff + r): the Philox stream of the action noise with
// bit 31 of the draw-index word set, so latents never collide with action draws
__global__ void k_sample_latents(float* zs, int n, int steps, int nz, int goff, uint64_t seed, uint32_t plan, uint32_t iter) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * steps * nz) return;
  const int r = tid / (steps * nz), j = tid - r * steps * nz;
  zs[tid] = (float)philox_normal(seed, plan, iter, (uint32_t)(goff + r), 0x80000000u | (uint32_t)j);
}

// scores[m] = mean_k s[m*K + k] + lambda * var_k (population variance, np.var); fixed summation order
__global__ void k_reduce_futures(const double* __restrict__ s, int M, int K, double lambda, double* out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double mean = 0.0;
  for (int k = 0; k < K; ++k) mean += s[(long long)m * K + k];
  mean /= (double)K;
  double var = 0.0;
  for (int k = 0; k < K; ++k) { const double d = s[(long long)m * K + k] - mean; var += d * d; }
  out[m] = mean + lambda * (var / (double)K);
}

// thread per (row i, coordinate d)
__global__ void k_sample_actions(SampleArgs a, int n) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * a.D) return;
  const int i = tid / a.D, d = tid % a.D;
  const int gidx = a.indices ? a.indices[i] : a.offset + i;
  double x = a.mean[d];
  if (a.K == 0) {
    const double z = a.noise ? (double)a.noise[(long long)gidx * a.noise_stride + d]
                             : philox_normal(a.seed, a.plan_index, a.iteration, (uint32_t)gidx, (uint32_t)d);
    x += a.std0[d] * z;
  } else {
    for (int k = 0; k < a.K; ++k) {
      const double z = a.noise ? (double)a.noise[(long long)gidx * a.noise_stride + k]
                               : philox_normal(a.seed, a.plan_index, a.iteration, (uint32_t)gidx, (uint32_t)k);
      x += a.factor[d * a.K + k] * z;
    }
  }
  const int ad = d % a.adim, na = d / a.adim;
  if ((a.discrete_mask >> ad) & 1u) x = fmin(fmax(floor(x), 0.0), 4.0);   // discretize (controller_utils.py:107-117), before the clip
  x = fmin(fmax(x, a.clip_lo[ad]), a.clip_hi[ad]);       // truncate_movement
  a.out_nr[(long long)i * a.D + d] = x;
  const int T = a.nactions * a.repeat;
  for (int r = 0; r < a.repeat; ++r) {                     // np.repeat(actions, repeat, axis=1)
    const long long o = ((long long)i * T + na * a.repeat + r) * a.adim_out + ad;
    if (a.out_actions) a.out_actions[o] = (float)x;
    if (a.out_actions64) a.out_actions64[o] = x;
    if (ad == 0)                                           // append_action: constant trailing dims of the step
      for (int e = a.adim; e < a.adim_out; ++e) {
        if (a.out_actions) a.out_actions[o + e] = (float)a.append[e - a.adim];
        if (a.out_actions64) a.out_actions64[o + e] = a.append[e - a.adim];
      }
  }
}

// CorrelatedNoiseSampler (samplers/correlated_noise.py:17-35): thread per (row i, action dim ad), sequential over the steps.
// noise_s = z_s * std + bias; out_s = beta0 * noise_s + beta1 * out_{s-1}, out_{-1} := noise_{last} (the reference's wrap);
// action_s = mean_s + out_s.  repeat == 1.
__global__ void k_sample_correlated(SampleArgs a, int n) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n * a.adim) return;
  const int i = tid / a.adim, ad = tid % a.adim;
  const int gidx = a.indices ? a.indices[i] : a.offset + i;
  auto draw = [&](int s) {
    const int j = s * a.adim + ad;
    const double z = a.noise ? (double)a.noise[(long long)gidx * a.noise_stride + j]
                             : philox_normal(a.seed, a.plan_index, a.iteration, (uint32_t)gidx, (uint32_t)j);
    return z * a.std0[ad] + a.bias[ad];
  };
  double prev = draw(a.nactions - 1);
  for (int s = 0; s < a.nactions; ++s) {
    const double cur = a.beta0 * draw(s) + a.beta1 * prev;
    prev = cur;
    const int d = s * a.adim + ad;
    const double x = a.mean[d] + cur;
    a.out_nr[(long long)i * a.D + d] = x;
    const long long o = ((long long)i * a.nactions + s) * a.adim_out + ad;
    if (a.out_actions) a.out_actions[o] = (float)x;
    if (a.out_actions64) a.out_actions64[o] = x;
    if (ad == 0)
      for (int e = a.adim; e < a.adim_out; ++e) {
        if (a.out_actions) a.out_actions[o + e] = (float)a.append[e - a.adim];
        if (a.out_actions64) a.out_actions64[o + e] = a.append[e - a.adim];
      }
  }
}

// single block: softmax-weighted elite mean (correlated_noise.py:56-60), elites summed in rank order
__global__ void k_refit_correlated(const double* __restrict__ x, const double* __restrict__ scores, const int* __restrict__ idx, int K,
                                   int D, double kappa, double* mean) {
  extern __shared__ double sw[];   // [K] weights
  __shared__ double s_norm;
  if (threadIdx.x == 0) {
    double rmax = -INFINITY;
    for (int k = 0; k < K; ++k) rmax = fmax(rmax, -scores[idx[k]]);
    double tot = 0.0;
    for (int k = 0; k < K; ++k) { sw[k] = exp(kappa * (-scores[idx[k]] - rmax)); tot += sw[k]; }
    s_norm = tot + 1e-4;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc += x[(long long)k * D + d] * sw[k];
    mean[d] = acc / s_norm;
  }
}

// strict total order: (score asc, NaN last, index asc) == np.argsort(kind='stable')
__device__ __forceinline__ bool key_less(double a, int ia, double b, int ib) {
  const bool pa = ia == 0x7fffffff, pb = ib == 0x7fffffff;      // padding entries sort after everything, NaN scores included
  if (pa != pb) return pb;
  const bool na = a != a, nb = b != b;
  if (na != nb) return nb;
  if (!na && a != b) return a < b;
  return ia < ib;
}

// single block bitonic sort over a power-of-two padded (key, index) array held in global/L2
__global__ void __launch_bounds__(1024) k_topk(const double* __restrict__ scores, int n, int npad, int k, int* out_idx,
                                               double* keys, int* idx) {
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    keys[i] = i < n ? scores[i] : __longlong_as_double(0x7ff0000000000000LL);   // +inf pad
    idx[i] = i < n ? i : 0x7fffffff;
  }
  __syncthreads();
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < npad / 2; t += blockDim.x) {
        const int lo = (t / stride) * 2 * stride + (t % stride);
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double ka = keys[lo], kb = keys[hi];
        const int ia = idx[lo], ib = idx[hi];
        const bool swap = up ? key_less(kb, ib, ka, ia) : key_less(ka, ia, kb, ib);
        if (swap) { keys[lo] = kb; keys[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) out_idx[i] = idx[i];
}

// single block.  mean, unbiased covariance (np.cov(rowvar=False, bias=False)), factor = Xc^T / sqrt(K-1)
__global__ void k_refit(const double* __restrict__ x, int K, int D, double* mean, double* factor, double* cov) {
  extern __shared__ double sm[];   // [D]
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += x[(long long)k * D + d];
    sm[d] = s / (double)K;
    mean[d] = sm[d];
  }
  __syncthreads();
  const double denom = (double)(K > 1 ? K - 1 : 1);
  const double isq = 1.0 / sqrt(denom);
  for (int i = threadIdx.x; i < D * K; i += blockDim.x) {
    const int d = i / K, k = i % K;
    factor[i] = (x[(long long)k * D + d] - sm[d]) * isq;
  }
  if (cov)
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
      const int d = i / D, e = i % D;
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += (x[(long long)k * D + d] - sm[d]) * (x[(long long)k * D + e] - sm[e]);
      cov[i] = s / denom;
    }
}

// One block.  Peer stores travel over NVLink / NVSwitch (or stay in local HBM when the "peer" is another handle on the same
// device); the release/acquire pair on the arrival counters orders them against the peers' reads.
__global__ void __launch_bounds__(256) k_score_exchange(ExchangeArgs a) {
  const double* mine = a.scores[a.rank] + a.row_off + a.offset;
  for (int r = 0; r < a.world; ++r) {
    if (r == a.rank) continue;
    double* dst = a.scores[r] + a.row_off + a.offset;
    for (int i = threadIdx.x; i < a.local; i += blockDim.x) dst[i] = mine[i];
  }
  __threadfence_system();
  __syncthreads();
  const int r = threadIdx.x;
  if (r < a.world && r != a.rank) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flags[r] + a.rank), "r"(a.epoch) : "memory");
    const unsigned* f = a.flags[a.rank] + r;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - a.epoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > a.timeout_ns) { *a.status = 1u; break; }          // reported by vf_cem_finish: never hang the GPU
      __nanosleep(200);
    }
  }
  __syncthreads();
  __threadfence_system();
}

}  // namespace

void launch_score_exchange(const ExchangeArgs& a, cudaStream_t s) {
  ++g_launch_counter;
  k_score_exchange<<<1, 256, 0, s>>>(a);
}

void launch_sample_actions(const SampleArgs& a, int n, cudaStream_t s) {
  ++g_launch_counter;
  if (a.kind == 1) {
    const int total = n * a.adim;
    k_sample_correlated<<<(total + 127) / 128, 128, 0, s>>>(a, n);
    return;
  }
  const int total = n * a.D;
  k_sample_actions<<<(total + 127) / 128, 128, 0, s>>>(a, n);
}
void launch_refit_correlated(const double* elites_nr, const double* scores, const int* idx, int K, int D, double kappa, double* mean,
                             cudaStream_t s) {
  ++g_launch_counter;
  k_refit_correlated<<<1, 128, K * sizeof(double), s>>>(elites_nr, scores, idx, K, D, kappa, mean);
}
void launch_sample_latents(float* zs, int n, int steps, int nz, int goff, uint64_t seed, uint32_t plan, uint32_t iter,
                           cudaStream_t st) {
  ++g_launch_counter;
  const int total = n * steps * nz;
  k_sample_latents<<<(total + 255) / 256, 256, 0, st>>>(zs, n, steps, nz, goff, seed, plan, iter);
}
void launch_reduce_futures(const double* s, int M, int K, double lambda, double* out, cudaStream_t st) {
  ++g_launch_counter;
  k_reduce_futures<<<(M + 127) / 128, 128, 0, st>>>(s, M, K, lambda, out);
}
int topk_padded(int n) {
  int p = 2;
  while (p < n) p <<= 1;
  return p;
}
void launch_topk(const double* scores, int n, int k, int* out_idx, double* work_keys, int* work_idx, cudaStream_t s) {
  ++g_launch_counter;
  k_topk<<<1, 1024, 0, s>>>(scores, n, topk_padded(n), k, out_idx, work_keys, work_idx);
}
void launch_refit(const double* elites_nr, int K, int D, double* mean, double* factor, double* cov, cudaStream_t s) {
  ++g_launch_counter;
  k_refit<<<1, 256, D * sizeof(double), s>>>(elites_nr, K, D, mean, factor, cov);
}

}  // namespace vf

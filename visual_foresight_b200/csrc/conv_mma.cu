// conv_mma.cu — tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a.
//
// D[cout, pixel] = sum_{tap, cin} W[cout, (tap, cin)] * X[(tap, cin), pixel]      (SAME zero padding, stride 1)
//
//   * Output channels sit on the MMA M dimension (UMMA_M = 128 TMEM lanes), pixels on N (TMEM columns).  One
//     epilogue thread therefore owns one output channel of every pixel in the tile: per-channel work is
//     thread-local, and a warp writes 32 consecutive channels of one pixel = one 128-byte NHWC line.
//   * No im2col and no per-tap re-fetch.  A tile's input pixels (+halo) are staged ONCE per 32-channel chunk in
//     shared memory as a FLAT zero-padded image (row pitch Wp = W + k - 1) in the UMMA "interleaved" (no-swizzle)
//     K-major layout [k-chunk of 8 ch][pixel][16 B].  In that layout the operand row of pixel p is at byte offset
//     16*p, so filter tap (dy,dx) is the SAME buffer read through a matrix descriptor whose start address is
//     advanced by (dy*Wp + dx)*16 bytes: 25 taps = 25 descriptors, zero data movement.  The k-1 wrap-around
//     columns per image row are computed and discarded (W/Wp efficiency: 89% at 32x32, 80% at 16x16).
//   * fp32-grade arithmetic on fp16 tensor cores: both operands are split x = hi + lo (fp16 each, weights
//     pre-scaled by a power of two so lo stays normal) and three MMAs hi*hi + hi*lo + lo*hi accumulate into the
//     same fp32 TMEM accumulator (relative error ~2^-21 per product; the dropped lo*lo term is ~2^-22).
//   * Weights are pre-packed on the host in exactly the shared-memory operand layout, so a pipeline stage is one
//     contiguous 16 KB cp.async.bulk (TMA 1-D) completing on an mbarrier: no tensor maps.
//   * Warp roles: warp 0 = weight producer (bulk copies), warp 1 = MMA issuer (one elected thread) + TMEM
//     allocator, warps 2-5 = activation stagers (fp32 -> fp16 hi/lo split on the fly) then epilogue
//     (tcgen05.ld -> scale, + border-class bias -> coalesced NHWC stores).  Persistent CTAs, one per SM.
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "conv_mma.cuh"

namespace vf {
namespace {

constexpr int CH = 32;                 // input channels per chunk
constexpr int KC = CH / 8;             // 16-byte k-chunks per chunk
constexpr int MT = 128;                // cout tile (UMMA M)
constexpr int NSTAGE = 4;              // weight pipeline stages
constexpr int A_HALF_BYTES = KC * MT * 16;          // one hi (or lo) weight block: 8 KB
constexpr int STAGE_BYTES = 2 * A_HALF_BYTES;       // hi + lo: 16 KB
constexpr int NTHREADS = 192;
constexpr int NLOAD = 128;             // stager/epilogue threads (warps 2..5)
constexpr int MAX_SEG = 8;

struct Geometry {
  int H, W, k, pad, Wp;
  int Cin, Cout;
  int nchunk, ntap, n_mt;
  int G;            // images per item
  int v_cnt;        // virtual pixels per image per item (multiple of 32)
  int npass;        // passes over the image's virtual pixel range
  int img_pix;      // staged flat pixels per image
  int npix;         // pixels per k-chunk plane (G*img_pix rounded so that npix % 8 == 2)
  int nseg;         // MMA column segments per image
  int seg_n[MAX_SEG];
  int seg_off[MAX_SEG];
  int ngroups, nitems;
  int passes;       // 1 or 3 MMA passes
  float out_scale;  // 2^-scale_log2
};

struct Params {
  Geometry g;
  View src, out;
  const float* sabias;
  const float* bias;
  const __half* w;     // packed [mt][chunk][tap][hi|lo][kc][128][8]
  int B;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && spins > (1u << 26)) __trap();      // a lost arrival must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread (lane) gets 32 consecutive columns of its TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, no swizzle ("interleave"): 8-row core matrices of 16-byte rows.  LBO = byte distance between the two
// 16-byte K chunks of one MMA (K = 16 halves), SBO = byte distance between 8-row groups.  version = 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(MT >> 4) << 24); }

__global__ void __launch_bounds__(NTHREADS, 1) k_conv_mma(const Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const Geometry& g = P.g;
  const int plane_bytes = KC * g.npix * 16;            // one hi (or lo) activation plane of one chunk buffer
  uint8_t* act[2] = {smem, smem + 2 * plane_bytes};    // [buf] -> hi plane, lo plane follows
  uint8_t* wst = smem + 4 * plane_bytes;               // weight stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + NSTAGE * STAGE_BYTES);
  uint64_t *w_full = bars, *w_empty = bars + NSTAGE, *a_full = bars + 2 * NSTAGE, *a_empty = a_full + 2;
  uint64_t *acc_full = a_empty + 2, *acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], NLOAD); mbar_init(&a_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, NLOAD);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int per_mt = g.ngroups * g.npass;

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = g.passes == 3 ? STAGE_BYTES : A_HALF_BYTES;
      for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
        const int mt = item / per_mt;
        const uint8_t* wbase = reinterpret_cast<const uint8_t*>(P.w) + (size_t)mt * g.nchunk * g.ntap * STAGE_BYTES;
        for (int ct = 0; ct < g.nchunk * g.ntap; ++ct) {
          mbar_wait(&w_empty[s], ph ^ 1);
          mbar_expect_tx(&w_full[s], bytes);
          bulk_g2s(wst + s * STAGE_BYTES, wbase + (size_t)ct * STAGE_BYTES, bytes, &w_full[s]);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, aph[2] = {0, 0}, acc_ph = 0;
      const uint32_t lbo_b = (uint32_t)g.npix * 16;
      for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
        mbar_wait(acc_empty, acc_ph ^ 1);
        acc_ph ^= 1;
        tc_fence_after();
        for (int c = 0; c < g.nchunk; ++c) {
          const int buf = c & 1;
          mbar_wait(&a_full[buf], aph[buf]);
          aph[buf] ^= 1;
          tc_fence_after();
          const uint32_t b_hi = smem_u32(act[buf]), b_lo = b_hi + plane_bytes;
          for (int tap = 0; tap < g.ntap; ++tap) {
            mbar_wait(&w_full[s], ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(wst + s * STAGE_BYTES), a_lo = a_hi + A_HALF_BYTES;
            const uint32_t tap_off = (uint32_t)((tap / g.k) * g.Wp + (tap % g.k)) * 16;
            for (int pass = 0; pass < g.passes; ++pass) {
              const uint32_t a0 = pass == 2 ? a_lo : a_hi;
              const uint32_t b0 = pass == 1 ? b_lo : b_hi;
#pragma unroll
              for (int j = 0; j < CH / 16; ++j) {
                const uint64_t da = make_desc(a0 + j * 2 * (MT * 16), MT * 16, 128);
                const uint32_t first = (c | tap | pass | j) == 0 ? 0u : 1u;
                for (int im = 0; im < g.G; ++im) {
                  for (int sg = 0; sg < g.nseg; ++sg) {
                    const uint32_t col = (uint32_t)(im * g.v_cnt + g.seg_off[sg]);
                    const uint32_t baddr = b0 + j * 2 * lbo_b + tap_off + (uint32_t)(im * g.img_pix + g.seg_off[sg]) * 16;
                    tc_mma_f16(tmem_base + col, da, make_desc(baddr, lbo_b, 128), make_idesc(g.seg_n[sg]), first);
                  }
                }
              }
            }
            tc_commit(&w_empty[s]);                 // stage reusable once these MMAs retire
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
          tc_commit(&a_empty[buf]);                 // chunk buffer reusable
        }
        tc_commit(acc_full);                        // accumulators complete -> epilogue
      }
    }
  } else {
    // ===== activation stagers, then epilogue (warps 2..5) =====
    const int t = threadIdx.x - 64;                 // 0..127
    const int q4 = warp & 3;                        // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;                 // output channel within the cout tile
    uint32_t aeph[2] = {0, 0}, accf_ph = 0;
    for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
      const int mt = item / per_mt;
      const int rem = item % per_mt;
      const int grp = rem / g.npass, ps = rem % g.npass;
      const int b0 = grp * g.G;
      const int v_lo = ps * g.v_cnt;
      // ---- stage input chunks ----
      for (int c = 0; c < g.nchunk; ++c) {
        const int buf = c & 1;
        mbar_wait(&a_empty[buf], aeph[buf] ^ 1);
        aeph[buf] ^= 1;
        uint4* hi = reinterpret_cast<uint4*>(act[buf]);
        uint4* lo = reinterpret_cast<uint4*>(act[buf] + plane_bytes);
        const int total = g.G * g.img_pix * KC;
        for (int idx = t; idx < total; idx += NLOAD) {
          const int kc = idx & (KC - 1);
          const int pl = idx >> 2;                   // pixel slot in the plane
          const int im = pl / g.img_pix, ql = pl - im * g.img_pix;
          const int q = v_lo + ql;                   // flat index in the zero-padded image
          const int yy = q / g.Wp - g.pad, xx = q % g.Wp - g.pad;
          const int b = b0 + im;
          uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
          if (b < P.B && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) {
            const float* sp = P.src.p + (long long)b * P.src.sample_stride + (long long)(yy * g.W + xx) * P.src.pix_stride +
                              P.src.ch_off + c * CH + kc * 8;
            const float4 f0 = __ldg(reinterpret_cast<const float4*>(sp));
            const float4 f1 = __ldg(reinterpret_cast<const float4*>(sp) + 1);
            const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            __half h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              h[e] = __float2half_rn(f[e]);
              l[e] = __float2half_rn(f[e] - __half2float(h[e]));
            }
            vh = *reinterpret_cast<uint4*>(h);
            vl = *reinterpret_cast<uint4*>(l);
          }
          hi[kc * g.npix + pl] = vh;
          lo[kc * g.npix + pl] = vl;
        }
        fence_proxy_async();                         // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(&a_full[buf]);
      }
      // ---- epilogue ----
      mbar_wait(acc_full, accf_ph);
      accf_ph ^= 1;
      tc_fence_after();
      const int n = mt * MT + row;
      for (int im = 0; im < g.G; ++im) {
        const int b = b0 + im;
        for (int cc = 0; cc < g.v_cnt; cc += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(im * g.v_cnt + cc), r);
          if (b >= P.B) continue;
          int v = v_lo + cc;
          int oy = v / g.Wp, ox = v - oy * g.Wp;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (ox < g.W && oy < g.H) {
              float val = __uint_as_float(r[j]) * g.out_scale;
              if (P.sabias) {
                const int cls = border_class(oy, g.H, g.pad) * g.k + border_class(ox, g.W, g.pad);
                val += __ldg(P.sabias + ((long long)b * g.ntap + cls) * g.Cout + n);
              } else if (P.bias) {
                val += __ldg(P.bias + n);
              }
              P.out.p[(long long)b * P.out.sample_stride + (long long)(oy * g.W + ox) * P.out.pix_stride + P.out.ch_off + n] = val;
            }
            if (++ox == g.Wp) { ox = 0; ++oy; }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
int g_num_sms = 0;

bool plan_geometry(int k, int Cin, int Cout, int H, int W, int B, int passes, Geometry* out) {
  Geometry g;
  memset(&g, 0, sizeof(g));
  g.H = H; g.W = W; g.k = k; g.pad = k / 2; g.Wp = W + k - 1; g.Cin = Cin; g.Cout = Cout;
  g.nchunk = Cin / CH; g.ntap = k * k; g.n_mt = Cout / MT; g.passes = passes;
  const int V = H * g.Wp;                               // virtual pixels per image (incl. k-1 wrap columns per row)
  // (G, v_cnt, npass): several whole small images per item, or an (almost) even slice of one large image.
  // v_cnt is rounded up to 32 columns; the overshoot reads zero-filled staging rows and is masked in the epilogue.
  if (V <= 256) {
    g.v_cnt = (V + 31) / 32 * 32;
    g.G = std::max(1, std::min(512 / g.v_cnt, 3));
    g.npass = 1;
  } else {
    g.G = 1;
    g.npass = (V + 511) / 512;
    g.v_cnt = ((V + g.npass - 1) / g.npass + 31) / 32 * 32;
  }
  // MMA column segments: n <= 256, multiple of 16
  g.nseg = 0;
  int left = g.v_cnt, off = 0;
  const int nsplit = (g.v_cnt + 255) / 256;
  const int base = ((g.v_cnt / nsplit) + 15) / 16 * 16;
  while (left > 0) {
    const int n = std::min(base, left);
    if (n % 16 || g.nseg >= MAX_SEG) return false;
    g.seg_n[g.nseg] = n; g.seg_off[g.nseg] = off; ++g.nseg;
    off += n; left -= n;
  }
  g.img_pix = g.v_cnt + (k - 1) * g.Wp + (k - 1);
  g.img_pix = (g.img_pix + 7) / 8 * 8;
  for (;; --g.G) {                                       // shrink the image group until the staging buffers fit
    int npix = g.G * g.img_pix;
    while (npix % 8 != 2) ++npix;                        // conflict-free 16-byte stores across k-chunks
    g.npix = npix;
    if (g.G == 1 || (size_t)4 * KC * npix * 16 + (size_t)NSTAGE * STAGE_BYTES + 256 <= (size_t)227 * 1024) break;
  }
  g.ngroups = (B + g.G - 1) / g.G;
  g.nitems = g.n_mt * g.ngroups * g.npass;
  *out = g;
  return true;
}

// >= 116 KB so that exactly one CTA is resident per SM: every CTA allocates all 512 TMEM columns
size_t smem_bytes(const Geometry& g) {
  return std::max((size_t)4 * KC * g.npix * 16 + (size_t)NSTAGE * STAGE_BYTES + 256, (size_t)116 * 1024);
}

}  // namespace

bool mma_conv_supported(int k, int cin, int cout, int H, int W) {
  if ((k != 3 && k != 5) || cin % CH || cout % MT) return false;
  Geometry g;
  if (!plan_geometry(k, cin, cout, H, W, 1, 3, &g)) return false;
  return smem_bytes(g) <= 227 * 1024;
}

int mma_conv_prepare_weights(const float* w_sp, int k, int cin, int cout, MmaConvWeights* out, std::vector<void*>* allocs,
                             std::string* err) {
  if (cin % CH || cout % MT) { if (err) *err = "cin % 32 or cout % 128"; return -1; }
  const int kk = k * k, nchunk = cin / CH, n_mt = cout / MT;
  float amax = 0.f;
  for (size_t i = 0; i < (size_t)kk * cin * cout; ++i) amax = std::max(amax, fabsf(w_sp[i]));
  int sl = 0;
  if (amax > 0.f) sl = (int)floorf(log2f(16384.0f / amax));           // max |w| * 2^sl in [8192, 16384]
  sl = std::max(-24, std::min(sl, 24));
  const float scale = ldexpf(1.0f, sl);
  const size_t total = (size_t)n_mt * nchunk * kk * 2 * KC * MT * 8;
  std::vector<__half> packed(total);
  for (int mt = 0; mt < n_mt; ++mt)
    for (int c = 0; c < nchunk; ++c)
      for (int t = 0; t < kk; ++t) {
        const size_t blk = ((((size_t)mt * nchunk + c) * kk + t) * 2) * KC * MT * 8;
        for (int kc = 0; kc < KC; ++kc)
          for (int r = 0; r < MT; ++r)
            for (int e = 0; e < 8; ++e) {
              const int ci = c * CH + kc * 8 + e, n = mt * MT + r;
              const float v = w_sp[((size_t)t * cin + ci) * cout + n] * scale;
              const __half h = __float2half_rn(v);
              const __half l = __float2half_rn(v - __half2float(h));
              packed[blk + ((size_t)kc * MT + r) * 8 + e] = h;
              packed[blk + (size_t)KC * MT * 8 + ((size_t)kc * MT + r) * 8 + e] = l;
            }
      }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, total * sizeof(__half));
  if (e != cudaSuccess) { if (err) *err = cudaGetErrorString(e); return -1; }
  allocs->push_back(d);
  e = cudaMemcpy(d, packed.data(), total * sizeof(__half), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { if (err) *err = cudaGetErrorString(e); return -1; }
  out->w_hi = reinterpret_cast<__half*>(d);
  out->w_lo = nullptr;
  out->k = k; out->cin = cin; out->cout = cout; out->scale_log2 = sl; out->ready = true;
  return 0;
}

int mma_conv_launch(const MmaConvWeights& w, const MmaConvCall& c, int B, cudaStream_t s) {
  Params P;
  if (!plan_geometry(w.k, w.cin, w.cout, c.H, c.W, B, c.passes == 3 ? 3 : 1, &P.g)) return -1;
  P.g.out_scale = ldexpf(1.0f, -w.scale_log2);
  P.src = c.src; P.out = c.out; P.sabias = c.sabias; P.bias = c.bias; P.w = w.w_hi; P.B = B;
  if ((c.src.pix_stride % 4) || (c.src.ch_off % 4) || (c.src.sample_stride % 4)) return -2;      // float4 loads
  const size_t smem = smem_bytes(P.g);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_conv_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -3;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const int grid = std::min(P.g.nitems, g_num_sms > 0 ? g_num_sms : 148);
  ++g_launch_counter;
  k_conv_mma<<<grid, NTHREADS, smem, s>>>(P);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace vf

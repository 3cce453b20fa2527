// conv_mma.cu — tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a, operands fed by TMA.
//
// D[cout, pixel] = sum_{tap, cin} W[cout, (tap, cin)] * X[(tap, cin), pixel]      (SAME zero padding, stride 1)
//
//   * Activations live in HBM in "split-half" storage (vf_common.cuh): two fp16 planes hi = rn16(x), lo = rn16(x - hi),
//     written once by the producing kernel.  The conv fetches a tile's input rows (+halo) per 32-channel chunk with ONE
//     cp.async.bulk.tensor (5-D tensor map: channel, x, y, sample, plane) per plane, box = 32 ch x (W+k-1) x R rows starting
//     at x = -pad: the hardware zero-fills the SAME padding and writes the SWIZZLE_64B operand layout directly.  No
//     register staging, no conversion pass, no im2col.
//   * The box in shared memory is a FLAT zero-padded image (row pitch Wp = W + k - 1), one operand row per pixel.  Filter
//     tap (dy,dx) is the SAME buffer read through a matrix descriptor whose start address is advanced by (dy*Wp + dx)
//     pixel rows: 25 taps = 25 descriptors, zero data movement.  In the generic tiling the k-1 wrap-around columns per
//     image row are computed and discarded (W/Wp efficiency: 89% at 32x32, 80% at 16x16); the "row group" modes
//     (Geometry::rg) avoid them for maps whose width is a multiple of 8: the descriptor's 8-row group pitch is set to the
//     padded row, so an MMA's N rows are 8-pixel groups of consecutive IMAGE rows — stacked 8-wide images (rg 1), the two
//     column groups of a 16-wide image (rg 2), or one 8-pixel column STRIP of a wider image over all its rows (rg 3, the
//     32x32 gate convs: N = 256 per MMA, staged box = the 12-pixel-wide strip).
//   * Two orientations.  Cout >= 128: output channels on the MMA M dimension (TMEM lanes), pixels on N; a warp stores 32
//     consecutive channels of one pixel = one 128-byte NHWC line.  Cout <= 64 ("swap"): pixels on M in 128-row units,
//     channels on N; one epilogue thread owns one pixel and writes its channels with 16-byte stores.
//   * fp32-grade arithmetic on fp16 tensor cores: three MMAs hi*hi + hi*lo + lo*hi accumulate into the same fp32 TMEM
//     accumulator (weights pre-scaled by a power of two so their lo part stays normal).
//   * Weights are pre-packed on the host in exactly the shared-memory operand layout, so a pipeline stage is one
//     contiguous cp.async.bulk (TMA 1-D) completing on an mbarrier.
//   * Warp roles (384 threads, persistent CTAs, one per SM): warp 0 = weight producer, warp 1 = MMA issuer (one elected
//     thread) + TMEM allocator, warp 2 = activation producer (tensor TMA), warps 4-11 = epilogue (tcgen05.ld -> scale,
//     + border-class bias, optional sigmoid / instance-norm partial sums -> NHWC stores).  Two TMEM accumulator sets and two
//     activation buffers: the epilogue of item i and the loads of item i+1 overlap the MMAs.
//
//   * k_conv_pair (further down): the 16x16 layers with whole 256-channel tiles run on CTA PAIRS (tcgen05 cta_group::2, M = 256):
//     each SM of a TPC stages half of both operands, one MMA per K step spans both.
//   * Row-stacked thin layers with np <= 16 and >= 2 channel chunks (the mask-logit conv) multiply X_hi by BOTH weight halves
//     in one MMA of N = 2*k*np (Geometry::stk): 2 MMAs per K step instead of 3, the W_lo products are added in the epilogue.
//
// Operand layouts (VF_MMA_LAYOUT): 1 = SWIZZLE_64B, 32-channel chunks, pixel rows of 64 B (default);
//                                  2 = SWIZZLE_128B, 64-channel chunks, pixel rows of 128 B.
// The swizzle XOR acts on absolute shared-memory address bits for TMA and for the MMA descriptors alike (measured:
// tap-shifted descriptors are exact with base_offset = 0), so arbitrary 64-byte row shifts stay consistent.
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "conv_mma.cuh"

namespace vf {
namespace {

constexpr int MT = 128;                // cout tile (UMMA M)
constexpr int NTHREADS = 384;
constexpr int MAX_SEG = 8;
constexpr int MAX_STAGE = 8;
constexpr int MAX_UNIT = 4;            // (images x column segments) per item
constexpr int MAX_UNIT_B = 8;          // swapped orientation: 128-pixel units per item

struct Geometry {
  int layout, bo_mode;
  int ch, kc, ksteps, row_bytes, swz_mask, half_bytes, stage_bytes, nstage, nbuf, plane_bytes;
  int H, W, k, pad, Wp;   // k = filter rows, pad = k/2
  int kw, padx;           // filter columns actually multiplied (== k, or 1 when the producer folded the dx taps into channels)
  int kcl, padc;          // kernel size of the border classes of the tiled action/state bias (the layer's nominal k)
  int Cin, Cout;
  int nchunk, ntap, n_mt;
  int G;            // images per item
  int v_cnt;        // virtual pixels per image per item (multiple of 32)
  int npass;        // passes over the image's virtual pixel range
  int img_pix;      // shared-memory pixel rows reserved per image (>= R*Wp, multiple of 8)
  int R;            // rows of the TMA box (covers v_cnt + halo for every pass offset)
  int box_bytes;    // bytes one TMA box delivers (R * Wp * row_bytes)
  int nchunk0;      // channel chunks taken from the first source (the rest come from the second)
  int nseg;         // MMA column segments per image
  int seg_n[MAX_SEG];
  int seg_off[MAX_SEG];   // first accumulator column of the segment
  int seg_px[MAX_SEG];    // first shared-memory pixel row of the segment (== seg_off except in row-group mode 2)
  int ngroups, nitems;
  int passes;       // 1 or 3 MMA passes
  int nacc;         // TMEM accumulator sets (2 = epilogue overlaps the next item's MMAs)
  int swap;         // thin layers (Cout <= 64): pixels on the MMA M dimension (TMEM lanes), output channels on N.
                    //   2 = row-stacked: the k taps of one filter ROW sit side by side on N (N = k*np), so one MMA serves k
                    //       taps; the dx shifts are undone in the epilogue (lane shuffles).  1 = one MMA per tap (k*np > 256).
  int np;           // swap: output channels padded to a multiple of 16
  int units;        // swap: 128-pixel units per item
  int ncols;        // swap: weight rows per stage half (np, or k*np when row-stacked) = MMA N of one pass
  int stk;          // row-stacked, 3 passes: 1 = the hi and lo weight halves of a stage (contiguous rows) are ONE B operand of
                    //   N = 2*ncols, so X_hi * [W_hi | W_lo] is one MMA (2 MMAs per K step instead of 3; the W_lo products
                    //   land in their own ncols TMEM columns and are added in the epilogue).  A unit then owns 2*ncols columns.
  int ustride;      // swap: output pixels per unit (128, or 128-(k-1) when row-stacked: units overlap by the dx halo)
  int nst;          // weight stages per channel chunk (k*k taps, or k filter rows when row-stacked)
  int ksteps_last;  // K=16 steps of the last channel chunk (its zero-padded tail is not multiplied)
  int rg;           // wide path, W == 8 ("row groups"): an image row is exactly one 8-pixel core-matrix group of the pixel operand, so
                    //   the descriptor's group stride (SBO) is the PADDED row pitch Wp and the k-1 wrap columns are never multiplied;
                    //   the G images of an item are stacked pad rows apart (the zero halo rows between two images are shared) and
                    //   ONE MMA of N = ((G-1)(H+pad)+H)*8 columns spans all of them.
                    //   rg == 2 (W == 16): one image per item, two MMAs per K step — the left and the right 8-pixel group of every
                    //   row (N = 8*H each, group stride Wp, the right one starting 8 pixels later) -> columns [0, 8H) and [8H, 16H).
                    //   rg == 3 (W == 32, or any W % 8 == 0 wider than 16 with 8*H <= 256): an item is ONE 8-pixel column group of
                    //   an image over all rows (npass = W/8 items per image): the staged box is the 8+k-1 pixel wide strip
                    //   (Wp = strip width), one MMA of N = 8*H per K step, no wrap columns and no row-block halo re-reads.
  int tsplit;       // rg == 3: 1 = the strips of the LAST, partial round of the persistent grid are split in an upper and a lower
  int n_full;       //   half (N = 4*H each): work items [0, n_full) are whole strips, the rest half strips (strip_of())
  int half_cols;    // rg == 2: accumulator columns per 8-pixel column group (8*H)
  int col_stride;   // rg: TMEM columns between consecutive images of an item ((H+pad)*8)
  int ncols_item;   // rg: TMEM columns of one accumulator set (the MMA's N)
  int We;           // row pitch of the epilogue's column -> pixel map (Wp, or W when rg)
  float out_scale;  // 2^-scale_log2
  // division by kernel-invariant divisors as one multiply-high: x / d == __umulhi(x, ceil(2^32 / d)) for x * d < 2^32 (d > 1)
  uint32_t m_Wp, m_npass, m_np;
  int dual;         // host switch of the two-issuer mode (row-stacked path)
};

__host__ __device__ inline uint32_t fd_magic(int d) { return d > 1 ? (uint32_t)(((1ull << 32) + (uint32_t)d - 1) / (uint32_t)d) : 0u; }

struct Params {
  alignas(64) CUtensorMap tmap0;   // activations of source 0: (channel, x, y, sample, plane)
  alignas(64) CUtensorMap tmap1;   // source 1 (channel chunks >= nchunk0) or a copy of tmap0
  Geometry g;
  View out;
  const float* sabias;
  const float* bias;
  const __half* w;     // packed [mt][chunk][tap][hi|lo][operand tile]
  int B;
  int act;
  double* stats_partial;   // optional: per-(sample, channel) sum / sum-of-squares partials of the stored outputs
  int stats_S;             // partial slots per (sample, channel)
  // optional "last arriver finalises": the warp that completes a (sample, 32- or 16-channel group) — counted with one
  // atomic per warp per item in stats_cnt[b * VF_STAT_CNT_STRIDE + group] — turns the S partials into (mean, rstd) pairs in
  // stats_fin[(b * Cout + c) * 2] and resets the counter: no separate finalize launch, consumers read two floats per channel
  float* stats_fin;
  int* stats_cnt;
  int stats_npix;
  float stats_eps;
};

// (mean, rstd) of one plane from its S partial slots, combined in slot order (== k_stats_finalize); the slots were written by
// other SMs: read them from L2 (ld.cg)
__device__ __forceinline__ void stats_finalize_plane(const double* part, int S, int npix, float eps, float* fin) {
  double ts = 0.0, tq = 0.0;
  for (int k = 0; k < S; ++k) { ts += __ldcg(part + 2 * k); tq += __ldcg(part + 2 * k + 1); }
  const double mean = ts / npix;
  double var = tq / npix - mean * mean;
  if (var < 0.0) var = 0.0;
  fin[0] = (float)mean;
  fin[1] = (float)(1.0 / sqrt(var + (double)eps));
}

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
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try(addr, parity)) return;             // fast path: already complete
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
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
// Warp-uniform issue: EVERY lane executes the surrounding address arithmetic (so the compiler can keep the warp-uniform
// descriptors in uniform registers instead of moving them there per MMA), the election happens inside the instruction
// sequence and only the elected lane's tcgen05 instruction takes effect.  elect.sync picks the same lane every time for the
// same member mask, so the commits below track the MMAs above.
__device__ __forceinline__ void tc_mma_f16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// explicit shared-space accesses on 32-bit shared addresses (generic pointers cost a cvta sequence per access)
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Shared-memory matrix descriptor (K-major).  version = 1 (sm_100).  layout_type: 0 none, 4 = 64B, 2 = 128B swizzle.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type,
                                              uint32_t base_offset) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) |
         ((uint64_t)(layout_type & 7) << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(MT >> 4) << 24); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ int fdiv(int x, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)x, magic) : x; }

// All MMAs of one filter tap (one weight stage): PASSES x KSTEPS x U instructions, straight-line, issued by the single
// elected thread.  Descriptors only differ in their low word (start address >> 4), advanced by pre-shifted offsets.
template <int PASSES, int KSTEPS, int U>
__device__ __forceinline__ void issue_tap(uint64_t da_st, uint64_t db_tap, uint64_t a_half, uint64_t b_plane, uint64_t kstep_a,
                                          uint64_t kstep_b, const uint32_t (&ucol)[MAX_UNIT], const uint64_t (&uoff)[MAX_UNIT],
                                          const uint32_t (&uidesc)[MAX_UNIT], uint32_t acc_first) {
#pragma unroll
  for (int pass = 0; pass < PASSES; ++pass) {
    const uint64_t da_p = da_st + (pass == 2 ? a_half : 0);
    const uint64_t db_p = db_tap + (pass == 1 ? b_plane : 0);
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
      const uint64_t da = da_p + j * kstep_a, db = db_p + j * kstep_b;
#pragma unroll
      for (int u = 0; u < U; ++u) tc_mma_f16(ucol[u], da, db + uoff[u], uidesc[u], (pass | j) == 0 ? acc_first : 1u);
    }
  }
}

// Swapped orientation: A = activation rows (128 pixels per unit, tap-shifted), B = the weight tile (np output channels).
// The thin layers are ISSUE-bound (ncu source view of masks1: 144 MMAs of N = 48 per item, ~14 SASS instructions each on one
// warp): descriptors are therefore advanced in their LOW word only (start address >> 4 and LBO live there; every offset
// added is < 2^14, so nothing carries into the constant high word), the K-step count is a template parameter (straight-line
// code), and the accumulator column / pixel-row offset of every unit come from small precomputed tables.
__device__ __forceinline__ void tc_mma_f16_elect32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// STK (3 passes only): pass 0 multiplies X_hi by BOTH weight halves (idesc2: N = 2*ncols), pass 1 X_lo by W_hi; no pass 2.
template <int PASSES, int KSTEPS, int U, int STK = 0>
__device__ __forceinline__ void issue_tap_swap(uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t a_plane, uint32_t b_half,
                                               uint32_t kstep, const uint32_t (&uoff)[MAX_UNIT_B], const uint32_t (&ucol)[MAX_UNIT_B],
                                               uint32_t idesc, uint32_t acc_first, uint32_t idesc2 = 0) {
  if constexpr (STK == 1) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const uint32_t a_p = a_lo + (pass == 1 ? a_plane : 0u);
#pragma unroll
      for (int j = 0; j < KSTEPS; ++j) {
        const uint32_t a = a_p + j * kstep, b = b_lo + j * kstep;
#pragma unroll
        for (int u = 0; u < U; ++u)
          tc_mma_f16_elect32(ucol[u], a + uoff[u], a_hi, b, b_hi, pass == 0 ? idesc2 : idesc, (pass | j) == 0 ? acc_first : 1u);
      }
    }
    return;
  }
#pragma unroll
  for (int pass = 0; pass < PASSES; ++pass) {
    const uint32_t a_p = a_lo + (pass == 1 ? a_plane : 0u);
    const uint32_t b_p = b_lo + (pass == 2 ? b_half : 0u);
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
      const uint32_t a = a_p + j * kstep, b = b_p + j * kstep;
#pragma unroll
      for (int u = 0; u < U; ++u)      // unit u: pixel rows [u*ustride, u*ustride + 128) of the item -> its own TMEM columns
        tc_mma_f16_elect32(ucol[u], a + uoff[u], a_hi, b, b_hi, idesc, (pass | j) == 0 ? acc_first : 1u);
    }
  }
}

struct IssueCtx {
  uint64_t da_zero, db_zero, a_half, b_plane, kstep_a, kstep_b;
  uint32_t act_base, wst_base, tmem_base, pix_b;
  uint64_t* w_full; uint64_t* w_empty; uint64_t* a_full; uint64_t* a_empty; uint64_t* acc_full; uint64_t* acc_empty;
};

// work item -> strip (the item index of the unsplit schedule) and part: -1 = whole item, 0 / 1 = upper / lower half of a strip of
// the split last round (Geometry::tsplit)
__device__ __forceinline__ int strip_of(const Geometry& g, int item, int& part) {
  part = -1;
  if (!g.tsplit || item < g.n_full) return item;
  const int t = item - g.n_full;
  part = t & 1;
  return g.n_full + (t >> 1);
}

// Whole-kernel MMA issue loop, specialised at compile time on (passes, k-steps, accumulator units): inside the tap loop
// there is one mbarrier wait, one election, PASSES*KSTEPS*U back-to-back UTCHMMA and one commit.
template <int PASSES, int KSTEPS, int U>
__device__ __forceinline__ void issuer_loop(const Geometry& g, const IssueCtx& cx, int acc_cols) {
  uint32_t ucol0[MAX_UNIT], ucol[MAX_UNIT], uidesc[MAX_UNIT];
  uint64_t uoff[MAX_UNIT];
#pragma unroll
  for (int u = 0; u < MAX_UNIT; ++u) {
    const int im = u / g.nseg, sg = u % g.nseg;
    ucol0[u] = cx.tmem_base + (uint32_t)(im * g.v_cnt + g.seg_off[sg]);
    uoff[u] = (uint64_t)(((uint32_t)(im * g.img_pix + g.seg_px[sg]) * cx.pix_b) >> 4);
    uidesc[u] = make_idesc(g.seg_n[sg]);
  }
  const uint32_t idesc_half = make_idesc(g.v_cnt >> 1);                  // half strips of the split last round (rg 3: one unit)
  const uint64_t half_off16 = (uint64_t)(((uint32_t)((g.H >> 1) * g.Wp) * cx.pix_b) >> 4);
  const uint32_t idesc_full0 = uidesc[0];
  const int nitems = g.nitems, nchunk = g.nchunk, ntap = g.nst, kk = g.kw, nbuf = g.nbuf, nstage = g.nstage, nacc = g.nacc;
  const uint32_t stage16 = (uint32_t)(g.stage_bytes >> 4), buf16 = (uint32_t)((2 * g.plane_bytes) >> 4);
  const uint64_t pix16 = (uint64_t)(cx.pix_b >> 4), rowskip16 = (uint64_t)(((uint32_t)(g.Wp - g.kw) * cx.pix_b) >> 4);
  const uint64_t da_base = cx.da_zero + (uint64_t)(cx.wst_base >> 4), db_base = cx.db_zero + (uint64_t)(cx.act_base >> 4);
  int s = 0;
  uint32_t ph = 0, job = 0, it = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
    const int a = it % nacc;
    // the box starts at the image row containing the item's first virtual pixel: skip (v_lo mod Wp) pixel rows
    int part;
    const int strip = strip_of(g, item, part);
    const uint64_t item_off16 = g.rg == 3 ? (part > 0 ? half_off16 : 0ull)
                                          : (uint64_t)(((uint32_t)(((strip % g.npass) * g.v_cnt) % g.Wp) * cx.pix_b) >> 4);
    uidesc[0] = part >= 0 ? idesc_half : idesc_full0;
    mbar_wait(&cx.acc_empty[a], ((it / nacc) & 1) ^ 1);
    tc_fence_after();
#pragma unroll
    for (int u = 0; u < MAX_UNIT; ++u) ucol[u] = ucol0[u] + (uint32_t)(a * acc_cols);
    uint32_t acc = 0;                                                    // first MMA of every unit overwrites
    for (int c = 0; c < nchunk; ++c, ++job) {
      const int buf = job % nbuf;
      mbar_wait(&cx.a_full[buf], (job / nbuf) & 1);
      tc_fence_after();
      uint64_t db_tap = db_base + (uint64_t)(buf * buf16) + item_off16;
      int tx = 0;
      for (int tap = 0; tap < ntap; ++tap) {
        mbar_wait(&cx.w_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          issue_tap<PASSES, KSTEPS, U>(da_base + (uint64_t)(s * stage16), db_tap, cx.a_half, cx.b_plane, cx.kstep_a, cx.kstep_b,
                                       ucol, uoff, uidesc, acc);
          tc_commit(&cx.w_empty[s]);                                     // stage reusable once these MMAs retire
        }
        acc = 1;
        if (++s == nstage) { s = 0; ph ^= 1; }
        db_tap += pix16;
        if (++tx == kk) { tx = 0; db_tap += rowskip16; }
      }
      if (elect_one()) tc_commit(&cx.a_empty[buf]);                      // chunk buffer reusable
    }
    if (elect_one()) tc_commit(&cx.acc_full[a]);                         // accumulators complete -> epilogue
  }
}

template <int PASSES, int U>
__device__ __forceinline__ void issuer_loop_swap(const Geometry& g, const IssueCtx& cx, int acc_cols) {
  const uint32_t unit_step = (128u * cx.pix_b) >> 4;
  const uint32_t idesc_b = make_idesc(g.np), np = (uint32_t)g.np;
  const int nitems = g.nitems, nchunk = g.nchunk, ntap = g.nst, kk = g.kw, nbuf = g.nbuf, nstage = g.nstage, nacc = g.nacc;
  const int ksteps = g.ksteps;
  const uint32_t stage16 = (uint32_t)(g.stage_bytes >> 4), buf16 = (uint32_t)((2 * g.plane_bytes) >> 4);
  const uint32_t pix16 = cx.pix_b >> 4, rowskip16 = ((uint32_t)(g.Wp - g.kw) * cx.pix_b) >> 4;
  const uint32_t w_hi = (uint32_t)(cx.da_zero >> 32), x_hi = (uint32_t)(cx.db_zero >> 32);
  const uint32_t dw_base = (uint32_t)cx.da_zero + (cx.wst_base >> 4), dx_base = (uint32_t)cx.db_zero + (cx.act_base >> 4);
  const uint32_t a_plane = (uint32_t)cx.b_plane, b_half = (uint32_t)cx.a_half, kstep = (uint32_t)cx.kstep_b;
  uint32_t uoff[MAX_UNIT_B], ucol0[MAX_UNIT_B], ucol[MAX_UNIT_B];
#pragma unroll
  for (int u = 0; u < MAX_UNIT_B; ++u) { uoff[u] = (uint32_t)u * unit_step; ucol0[u] = cx.tmem_base + (uint32_t)u * np; }
  int s = 0;
  uint32_t ph = 0, job = 0, it = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
    const int a = nacc == 2 ? (int)(it & 1u) : 0;
    mbar_wait(&cx.acc_empty[a], (((nacc == 2 ? it >> 1 : it)) & 1) ^ 1);
    tc_fence_after();
#pragma unroll
    for (int u = 0; u < MAX_UNIT_B; ++u) ucol[u] = ucol0[u] + (uint32_t)(a * acc_cols);
    const uint32_t item_off16 = ((uint32_t)(((item % g.npass) * g.v_cnt) % g.Wp) * cx.pix_b) >> 4;
    uint32_t acc = 0;
    for (int c = 0; c < nchunk; ++c, ++job) {
      const int buf = job % nbuf;
      mbar_wait(&cx.a_full[buf], (job / nbuf) & 1);
      tc_fence_after();
      uint32_t da_tap = dx_base + (uint32_t)(buf * buf16) + item_off16;
      int tx = 0;
      for (int tap = 0; tap < ntap; ++tap) {
        mbar_wait(&cx.w_full[s], ph);
        tc_fence_after();
        if (ksteps == 2) issue_tap_swap<PASSES, 2, U>(da_tap, x_hi, dw_base + (uint32_t)(s * stage16), w_hi, a_plane, b_half, kstep, uoff, ucol, idesc_b, acc);
        else if (ksteps == 1) issue_tap_swap<PASSES, 1, U>(da_tap, x_hi, dw_base + (uint32_t)(s * stage16), w_hi, a_plane, b_half, kstep, uoff, ucol, idesc_b, acc);
        else issue_tap_swap<PASSES, 4, U>(da_tap, x_hi, dw_base + (uint32_t)(s * stage16), w_hi, a_plane, b_half, kstep, uoff, ucol, idesc_b, acc);
        tc_commit_elect(&cx.w_empty[s]);
        acc = 1;
        if (++s == nstage) { s = 0; ph ^= 1; }
        da_tap += pix16;
        if (++tx == kk) { tx = 0; da_tap += rowskip16; }
      }
      tc_commit_elect(&cx.a_empty[buf]);
    }
    tc_commit_elect(&cx.acc_full[a]);
  }
}

// Row-stacked thin layers: one weight stage = the k taps of filter row dy side by side on N (ncols = k*np); the activation
// descriptor advances by one padded image row (Wp pixel rows) per stage.  k MMAs-rows instead of k*k taps.
// u0: first unit this warp issues (the row-stacked thin layers split the units of an item over TWO issuer warps, 1 and 3:
// their MMAs go to disjoint accumulator columns, and these layers are bound by the ~10 SASS instructions it takes one warp
// to get an MMA of N = 48..96 out of the door)
template <int PASSES, int U, int STK = 0>
__device__ __forceinline__ void issuer_loop_swap2(const Geometry& g, const IssueCtx& cx, int acc_cols, int u0) {
  const uint32_t unit_step = ((uint32_t)g.ustride * cx.pix_b) >> 4;
  const uint32_t idesc_b = make_idesc(g.ncols), idesc_2 = make_idesc(2 * g.ncols), ncols = (uint32_t)(g.ncols * (1 + STK));
  const int nitems = g.nitems, nchunk = g.nchunk, kk = g.k, nbuf = g.nbuf, nstage = g.nstage, nacc = g.nacc;
  const uint32_t stage16 = (uint32_t)(g.stage_bytes >> 4), buf16 = (uint32_t)((2 * g.plane_bytes) >> 4);
  const uint32_t row16 = ((uint32_t)g.Wp * cx.pix_b) >> 4;
  const uint32_t w_hi = (uint32_t)(cx.da_zero >> 32), x_hi = (uint32_t)(cx.db_zero >> 32);
  const uint32_t dw_base = (uint32_t)cx.da_zero + (cx.wst_base >> 4), dx_base = (uint32_t)cx.db_zero + (cx.act_base >> 4);
  const uint32_t a_plane = (uint32_t)cx.b_plane, b_half = (uint32_t)cx.a_half, kstep = (uint32_t)cx.kstep_b;
  uint32_t uoff[MAX_UNIT_B], ucol0[MAX_UNIT_B], ucol[MAX_UNIT_B];
#pragma unroll
  for (int u = 0; u < MAX_UNIT_B; ++u) { uoff[u] = (uint32_t)(u0 + u) * unit_step; ucol0[u] = cx.tmem_base + (uint32_t)(u0 + u) * ncols; }
  int s = 0;
  uint32_t ph = 0, job = 0, it = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
    const int a = nacc == 2 ? (int)(it & 1u) : 0;
    mbar_wait(&cx.acc_empty[a], (((nacc == 2 ? it >> 1 : it)) & 1) ^ 1);
    tc_fence_after();
#pragma unroll
    for (int u = 0; u < MAX_UNIT_B; ++u) ucol[u] = ucol0[u] + (uint32_t)(a * acc_cols);
    const int vlo = (item - fdiv(item, g.m_npass) * g.npass) * g.v_cnt;             // first virtual pixel of the item's pass
    const uint32_t item_off16 = ((uint32_t)(vlo - fdiv(vlo, g.m_Wp) * g.Wp) * cx.pix_b) >> 4;
    uint32_t acc = 0;
    for (int c = 0; c < nchunk; ++c, ++job) {
      const int buf = job % nbuf;
      const int ksteps = c == nchunk - 1 ? g.ksteps_last : g.ksteps;
      mbar_wait(&cx.a_full[buf], (job / nbuf) & 1);
      tc_fence_after();
      uint32_t da_row = dx_base + (uint32_t)(buf * buf16) + item_off16;
      for (int dy = 0; dy < kk; ++dy) {
        mbar_wait(&cx.w_full[s], ph);
        tc_fence_after();
        if (ksteps == 2) issue_tap_swap<PASSES, 2, U, STK>(da_row, x_hi, dw_base + (uint32_t)(s * stage16), w_hi, a_plane, b_half, kstep, uoff, ucol, idesc_b, acc, idesc_2);
        else issue_tap_swap<PASSES, 1, U, STK>(da_row, x_hi, dw_base + (uint32_t)(s * stage16), w_hi, a_plane, b_half, kstep, uoff, ucol, idesc_b, acc, idesc_2);
        tc_commit_elect(&cx.w_empty[s]);
        acc = 1;
        if (++s == nstage) { s = 0; ph ^= 1; }
        da_row += row16;
      }
      tc_commit_elect(&cx.a_empty[buf]);
    }
    tc_commit_elect(&cx.acc_full[a]);
  }
}

// Epilogue of the row-stacked thin path.  TMEM holds, for lane l (pixel row r0 + l of the flat padded image) and filter
// column dx, D[l][dx*np + c] = sum_{dy,cin} W[dy,dx][cin][c] * F[r0 + l + dy*Wp]; the output of home lane l0 is
// sum_dx D[l0 + dx][dx*np + c].  Lanes of the same warp exchange by shuffle, the first KS-1 lanes of every warp publish their
// blocks in shared memory for the previous warp (double-buffered, one 128-thread named barrier per 16-channel block).
// Only home lanes < 128-(KS-1) produce outputs: consecutive units overlap by KS-1 rows.
template <int KS, int NG>
__device__ __forceinline__ void epilogue_swap2(const Params& P, const Geometry& g, uint32_t tmem_base, int a, int acc_cols, int q4,
                                               int lane, int half, int b, int v_lo, uint32_t xch_half, uint32_t tab, uint32_t stg,
                                               float (&stat_acc)[NG >= 3 ? 1 : 4]) {
  const int row = q4 * 32 + lane;
  const int ps = P.out.pix_stride, W = g.W, H = g.H, Wp = g.Wp, pad = g.padc;
  const float scale = g.out_scale;
  const bool sigm = P.act == ACT_SIGMOID;
  const bool vec4 = ((g.Cout | P.out.ch_off | ps) & 3) == 0 && (P.out.sample_stride & 3) == 0 && (P.out.lo_off & 3) == 0;
  const bool split = P.out.lo_off != 0;
  // float32 outputs with whole 16-channel blocks leave through a per-warp shared-memory tile (stg: 32 pixels x 64 bytes,
  // XOR-swizzled 16-byte chunks): a store instruction then covers 8 pixels x 64 contiguous bytes (full sectors) instead of
  // 16 bytes in each of 32 different lines — the uncoalesced form kept the store queue full (STG operand-release stalls).
  const bool tiled = vec4 && !split && !sigm && (g.Cout & 15) == 0;
  uint32_t par = 0;
  const bool home = row < g.ustride && b < P.B;
  const long long obase = (long long)b * P.out.sample_stride + P.out.ch_off;      // the item's sample: once per item
  float* const out_b = P.out.p + obase;
  // work items = (unit, 16-channel block), dealt round-robin to the NG groups of the epilogue (4 warps each); (u, cb) advance
  // incrementally (ncb is 1..4: no division)
  const int ncb = g.np >> 4;
  int u = 0, cb = half;
  while (cb >= ncb) { cb -= ncb; ++u; }
  for (; u < g.units;) {
    const int c16 = cb << 4;
    const int v = v_lo + u * g.ustride + row;
    const int oy = fdiv(v, g.m_Wp), ox = v - oy * Wp;
    const bool valid = home && ox < W && oy < H;
    uint32_t sb = tab;                        // this sample's (border class, channel) bias table, staged in shared memory (address)
    int pixo = -1;                            // element offset of the output pixel inside the sample (-1: no output from this lane)
    if (valid) {
      if (P.sabias) sb = tab + (uint32_t)((border_class(oy, H, pad) * g.kcl + border_class(ox, W, pad)) * g.np) * 4u;
      pixo = (oy * W + ox) * ps;
    }
    {
      uint32_t r[KS][16];
      const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * acc_cols + u * g.ncols * (1 + g.stk) + c16);
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) tmem_ld16_nowait(t0 + (uint32_t)(dx * g.np), r[dx]);
      tmem_wait_ld();
      if (NG == 3 && g.stk) {                   // + the X_hi * W_lo products of the stacked pass (their own ncols columns); NG = 3 only:
                                                // the host launches stacked layers on the 128-register instance (no room at 96 / in the 5x5 epilogue)
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
#pragma unroll
          for (int h8 = 0; h8 < 16; h8 += 8) {        // 8 columns at a time: the 96-register epilogue (NG = 4) has no room for 16
            uint32_t t[8];
            tmem_ld8(t0 + (uint32_t)(g.ncols + dx * g.np + h8), t);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[dx][h8 + j] = __float_as_uint(__uint_as_float(r[dx][h8 + j]) + __uint_as_float(t[j]));
          }
        }
      }
      const uint32_t xw = xch_half + par * (uint32_t)(4 * (KS - 1) * (KS - 1) * 16 * 4);
      if (lane < KS - 1) {
#pragma unroll
        for (int dx = 1; dx < KS; ++dx) {
          const uint32_t d = xw + (uint32_t)((((q4 * (KS - 1) + (dx - 1)) * (KS - 1) + lane) * 16) * 4);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(d + 16 * j, make_float4(__uint_as_float(r[dx][4 * j]), __uint_as_float(r[dx][4 * j + 1]),
                                           __uint_as_float(r[dx][4 * j + 2]), __uint_as_float(r[dx][4 * j + 3])));
        }
      }
      named_bar_sync(1 + half, 128);
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[0][j]);
#pragma unroll
      for (int dx = 1; dx < KS; ++dx) {
        const bool from_next = lane + dx >= 32;
        const bool use_x = from_next && q4 < 3;          // the last warp's trailing lanes are not home lanes
        const uint32_t xs = xw + (uint32_t)(((((q4 + 1) * (KS - 1) + (dx - 1)) * (KS - 1) + (use_x ? lane + dx - 32 : 0)) * 16) * 4);
        float4 x4[4];
        if (use_x) {
#pragma unroll
          for (int j = 0; j < 4; ++j) x4[j] = lds128(xs + 16 * j);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float t = __shfl_down_sync(0xffffffffu, __uint_as_float(r[dx][j]), dx);
          if (from_next) {
            const float4 q = x4[j >> 2];
            t = use_x ? ((j & 3) == 0 ? q.x : ((j & 3) == 1 ? q.y : ((j & 3) == 2 ? q.z : q.w))) : 0.f;
          }
          acc[j] += t;
        }
      }
      par ^= 1;
      if (tiled) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = lds128(sb + (uint32_t)(c16 + j) * 4u);
          sts128(stg + (uint32_t)(lane * 64 + ((((j >> 2) ^ (lane >> 1)) & 3) << 4)),
                 make_float4(fmaf(acc[j], scale, bb.x), fmaf(acc[j + 1], scale, bb.y), fmaf(acc[j + 2], scale, bb.z), fmaf(acc[j + 3], scale, bb.w)));
        }
        __syncwarp();
        const int q = lane & 3;
        float* const out_q = out_b + c16 + 4 * q;
        float sv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // sums / sums of squares of this lane's 4 channels over its 4 pixels
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int pp = 8 * s4 + (lane >> 2);
          const float4 v4 = lds128(stg + (uint32_t)(pp * 64 + (((q ^ (pp >> 1)) & 3) << 4)));
          const int po = __shfl_sync(0xffffffffu, pixo, pp);
          if (po >= 0) {
            *reinterpret_cast<float4*>(out_q + po) = v4;
            sv[0] += v4.x; sv[1] += v4.y; sv[2] += v4.z; sv[3] += v4.w;
            sv[4] = fmaf(v4.x, v4.x, sv[4]); sv[5] = fmaf(v4.y, v4.y, sv[5]);
            sv[6] = fmaf(v4.z, v4.z, sv[6]); sv[7] = fmaf(v4.w, v4.w, sv[7]);
          }
        }
        if (P.stats_partial) {
          // Instance-norm statistics of the stored outputs (no separate pass over the conv output): reduce-scatter over the 8
          // lanes that hold the same channel quad (fixed tree -> bit-reproducible).  Afterwards lane l holds ONE value:
          // kind = l >> 4 (0 sum, 1 sum of squares) of channel c16 + 4*(l & 3) + 2*((l >> 3) & 1) + ((l >> 2) & 1),
          // summed over the warp's 32 pixel rows.
          const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
          float t4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float snd = b4 ? sv[i] : sv[4 + i], kp = b4 ? sv[4 + i] : sv[i];
            t4[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 16);
          }
          float t2[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float snd = b3 ? t4[i] : t4[2 + i], kp = b3 ? t4[2 + i] : t4[i];
            t2[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 8);
          }
          const float snd = b2 ? t2[0] : t2[1], kp = b2 ? t2[1] : t2[0];
          const float tot = kp + __shfl_xor_sync(0xffffffffu, snd, 4);
          if constexpr (NG >= 3) {
            stat_acc[0] += tot;                 // NG % (channel blocks) == 0 (host-checked): this group always sees the same block
          } else {
            stat_acc[0] += cb == 0 ? tot : 0.f; stat_acc[1] += cb == 1 ? tot : 0.f;
            stat_acc[2] += cb == 2 ? tot : 0.f; stat_acc[3] += cb == 3 ? tot : 0.f;
          }
        }
        __syncwarp();
      } else if (valid) {
        const long long oo = obase + pixo;
        if (vec4) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (c16 + j < g.Cout) {
              const float4 bb = lds128(sb + (uint32_t)(c16 + j) * 4u);
              float4 o;
              o.x = fmaf(acc[j], scale, bb.x);
              o.y = fmaf(acc[j + 1], scale, bb.y);
              o.z = fmaf(acc[j + 2], scale, bb.z);
              o.w = fmaf(acc[j + 3], scale, bb.w);
              if (sigm) {
                o.x = __fdividef(1.f, 1.f + __expf(-o.x)); o.y = __fdividef(1.f, 1.f + __expf(-o.y));
                o.z = __fdividef(1.f, 1.f + __expf(-o.z)); o.w = __fdividef(1.f, 1.f + __expf(-o.w));
              }
              if (split) vst4(P.out, oo + c16 + j, o);
              else *reinterpret_cast<float4*>(P.out.p + oo + c16 + j) = o;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c16 + j < g.Cout) {
              float o = fmaf(acc[j], scale, lds32(sb + (uint32_t)(c16 + j) * 4u));
              if (sigm) o = __fdividef(1.f, 1.f + __expf(-o));
              if (split) vst1(P.out, oo + c16 + j, o);
              else P.out.p[oo + c16 + j] = o;
            }
          }
        }
      }
    }
    cb += NG;
    while (cb >= ncb) { cb -= ncb; ++u; }
  }
}

// thread layout: warp 0 weight producer, warp 1 MMA issuer (+TMEM alloc), warp 2 activation producer (tensor TMA),
// warp 3 idle, warps 4.. epilogue in NG groups of 4 warps (warp % 4 = TMEM lane quarter; the groups split the columns /
// units).  NG = 2 serves every tiling; NG = 3 / 4 (512 / 640 threads, <= 128 / 96 registers) are the row-stacked 3x3 thin path
// only, whose epilogue is latency-bound (TMEM load -> exchange -> barrier -> shuffles -> stores per work item): an item of
// 2 units x 2 channel blocks is 4 work items, one per group with NG = 4.
constexpr int EPI_WARP0 = 4;
constexpr int TAB_BAR = 7;             // named barrier of the bias-table hand-over (ids 1..NG: the groups' exchange barriers)
constexpr int NPRE = 5;                // bias-table entries prefetched per epilogue thread (ntap * np <= 25 * 48 = 1200)
constexpr uint32_t STG_OFF = 2 * 25 * MT * 4 + 256;   // output tiles of the row-stacked epilogue: behind the bias tables + barriers
constexpr size_t STG_TILE_BYTES = 16 * 2048;          // 2 KB per epilogue warp, up to 16 warps (NG = 4)
constexpr uint32_t STAT_OFF = STG_OFF + (uint32_t)STG_TILE_BYTES;   // fused instance-norm statistics: [2 buffers][16 warps][4 channel blocks][32 lanes] floats
constexpr size_t STAT_BYTES = 2 * 16 * 4 * 32 * 4;
constexpr size_t STG_BYTES = STG_TILE_BYTES + STAT_BYTES;

// row-stacked thin layers with at least two units per item: warps 1 and 3 both issue (VF_DUAL_ISSUE=0 switches it off)
__device__ __forceinline__ bool dual_issue(const Geometry& g) { return g.swap == 2 && g.units >= 2 && g.dual; }

template <int NG>
__global__ void __launch_bounds__(128 + 128 * NG, 1) k_conv_mma(const __grid_constant__ Params P) {
  constexpr int NEPI = 128 * NG;
  extern __shared__ uint8_t smem_raw[];
  const Geometry& g = P.g;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;        // 1024-aligned: swizzle phases are address based
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t act_base = sbase;                                       // [buf][hi|lo] planes
  const uint32_t wst_base = sbase + (uint32_t)(g.nbuf * 2 * g.plane_bytes);
  uint8_t* tail = smem + g.nbuf * 2 * g.plane_bytes + g.nstage * g.stage_bytes;
  float* s_sab_all = reinterpret_cast<float*>(tail);                     // [2 halves][25 classes][128 channels] border-class bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 2 * 25 * MT * sizeof(float));
  uint64_t *w_full = bars, *w_empty = bars + MAX_STAGE, *a_full = bars + 2 * MAX_STAGE, *a_empty = a_full + 2;
  uint64_t *acc_full = a_empty + 2, *acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    const uint32_t nissue = dual_issue(g) ? 2u : 1u;      // issuer warps committing to the stage / buffer / accumulator barriers
    for (int i = 0; i < MAX_STAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], nissue); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], nissue);
      mbar_init(&acc_full[i], nissue); mbar_init(&acc_empty[i], NEPI);
    }
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
  // barrier init and the TMEM allocation above overlap the tail of the previous kernel (programmatic dependent launch);
  // nothing before this point touches global memory
  pdl_wait();
  pdl_trigger_conv();

  const int per_mt = g.ngroups * g.npass;
  const uint32_t ltype = g.layout == 1 ? 4u : 2u;
  const int acc_cols = g.swap ? g.units * g.ncols * (1 + g.stk) : (g.rg ? g.ncols_item : g.G * g.v_cnt);   // TMEM columns of one accumulator set
  const int nacc = g.nacc;                               // 2 when two sets fit in the 512 columns

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = g.passes == 3 ? (uint32_t)g.stage_bytes : (uint32_t)g.half_bytes;
      for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
        int part_;
        const int mt = strip_of(g, item, part_) / per_mt;
        const uint8_t* wbase = reinterpret_cast<const uint8_t*>(P.w) + (size_t)mt * g.nchunk * g.nst * g.stage_bytes;
        for (int ct = 0; ct < g.nchunk * g.nst; ++ct) {
          mbar_wait(&w_empty[s], ph ^ 1);
          mbar_expect_tx(&w_full[s], bytes);
          bulk_g2s(wst_base + s * g.stage_bytes, wbase + (size_t)ct * g.stage_bytes, bytes, &w_full[s]);
          if (++s == g.nstage) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || (warp == 3 && dual_issue(g))) {
    // ===== MMA issuer: the whole warp walks the loops (warp-uniform), one elected lane issues =====
    const uint32_t sbo_b = (uint32_t)(8 * g.row_bytes);                   // 8-row core-matrix group pitch
    IssueCtx cx;
    cx.kstep_a = cx.kstep_b = 32u >> 4;                                   // start-address advance per K=16 step (16-byte units)
    cx.pix_b = (uint32_t)g.row_bytes;
    cx.da_zero = make_desc(0, 16u, sbo_b, ltype, 0);
    cx.db_zero = make_desc(0, 16u, g.rg ? (uint32_t)(g.Wp * g.row_bytes) : sbo_b, ltype, 0);   // rg: group pitch = padded image row
    cx.a_half = (uint64_t)(g.half_bytes >> 4); cx.b_plane = (uint64_t)(g.plane_bytes >> 4);
    cx.act_base = act_base; cx.wst_base = wst_base; cx.tmem_base = tmem_base;
    cx.w_full = w_full; cx.w_empty = w_empty; cx.a_full = a_full; cx.a_empty = a_empty; cx.acc_full = acc_full; cx.acc_empty = acc_empty;
    if (NG >= 3 || g.swap == 2) {
      // units of this warp: all of them, or with two issuers the first ceil(U/2) (warp 1) / the rest (warp 3)
      const bool dual = dual_issue(g);
      const int first = (dual && warp == 3) ? (g.units + 1) / 2 : 0;
      const int count = dual ? (warp == 3 ? g.units - first : (g.units + 1) / 2) : g.units;
#define VF_S2(PA, UU) issuer_loop_swap2<PA, UU>(g, cx, acc_cols, first)
      if (g.passes == 3 && g.stk) {                     // stacked weight halves: <= 2 units per accumulator set
        if (count == 1) issuer_loop_swap2<3, 1, 1>(g, cx, acc_cols, first);
        else issuer_loop_swap2<3, 2, 1>(g, cx, acc_cols, first);
      } else if (g.passes == 3) {
        switch (count) {
          case 1: VF_S2(3, 1); break; case 2: VF_S2(3, 2); break; case 3: VF_S2(3, 3); break; case 4: VF_S2(3, 4); break;
          default: VF_S2(3, 5); break;
        }
      } else {
        switch (count) {
          case 1: VF_S2(1, 1); break; case 2: VF_S2(1, 2); break; case 3: VF_S2(1, 3); break; case 4: VF_S2(1, 4); break;
          default: VF_S2(1, 5); break;
        }
      }
#undef VF_S2
    } else if constexpr (NG != 2) {
    } else if (g.swap) {
#define VF_SW(PA, UU) issuer_loop_swap<PA, UU>(g, cx, acc_cols)
      if (g.passes == 3) {
        switch (g.units) {
          case 1: VF_SW(3, 1); break; case 2: VF_SW(3, 2); break; case 3: VF_SW(3, 3); break; case 4: VF_SW(3, 4); break;
          case 5: VF_SW(3, 5); break; case 6: VF_SW(3, 6); break; case 7: VF_SW(3, 7); break; default: VF_SW(3, 8); break;
        }
      } else {
        switch (g.units) {
          case 1: VF_SW(1, 1); break; case 2: VF_SW(1, 2); break; case 3: VF_SW(1, 3); break; case 4: VF_SW(1, 4); break;
          case 5: VF_SW(1, 5); break; case 6: VF_SW(1, 6); break; case 7: VF_SW(1, 7); break; default: VF_SW(1, 8); break;
        }
      }
#undef VF_SW
    } else {
      const int U = g.rg == 1 ? 1 : g.G * g.nseg;
#define VF_IS(PA, UU)                                       \
  if (g.ksteps == 2) issuer_loop<PA, 2, UU>(g, cx, acc_cols); \
  else issuer_loop<PA, 4, UU>(g, cx, acc_cols)
      if (g.passes == 3) {
        switch (U) { case 1: VF_IS(3, 1); break; case 2: VF_IS(3, 2); break; case 3: VF_IS(3, 3); break; default: VF_IS(3, 4); break; }
      } else {
        switch (U) { case 1: VF_IS(1, 1); break; case 2: VF_IS(1, 2); break; case 3: VF_IS(1, 3); break; default: VF_IS(1, 4); break; }
      }
#undef VF_IS
    }
  } else if (warp == 2) {
    // ===== activation producer: one tensor-TMA box per (image, plane) and channel chunk =====
    if (lane == 0) {
      const int nplanes = g.passes == 3 ? 2 : 1;
      uint32_t job = 0;
      for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
        int part_;
        const int rem = strip_of(g, item, part_) % per_mt;      // a half strip stages the whole strip's box (same tensor map)
        const int grp = rem / g.npass, ps = rem % g.npass;
        const int b0 = grp * g.G;
        const int qy0 = g.rg == 3 ? 0 : (ps * g.v_cnt) / g.Wp;  // first padded-image row the item touches (rg 3: the whole column strip)
        const int qx0 = (g.rg == 3 ? ps * 8 : 0) - g.padx;      // rg 3: the item's 8-pixel column group
        const int nimg = min(g.G, P.B - b0);
        for (int c = 0; c < g.nchunk; ++c, ++job) {
          const int buf = job % g.nbuf;
          mbar_wait(&a_empty[buf], ((job / g.nbuf) & 1) ^ 1);
          const CUtensorMap* tm = c < g.nchunk0 ? &P.tmap0 : &P.tmap1;
          const int c0 = (c < g.nchunk0 ? c : c - g.nchunk0) * g.ch;
          mbar_expect_tx(&a_full[buf], (uint32_t)(nimg * nplanes * g.box_bytes));
          for (int im = 0; im < nimg; ++im)
            for (int pl = 0; pl < nplanes; ++pl)
              tma_load_5d(act_base + (uint32_t)((buf * 2 + pl) * g.plane_bytes + im * g.img_pix * g.row_bytes), tm, c0, qx0,
                          qy0 - g.pad, b0 + im, pl, &a_full[buf]);
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue: TMEM -> registers -> (x 2^-s, + border-class bias) -> NHWC global =====
    const int q4 = warp & 3;                        // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;                 // output channel within the cout tile
    const int half = (warp - EPI_WARP0) >> 2;       // the two warps of a quarter take alternate column chunks / units
    float* s_sab = s_sab_all + (half & 1) * 25 * MT;
    // Row-stacked path: the sample's bias table [border class][channel] is staged in shared memory (the L1 is carved out to
    // almost nothing by the operand buffers: a global bias load per output vector costs an L2 round trip).  The table of
    // item i+1 is fetched into registers BEFORE the epilogue of item i and handed over after it, so the L2 latency of the
    // fetch never sits on the epilogue's critical path.
    const int etid = (int)threadIdx.x - EPI_WARP0 * 32;
    const int tabn = (P.sabias ? g.ntap : 1) * g.np;
    float pre[NPRE];
    auto tab_fetch = [&](int item_) {
      const int b_ = fdiv(item_, g.m_npass);       // row-stacked path: one Cout tile, one sample per item -> item = sample * npass + pass
#pragma unroll
      for (int j = 0; j < NPRE; ++j) {
        const int i = etid + j * NEPI;
        float bv = 0.f;
        if (i < tabn) {
          const int cls = fdiv(i, g.m_np), c = i - cls * g.np;
          if (c < g.Cout) bv = P.sabias ? __ldg(P.sabias + ((long long)b_ * g.ntap + cls) * g.Cout + c) : (P.bias ? __ldg(P.bias + c) : 0.f);
        }
        pre[j] = bv;
      }
    };
    auto tab_store = [&](float* t) {
#pragma unroll
      for (int j = 0; j < NPRE; ++j) {
        const int i = etid + j * NEPI;
        if (i < tabn) t[i] = pre[j];
      }
    };
    const bool stacked = NG >= 3 || g.swap == 2;
    float* const tab0 = s_sab_all + 4096;
    if (stacked && (int)blockIdx.x < g.nitems) {
      tab_fetch(blockIdx.x);
      tab_store(tab0);
      named_bar_sync(TAB_BAR, NEPI);
    }
    uint32_t it = 0;
    constexpr int NACC = NG >= 3 ? 1 : 4;       // row-stacked path: this lane's statistic per channel block over the item (epilogue_swap2);
    float stat_acc[NACC];                       // with NG >= 3 groups a group always handles the same block (NG % blocks == 0)
#pragma unroll
    for (int i = 0; i < NACC; ++i) stat_acc[i] = 0.f;
    const bool per_item_tab = stacked && P.sabias != nullptr;   // without the action/state bias the table is the layer's bias: staged once
    // "Last arriver finalises" (Params::stats_fin).  The arrival of an item is issued at the START of the next item's epilogue
    // (and after the loop for the last one): a fence right behind the item's output stores would wait for all of them to
    // drain (measured: thin layers 2x slower), one item later they have long left the SM and the fence is free.
    int pend_b = -1, pend_mt = 0;
    auto thin_arrive = [&](int b_) {                 // row-stacked path: warp etid >> 5 combined channel block cb of the item
      const int cb = etid >> 5, l = etid & 31;
      __threadfence();
      __syncwarp();
      int last = 0;
      if (l == 0) {
        int* cnt = P.stats_cnt + (long long)b_ * VF_STAT_CNT_STRIDE + cb;
        last = atomicAdd(cnt, 1) == g.npass - 1;     // the sample's last pass finalises the block
        if (last) *cnt = 0;
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        const int c = cb * 16 + l;
        if (l < 16 && c < g.Cout)
          stats_finalize_plane(P.stats_partial + ((long long)b_ * g.Cout + c) * P.stats_S * 2, P.stats_S, P.stats_npix, P.stats_eps,
                               P.stats_fin + ((long long)b_ * g.Cout + c) * 2);
      }
    };
    auto wide_arrive = [&](int b0_, int mt_) {       // wide path: one arrival per (image of the item, warp): 2 column-halves x npass
      if (mt_ * MT + q4 * 32 >= g.Cout) return;      //   per 32-channel group (warp-uniform: no live channel in this warp)
      const int n_ = mt_ * MT + row;
      __threadfence();
      __syncwarp();
      for (int im = 0; im < g.G; ++im) {
        const int b_ = b0_ + im;
        if (b_ >= P.B) break;
        int last = 0;
        if (lane == 0) {
          int* cnt = P.stats_cnt + (long long)b_ * VF_STAT_CNT_STRIDE + (mt_ * 4 + q4);
          last = atomicAdd(cnt, 1) == 2 * g.npass - 1;
          if (last) *cnt = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          __threadfence();
          if (n_ < g.Cout)
            stats_finalize_plane(P.stats_partial + ((long long)b_ * g.Cout + n_) * P.stats_S * 2, P.stats_S, P.stats_npix, P.stats_eps,
                                 P.stats_fin + ((long long)b_ * g.Cout + n_) * 2);
        }
      }
    };
    auto flush_arrival = [&]() {
      if (pend_b < 0) return;
      if (stacked) { if (etid < (g.np >> 4) * 32) thin_arrive(pend_b); }
      else if (!g.swap) wide_arrive(pend_b, pend_mt);
      pend_b = -1;
    };
    for (int item = blockIdx.x; item < g.nitems; item += gridDim.x, ++it) {
      int mt, grp, ps_, part = -1;
      if (stacked) {                               // one Cout tile, G = 1: no divisions by runtime values on this path
        mt = 0; grp = fdiv(item, g.m_npass); ps_ = item - grp * g.npass;
      } else {
        const int strip = strip_of(g, item, part);
        mt = strip / per_mt;
        const int rem = strip % per_mt;
        grp = rem / g.npass; ps_ = rem % g.npass;
      }
      const int hcols = g.v_cnt >> 1;                 // rg 3: columns of a half strip
      const int b0 = grp * g.G, v_lo = g.rg == 3 ? (part > 0 ? hcols : 0) : ps_ * g.v_cnt;
      const int ncols_it = part >= 0 ? hcols : g.v_cnt;
      const int n = mt * MT + row;
      const int a = nacc == 2 ? (int)(it & 1u) : 0;
      const int next = item + (int)gridDim.x;
      if (per_item_tab && next < g.nitems) tab_fetch(next);
      mbar_wait(&acc_full[a], (nacc == 2 ? it >> 1 : it) & 1);
      tc_fence_after();
      flush_arrival();                               // the previous item's statistics (its stores have drained by now)
      if (stacked) {
        float* tab = per_item_tab ? tab0 + (it & 1) * (g.ntap * g.np) : tab0;
        // lane-exchange scratch (the wide path's bias tables live here): 2 parities x 4 quarters x (KS-1)^2 rows x 16 floats
        const uint32_t stg = smem_u32(tail) + STG_OFF + (uint32_t)(warp - EPI_WARP0) * 2048u;   // this warp's output tile
        if constexpr (NG >= 3) {
          epilogue_swap2<3, NG>(P, g, tmem_base, a, acc_cols, q4, lane, half, b0, v_lo, smem_u32(s_sab_all + half * 512), smem_u32(tab), stg, stat_acc);
        } else {
          const uint32_t xch_half = smem_u32(s_sab_all + half * 2048);
          if (g.k == 3) epilogue_swap2<3, 2>(P, g, tmem_base, a, acc_cols, q4, lane, half, b0, v_lo, xch_half, smem_u32(tab), stg, stat_acc);
          else epilogue_swap2<5, 2>(P, g, tmem_base, a, acc_cols, q4, lane, half, b0, v_lo, xch_half, smem_u32(tab), stg, stat_acc);
        }
        const bool do_stats = P.stats_partial != nullptr;
        const uint32_t stat_buf = smem_u32(tail) + STAT_OFF + (it & 1u) * (uint32_t)(STAT_BYTES / 2);
        if (do_stats) {                      // publish this warp's per-channel partial sums of the item (double-buffered by item parity)
          const uint32_t my = stat_buf + (uint32_t)(((warp - EPI_WARP0) * 4) * 32 + lane) * 4u;
          if constexpr (NACC == 1) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(my + (uint32_t)((half % (g.np >> 4)) * 128)), "f"(stat_acc[0]) : "memory");
            stat_acc[0] = 0.f;
          } else {
#pragma unroll
            for (int cb = 0; cb < NACC; ++cb) {
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(my + (uint32_t)(cb * 128)), "f"(stat_acc[cb]) : "memory");
              stat_acc[cb] = 0.f;
            }
          }
        }
        if (per_item_tab && next < g.nitems) tab_store(tab0 + ((it + 1) & 1) * (g.ntap * g.np));   // every reader of the other table finished one item ago
        if ((per_item_tab && next < g.nitems) || do_stats) named_bar_sync(TAB_BAR, NEPI);
        if (do_stats && etid < (g.np >> 4) * 32) {
          // slot = pass: every (sample, channel, pass) is written exactly once; the epilogue warps are combined in a fixed order
          const int cb = etid >> 5, l = etid & 31;
          double tot = 0.0;
          for (int w = 0; w < 4 * NG; ++w)
            if (NACC > 1 || (w >> 2) % (g.np >> 4) == cb) tot += (double)lds32(stat_buf + (uint32_t)((w * 4 + cb) * 32 + l) * 4u);
          const int ch = cb * 16 + 4 * (l & 3) + 2 * ((l >> 3) & 1) + ((l >> 2) & 1);
          if (ch < g.Cout && b0 < P.B)
            P.stats_partial[(((long long)b0 * g.Cout + ch) * P.stats_S + ps_) * 2 + (l >> 4)] = tot;
          if (P.stats_fin && b0 < P.B) pend_b = b0;    // arrival deferred to the next item (see thin_arrive)
        }
      } else if constexpr (NG != 2) {
      } else if (g.swap) {
        // pixels on lanes: this thread owns pixel row `row` of every 128-pixel unit and all output channels of it
        const int b = b0;
        const int ps = P.out.pix_stride, W = g.W, H = g.H, Wp = g.Wp, pad = g.padc;
        const float scale = g.out_scale;
        const bool sigm = P.act == ACT_SIGMOID;
        const bool vec4 = ((g.Cout | P.out.ch_off | ps) & 3) == 0 && (P.out.sample_stride & 3) == 0 && (P.out.lo_off & 3) == 0;
        const bool split = P.out.lo_off != 0;                  // split-half output (feeds another convolution)
        for (int u = half; u < g.units; u += 2) {
          const int v = v_lo + u * 128 + row;
          const int oy = v / Wp, ox = v - oy * Wp;
          const bool valid = b < P.B && ox < W && oy < H;
          const float* sb = nullptr;
          float* op = nullptr;
          long long oo = 0;
          if (valid) {
            if (P.sabias) {
              const int cls = border_class(oy, H, pad) * g.kcl + border_class(ox, W, pad);
              sb = P.sabias + ((long long)b * g.ntap + cls) * g.Cout;
            } else {
              sb = P.bias;
            }
            oo = (long long)b * P.out.sample_stride + (long long)(oy * W + ox) * ps + P.out.ch_off;
            op = P.out.p + oo;
          }
          for (int c16 = 0; c16 < g.np; c16 += 16) {
            float4 bbv[4];
            if (valid && vec4) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4)
                bbv[j4] = (sb && c16 + 4 * j4 < g.Cout) ? __ldg(reinterpret_cast<const float4*>(sb + c16 + 4 * j4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * acc_cols + u * g.np + c16), r);
            if (!valid) continue;
            if (vec4) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                if (c16 + j < g.Cout) {
                  const float4 bb = bbv[j >> 2];
                  float4 o;
                  o.x = fmaf(__uint_as_float(r[j]), scale, bb.x);
                  o.y = fmaf(__uint_as_float(r[j + 1]), scale, bb.y);
                  o.z = fmaf(__uint_as_float(r[j + 2]), scale, bb.z);
                  o.w = fmaf(__uint_as_float(r[j + 3]), scale, bb.w);
                  if (sigm) {
                    o.x = __fdividef(1.f, 1.f + __expf(-o.x)); o.y = __fdividef(1.f, 1.f + __expf(-o.y));
                    o.z = __fdividef(1.f, 1.f + __expf(-o.z)); o.w = __fdividef(1.f, 1.f + __expf(-o.w));
                  }
                  if (split) vst4(P.out, oo + c16 + j, o);
                  else *reinterpret_cast<float4*>(op + c16 + j) = o;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c16 + j < g.Cout) {
                  float o = fmaf(__uint_as_float(r[j]), scale, sb ? __ldg(sb + c16 + j) : 0.f);
                  if (sigm) o = __fdividef(1.f, 1.f + __expf(-o));
                  if (split) vst1(P.out, oo + c16 + j, o);
                  else op[c16 + j] = o;
                }
              }
            }
          }
        }
      } else
      for (int im = 0; im < g.G; ++im) {
        const int b = b0 + im;
        const bool live = b < P.B && n < g.Cout;
        if (live) {                                 // this thread's column of the per-sample class table (private: no sync)
          if (P.sabias) {
            const float* sp = P.sabias + (long long)b * g.ntap * g.Cout + n;
            for (int cls = 0; cls < g.ntap; ++cls) s_sab[cls * MT + row] = __ldg(sp + (long long)cls * g.Cout);
          } else {
            const float bv = P.bias ? __ldg(P.bias + n) : 0.f;
            for (int cls = 0; cls < g.ntap; ++cls) s_sab[cls * MT + row] = bv;
          }
        }
        float* op = P.out.p + (long long)b * P.out.sample_stride + P.out.ch_off + n;
        const int ps = P.out.pix_stride, W = g.W, H = g.H, Wp = g.We, pad = g.padc, kk = g.kcl;
        const int im_cols = g.rg ? g.col_stride : g.v_cnt;
        const float scale = g.out_scale;
        const bool sigm = P.act == ACT_SIGMOID;
        double st_s = 0.0, st_q = 0.0;              // instance-norm statistics of this thread's channel (fused: no extra pass)
        double st_s1 = 0.0, st_q1 = 0.0;            // rg 3: the lower half of the strip has its own slot (see below)
        for (int cc = half * 32; cc < ncols_it; cc += 64) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * acc_cols + im * im_cols + cc), r);
          if (!live) continue;
          // Every lane handles the SAME pixel (a different channel).  Branch-free and without loop-carried state so the
          // 32 elements overlap: a 32-column chunk spans at most 4 image rows (Wp >= 10), selected by compare-and-add.
          const int xoff = g.rg == 2 ? (cc / g.half_cols) * 8 : (g.rg == 3 ? ps_ * 8 : 0);   // rg 2: second column group of the rows; rg 3: the item's
          const int v = v_lo + (g.rg == 2 ? cc % g.half_cols : cc);
          const int oy0 = v / Wp, ox0 = v - oy0 * Wp;
          int cyk[4], rbase[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            cyk[w] = border_class(oy0 + w, H, pad) * kk;
            rbase[w] = (oy0 + w) * W * ps;
          }
          float cs = 0.f, cq = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int x0 = ox0 + j;
            const int wr = (x0 >= Wp) + (x0 >= 2 * Wp) + (x0 >= 3 * Wp);
            const int x = x0 - wr * Wp + xoff;
            const bool valid = (x < W) & (oy0 + wr < H);
            const int cx = x < pad ? x : (x >= W - pad ? x - (W - 1 - 2 * pad) : pad);
            const int cyw = wr == 0 ? cyk[0] : (wr == 1 ? cyk[1] : (wr == 2 ? cyk[2] : cyk[3]));
            const int rb = wr == 0 ? rbase[0] : (wr == 1 ? rbase[1] : (wr == 2 ? rbase[2] : rbase[3]));
            const float bias = s_sab[(valid ? cyw + cx : 0) * MT + row];
            float val = fmaf(__uint_as_float(r[j]), scale, bias);
            if (sigm) val = __fdividef(1.f, 1.f + __expf(-val));
            if (valid) {
              op[rb + x * ps] = val;
              cs += val;                            // fp32 within the 32-element chunk, float64 across chunks
              cq = fmaf(val, val, cq);
            }
          }
          if (g.rg == 3 && v >= hcols) { st_s1 += (double)cs; st_q1 += (double)cq; }
          else { st_s += (double)cs; st_q += (double)cq; }
        }
        if (P.stats_partial && live) {              // slot = (pass, column-half): every slot is written exactly once
          if (g.rg == 3) {
            // (strip, column-half, upper / lower half of the strip): a whole strip writes both of its half slots, a half strip of
            // the split last round its own — the partial sums and their order never depend on whether the round was split
            double* o = P.stats_partial + (((long long)b * g.Cout + n) * P.stats_S + (ps_ * 2 + half) * 2) * 2;
            if (part <= 0) { o[0] = st_s; o[1] = st_q; }
            if (part != 0) { o[2] = st_s1; o[3] = st_q1; }
          } else {
            double* o = P.stats_partial + (((long long)b * g.Cout + n) * P.stats_S + (ps_ * 2 + half)) * 2;
            o[0] = st_s;
            o[1] = st_q;
          }
        }
        if (P.stats_fin && im == 0) { pend_b = b0; pend_mt = mt; }   // arrivals of the item's images deferred to the next item
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
    flush_arrival();
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// =====================================================================================================================
// CTA-PAIR variant (tcgen05 cta_group::2) of the wide tiling for 16x16 maps with Cout % 256 == 0 (the 16x16 conv-LSTM gate
// convolutions).  Two CTAs of a cluster (one TPC) compute ONE 256-channel x 256-pixel tile: CTA r holds the weight rows
// [128 r, 128 r + 128) of the tile (its own A half) and the 8-pixel COLUMN GROUP r of all 16 image rows (x in [8 r, 8 r + 8)
// + halo: its half of the B operand, a 12-pixel-wide box); the leader's elected lane issues ONE M = 256 x N = 256 MMA per K
// step (the single-CTA kernel needs two N = 128 MMAs, one per column group) whose B operand is read half from each SM: per
// MMA an SM reads 4 KB of weights + 4 KB of pixels for 128 clk of math, against 4 + 4 KB per 64 clk in the single-CTA kernel,
// which is bound by exactly that shared-memory traffic (ncu: L1/shared 69 %, tensor pipe 69 % on these layers).
// Accumulators: each CTA's TMEM gets its 128 channels x 256 pixels; epilogue as in the wide tiling.
//
// Synchronisation (same barrier offsets in both CTAs):
//   w_full[s] / a_full[buf]   live in the LEADER: both CTAs' loads are tensor TMA with .cta_group::2, whose completion bytes may
//                             target the peer's barrier — the leader expects the bytes of both halves (a first version relayed
//                             the peer's completions through a thread: its release-arrive cost a membar per stage and capped
//                             the kernel at 201 us).
//   w_empty / a_empty / acc_full   tcgen05.commit.cta_group::2 with multicast mask 0b11: released in both CTAs at once.
//   acc_empty[a] (leader)     16 arrivals: the 8 epilogue warps of each CTA (the peer's arrive remotely).
// =====================================================================================================================
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // acquire at cluster scope (peer-signalled barriers)
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && spins > (1u << 26)) __trap();
  }
}
// tensor TMA whose completion is counted on a barrier of EITHER CTA of the pair (mbar: shared::cluster address)
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, int c4, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {     // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

constexpr int PAIR_THREADS = 384;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1) k_conv_pair(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  const Geometry& g = P.g;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t act_base = sbase;
  const uint32_t wst_base = sbase + (uint32_t)(g.nbuf * 2 * g.plane_bytes);
  uint8_t* tail = smem + g.nbuf * 2 * g.plane_bytes + g.nstage * g.stage_bytes;
  float* s_sab_all = reinterpret_cast<float*>(tail);                     // [2 halves][25 classes][128 channels]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 2 * 25 * MT * sizeof(float));
  uint64_t *w_full = bars, *w_empty = bars + MAX_STAGE;
  uint64_t *a_full = bars + 2 * MAX_STAGE, *a_empty = a_full + 2;
  uint64_t *acc_full = a_empty + 2, *acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);       // 24 + 10 barriers = 272 bytes < the 512 reserved (PAIR_SLACK)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();                                                    // barriers of both CTAs exist before anyone signals them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger_conv();

  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int n_tile = g.Cout / 256;                                        // 256-channel tiles
  const int nitems = n_tile * P.B;
  const uint32_t ltype = 4u;                                              // SWIZZLE_64B
  const int nplanes = g.passes == 3 ? 2 : 1;

  if (warp == 0) {
    // ===== weight producer: this CTA's 128 rows of the 256-channel tile =====
    // (P.tmap1 = the packed weights as a 2-D tensor: rows of 512 bytes, one 16 KB stage = 32 rows)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t w_full_leader = map_to_cta(smem_u32(w_full), 0);
      const int rows_per_stage = g.stage_bytes / 512;
      for (int item = pair; item < nitems; item += npairs) {
        const int mt = (item / P.B) * 2 + (int)rank;
        const int stage0 = mt * g.nchunk * g.nst;
        for (int ct = 0; ct < g.nchunk * g.nst; ++ct) {
          mbar_wait(&w_empty[s], ph ^ 1);
          if (leader) mbar_expect_tx(&w_full[s], 2u * (uint32_t)g.stage_bytes);      // both CTAs' halves
          tma_load_2d_pair(wst_base + s * g.stage_bytes, &P.tmap1, 0, (stage0 + ct) * rows_per_stage, w_full_leader + (uint32_t)(s * 8));
          if (++s == g.nstage) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===== activation producer: this CTA's 8-pixel column group (+ halo) of every channel chunk =====
    if (lane == 0) {
      uint32_t job = 0;
      const uint32_t a_full_leader = map_to_cta(smem_u32(a_full), 0);
      for (int item = pair; item < nitems; item += npairs) {
        const int b = item % P.B;
        for (int c = 0; c < g.nchunk; ++c, ++job) {
          const int buf = job & 1;
          mbar_wait(&a_empty[buf], ((job >> 1) & 1) ^ 1);
          if (leader) mbar_expect_tx(&a_full[buf], 2u * (uint32_t)(nplanes * g.box_bytes));
          for (int pl = 0; pl < nplanes; ++pl)
            tma_load_5d_pair(act_base + (uint32_t)((buf * 2 + pl) * g.plane_bytes), &P.tmap0, c * g.ch, (int)rank * 8 - g.padx, -g.pad, b, pl,
                             a_full_leader + (uint32_t)(buf * 8));
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===== MMA issuer (leader CTA only): M = 256 x N = 256 over the pair, one MMA per K step =====
    const uint32_t pix_b = (uint32_t)g.row_bytes;
    const uint64_t da_zero = make_desc(0, 16u, 8u * pix_b, ltype, 0);
    const uint64_t db_zero = make_desc(0, 16u, (uint32_t)g.Wp * pix_b, ltype, 0);     // group pitch = padded image row
    const uint64_t a_half = (uint64_t)(g.half_bytes >> 4), b_plane = (uint64_t)(g.plane_bytes >> 4);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(g.v_cnt >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // D f32, A = B = f16, N = 16 * H (256 at 16x16), M = 256
    const uint32_t stage16 = (uint32_t)(g.stage_bytes >> 4), buf16 = (uint32_t)((2 * g.plane_bytes) >> 4);
    const uint64_t pix16 = (uint64_t)(pix_b >> 4), rowskip16 = (uint64_t)(((uint32_t)(g.Wp - g.kw) * pix_b) >> 4);
    const uint64_t da_base = da_zero + (uint64_t)(wst_base >> 4), db_base = db_zero + (uint64_t)(act_base >> 4);
    int s = 0;
    uint32_t ph = 0, job = 0, it = 0;
    for (int item = pair; item < nitems; item += npairs, ++it) {
      const int a = (int)(it & 1u);
      mbar_wait_cluster(&acc_empty[a], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(a * 256);
      uint32_t acc = 0;
      for (int c = 0; c < g.nchunk; ++c, ++job) {
        const int buf = job & 1;
        mbar_wait_cluster(&a_full[buf], (job >> 1) & 1);
        tc_fence_after();
        uint64_t db_tap = db_base + (uint64_t)(buf * buf16);
        int tx = 0;
        for (int tap = 0; tap < g.nst; ++tap) {
          mbar_wait_cluster(&w_full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da_st = da_base + (uint64_t)(s * stage16);
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              if (pass < g.passes) {
                const uint64_t da_p = da_st + (pass == 2 ? a_half : 0), db_p = db_tap + (pass == 1 ? b_plane : 0);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint64_t da = da_p + (uint64_t)(j * 2), db = db_p + (uint64_t)(j * 2);
                  tc_mma_f16_pair(d0, da, db, idesc, (pass | j) == 0 ? acc : 1u);   // columns [0,128) from CTA 0's group, [128,256) from CTA 1's
                }
              }
            }
            tc_commit_pair(&w_empty[s]);
          }
          acc = 1;
          if (++s == g.nstage) { s = 0; ph ^= 1; }
          db_tap += pix16;
          if (++tx == g.kw) { tx = 0; db_tap += rowskip16; }
        }
        if (elect_one()) tc_commit_pair(&a_empty[buf]);
      }
      if (elect_one()) tc_commit_pair(&acc_full[a]);
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue: this CTA's 128 channels x 256 pixels (columns [0,128) = left 8-pixel group of the 16 rows, [128,256) right) =====
    const int q4 = warp & 3, row = q4 * 32 + lane, half = (warp - EPI_WARP0) >> 2;
    float* s_sab = s_sab_all + half * 25 * MT;
    const int ps = P.out.pix_stride, W = g.W, H = g.H, pad = g.padc, kk = g.kcl;
    const float scale = g.out_scale;
    const uint32_t acc_empty_leader = map_to_cta(smem_u32(acc_empty), 0);
    uint32_t it = 0;
    for (int item = pair; item < nitems; item += npairs, ++it) {
      const int b = item % P.B;
      const int n = ((item / P.B) * 2 + (int)rank) * MT + row;
      const int a = (int)(it & 1u);
      {                                              // this thread's column of the per-sample class table (private: no sync)
        if (P.sabias) {
          const float* sp = P.sabias + (long long)b * g.ntap * g.Cout + n;
          for (int cls = 0; cls < g.ntap; ++cls) s_sab[cls * MT + row] = __ldg(sp + (long long)cls * g.Cout);
        } else {
          const float bv = P.bias ? __ldg(P.bias + n) : 0.f;
          for (int cls = 0; cls < g.ntap; ++cls) s_sab[cls * MT + row] = bv;
        }
      }
      mbar_wait_cluster(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      float* op = P.out.p + (long long)b * P.out.sample_stride + P.out.ch_off + n;
      double st_s = 0.0, st_q = 0.0;
      for (int cc = half * 32; cc < g.v_cnt; cc += 64) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * 256 + cc), r);
        const int grp = cc >= g.half_cols;           // second half of the columns = right 8-pixel group
        const int xoff = grp * 8;
        const int oy0 = (cc - grp * g.half_cols) >> 3;   // 32 columns = 4 image rows x 8 pixels
        float cs = 0.f, cq = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int oy = oy0 + (j >> 3), x = (j & 7) + xoff;
          const int cx = x < pad ? x : (x >= W - pad ? x - (W - 1 - 2 * pad) : pad);
          const float bias = s_sab[(border_class(oy, H, pad) * kk + cx) * MT + row];
          const float val = fmaf(__uint_as_float(r[j]), scale, bias);
          op[(oy * W + x) * ps] = val;
          cs += val;
          cq = fmaf(val, val, cq);
        }
        st_s += (double)cs;
        st_q += (double)cq;
      }
      if (P.stats_partial) {                         // slot = column half (2 slots per (sample, channel))
        double* o = P.stats_partial + (((long long)b * g.Cout + n) * P.stats_S + half) * 2;
        o[0] = st_s;
        o[1] = st_q;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)(a * 8));   // the leader's acc_empty[a] (8 bytes per barrier)
    }
  }
  tc_fence_before();
  cluster_sync_all();                                // nobody frees TMEM / exits while the peer still uses the pair's resources
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
int g_num_sms = 0;
constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr size_t SMEM_SLACK = 1024 + 2 * 25 * MT * 4 + 256;  // alignment slack + bias tables + barriers

void current_mode(int* layout, int* bo) {
  static int s_layout = -1, s_bo = 0;
  if (s_layout < 0) {
    const char* e = getenv("VF_MMA_LAYOUT");
    s_layout = e ? atoi(e) : 1;     // SWIZZLE_64B: verified correct with tap-shifted descriptors on B200
    if (s_layout < 1 || s_layout > 2) s_layout = 1;
    const char* b = getenv("VF_MMA_BO");
    s_bo = b ? atoi(b) : 0;
  }
  *layout = s_layout; *bo = s_bo;
}

// output tiles of the row-stacked epilogue: only layers whose epilogue can take the tiled path (whole 16-channel blocks)
size_t stg_bytes(const Geometry& g) { return (g.swap == 2 && (g.Cout & 15) == 0) ? STG_BYTES : 0; }
// host mirror of epilogue_swap2's `tiled` predicate (the path that also accumulates the instance-norm statistics)
bool thin_epilogue_tiled(const Geometry& g, const View& out, int act) {
  const bool vec4 = ((g.Cout | out.ch_off | out.pix_stride) & 3) == 0 && (out.sample_stride & 3) == 0 && (out.lo_off & 3) == 0;
  return vec4 && out.lo_off == 0 && act != ACT_SIGMOID && (g.Cout & 15) == 0 && g.np <= 64;
}

// TMA box rows: every pass starts its box at the padded-image row holding its first virtual pixel and must cover
// (offset inside that row) + v_cnt + the k-1 halo rows + k-1 pixels.
bool box_rows(Geometry& g) {
  // pixel rows one item reads past its first one: its v_cnt outputs + k-1 taps along x (row-stacked: the last unit's 128 lanes)
  const int span = g.swap == 2 ? (g.units - 1) * g.ustride + 128 : g.v_cnt + (g.kw - 1);
  int need = 0;
  for (int ps = 0; ps < g.npass; ++ps) need = std::max(need, (ps * g.v_cnt) % g.Wp + span + (g.k - 1) * g.Wp);
  g.R = (need + g.Wp - 1) / g.Wp;
  g.img_pix = (g.R * g.Wp + 7) / 8 * 8;
  g.box_bytes = g.R * g.Wp * g.row_bytes;
  return g.R <= 256 && g.Wp <= 256;
}

// VF_THIN_PERTAP=1 (experiment): thin layers run one MMA per filter tap (N = np, lean epilogue) instead of row-stacked
bool thin_pertap() {
  static const bool on = getenv("VF_THIN_PERTAP") && atoi(getenv("VF_THIN_PERTAP")) == 1;
  return on;
}

// Largest Cout that still takes the thin (pixels-on-M) tilings.  VF_THIN_MAX_COUT=32 sends the 64-channel layers to the wide
// tiling (channels on the 128 MMA rows, half of them zero): twice the MMA work, but the wide epilogue has no lane exchange.
int thin_max_cout() {
  static const int v = getenv("VF_THIN_MAX_COUT") ? atoi(getenv("VF_THIN_MAX_COUT")) : 64;
  return v;
}

bool plan_geometry(int layout, int bo_mode, int k, int kw, int kcl, int Cin, int Cout, int H, int W, int B, int passes, Geometry* out) {
  Geometry g;
  memset(&g, 0, sizeof(g));
  g.layout = layout; g.bo_mode = bo_mode;
  g.ch = layout == 2 ? 64 : 32;
  if (Cin % 8) return false;                            // 16-byte staging units; channels/cout beyond the real extent are zero-padded
  g.kc = g.ch / 8; g.ksteps = g.ch / 16; g.row_bytes = g.ch * 2; g.swz_mask = layout == 2 ? 7 : 3;
  g.half_bytes = MT * g.ch * 2; g.stage_bytes = 2 * g.half_bytes;
  g.H = H; g.W = W; g.k = k; g.pad = k / 2; g.kw = kw; g.padx = kw / 2; g.kcl = kcl; g.padc = kcl / 2;
  g.Wp = W + kw - 1; g.Cin = Cin; g.Cout = Cout;
  g.nchunk = (Cin + g.ch - 1) / g.ch; g.ntap = kcl * kcl; g.n_mt = (Cout + MT - 1) / MT; g.passes = passes;
  g.nchunk0 = g.nchunk;
  g.nst = k * kw;
  g.ksteps_last = (std::min(g.ch, Cin - (g.nchunk - 1) * g.ch) + 15) / 16;
  const int np_thin = (Cout + 15) / 16 * 16;
  if (!thin_pertap() && Cout <= thin_max_cout() && kw == k && k * np_thin <= 256 && 2 * kcl * kcl * np_thin <= 2304) {   // 2 bias tables behind the exchange scratch
    // ---- row-stacked thin path: pixels on M, the k taps of a filter row side by side on N (k*np columns) ----
    g.swap = 2;
    g.np = np_thin; g.ncols = k * g.np; g.ustride = 128 - (k - 1); g.nst = k;
    g.half_bytes = g.ncols * g.row_bytes; g.stage_bytes = 2 * g.half_bytes;
    g.n_mt = 1; g.G = 1;
    const int Vs = H * g.Wp;
    static const int stack_np = getenv("VF_STACK_NP") ? atoi(getenv("VF_STACK_NP")) : 16;   // stack the weight halves up to this np (0: never)
    // 3x3 only (runs on k_conv_mma<3>), and only with >= 2 channel chunks: masks1 (56 -> 7) is bound by MMA issue and gains
    // (110.7 -> 94.4 us), scratch1 (32 -> 3, one chunk, sigmoid + split-half stores) is bound by its epilogue and loses with the
    // smaller items / 3 epilogue groups (43.2 -> 56.7 us)
    g.stk = (passes == 3 && k == 3 && g.np <= stack_np && 2 * g.ncols <= 256 && g.nchunk >= 2) ? 1 : 0;
    const int ucols = g.ncols * (1 + g.stk);                             // TMEM columns of one unit
    const int umax = std::max(1, std::min(4, 256 / ucols));              // two accumulator sets in the 512 TMEM columns; <= 4 units
                                                                          // of one channel block = one work item per epilogue group
    bool found = false;
    for (int units = umax; units >= 1 && !found; --units) {
      g.npass = (Vs + g.ustride * units - 1) / (g.ustride * units);
      g.units = ((Vs + g.npass - 1) / g.npass + g.ustride - 1) / g.ustride;
      g.v_cnt = g.units * g.ustride;
      if (!box_rows(g)) return false;
      g.plane_bytes = (g.img_pix * g.row_bytes + 1023) / 1024 * 1024;
      for (g.nbuf = 2; g.nbuf >= 1; --g.nbuf) {
        const size_t act = (size_t)g.nbuf * 2 * g.plane_bytes + stg_bytes(g);  // + the epilogue's output tiles
        if (act + 2 * (size_t)g.stage_bytes + SMEM_SLACK > SMEM_LIMIT) continue;
        g.nstage = (int)std::min<size_t>(MAX_STAGE, (SMEM_LIMIT - SMEM_SLACK - act) / g.stage_bytes);
        if (g.nbuf == 2 || units == 1) found = true;
        break;
      }
    }
    if (!found) return false;
    g.nacc = (2 * g.units * ucols <= 512) ? 2 : 1;
    g.ngroups = B;
    g.nitems = g.ngroups * g.npass;
    *out = g;
    return true;
  }
  if (Cout <= thin_max_cout()) {
    // ---- swapped orientation (thin layers): pixels on M in 128-row units, np = Cout padded to 16 on N ----
    g.swap = 1;
    g.np = (Cout + 15) / 16 * 16; g.ncols = g.np; g.ustride = 128;
    g.half_bytes = g.np * g.row_bytes; g.stage_bytes = 2 * g.half_bytes;
    g.n_mt = 1; g.G = 1;
    const int Vs = H * g.Wp;
    const int umax = std::min(MAX_UNIT_B, 256 / g.np);
    for (int units = umax; units >= 1; --units) {
      g.npass = (Vs + 128 * units - 1) / (128 * units);
      g.units = ((Vs + g.npass - 1) / g.npass + 127) / 128;
      g.v_cnt = g.units * 128;
      if (!box_rows(g)) return false;
      g.plane_bytes = (g.img_pix * g.row_bytes + 1023) / 1024 * 1024;
      bool fit = false;
      for (g.nbuf = 2; g.nbuf >= 1; --g.nbuf) {
        const size_t act = (size_t)g.nbuf * 2 * g.plane_bytes;
        if (act + 2 * (size_t)g.stage_bytes + SMEM_SLACK > SMEM_LIMIT) continue;
        g.nstage = (int)std::min<size_t>(MAX_STAGE, (SMEM_LIMIT - SMEM_SLACK - act) / g.stage_bytes);
        fit = true;
        break;
      }
      if (fit && g.nbuf == 2) break;                      // largest unit count that still double-buffers the staging
      if (fit && units == 1) break;
      if (!fit && units == 1) return false;
    }
    g.nacc = (2 * g.units * g.np <= 512) ? 2 : 1;
    g.ngroups = B;
    g.nitems = g.ngroups * g.npass;
    *out = g;
    return true;
  }
  const int V = H * g.Wp;                               // virtual pixels per image (incl. k-1 wrap columns per row)
  // (G, v_cnt, npass): several whole small images per item, or an (almost) even slice of one large image.
  // v_cnt is rounded up to 32 columns; the overshoot reads zero-filled staging rows and is masked in the epilogue.
  g.We = g.Wp;
  static const bool rg_env = !(getenv("VF_ROWGROUPS") && atoi(getenv("VF_ROWGROUPS")) == 0);
  static const bool rg2_anyh = !(getenv("VF_RG2_ANYH") && atoi(getenv("VF_RG2_ANYH")) == 0);   // [1]: 8x16 / 12x16 maps too (48x64 inputs)
  static const bool rg3_env = !(getenv("VF_RG3") && atoi(getenv("VF_RG3")) == 0);   // [1]: column-strip items (row-group mode 3)
  // W == 16: strips only for layers that cannot run on CTA pairs (no whole 256-channel tiles) — twice as many, half as large
  // items fill the last round of the persistent grid better (enc2: 200 items on 148 CTAs -> 400)
  static const bool rg3_w16 = !(getenv("VF_RG3_W16") && atoi(getenv("VF_RG3_W16")) == 0);
  const bool strip16 = rg3_env && rg3_w16 && W == 16 && Cout % 256 != 0;
  if (rg_env && !strip16 && W == 16 && (H == 16 || (rg2_anyh && H >= 4 && H < 16 && H % 4 == 0)) && kw == k && layout == 1) {
    // ---- row-group mode 2 (see Geometry::rg) ----
    g.rg = 2; g.G = 1; g.npass = 1; g.We = 8;
    g.half_cols = H * 8; g.v_cnt = 2 * g.half_cols; g.col_stride = g.v_cnt; g.ncols_item = g.v_cnt;
    g.nseg = 2;
    g.seg_n[0] = g.seg_n[1] = g.half_cols;
    g.seg_off[0] = 0; g.seg_off[1] = g.half_cols;
    g.seg_px[0] = 0; g.seg_px[1] = 8;
    g.R = H + 2 * g.pad;
    g.img_pix = (g.R * g.Wp + 7) / 8 * 8;
    g.box_bytes = g.R * g.Wp * g.row_bytes;
    g.plane_bytes = ((g.img_pix + 16) * g.row_bytes + 1023) / 1024 * 1024;    // + the right group's reach past the last row
    bool ok_rg = false;
    for (g.nbuf = 2; g.nbuf >= 1; --g.nbuf) {
      const size_t act = (size_t)g.nbuf * 2 * g.plane_bytes;
      if (act + 2 * (size_t)g.stage_bytes + SMEM_SLACK > SMEM_LIMIT) continue;
      g.nstage = (int)std::min<size_t>(MAX_STAGE, (SMEM_LIMIT - SMEM_SLACK - act) / g.stage_bytes);
      ok_rg = true;
      break;
    }
    if (ok_rg) {
      g.nacc = (2 * g.ncols_item <= 512) ? 2 : 1;
      g.ngroups = B;
      g.nitems = g.n_mt * g.ngroups;
      *out = g;
      return true;
    }
    g.rg = 0; g.We = g.Wp;
  }
  if (rg_env && rg3_env && (W > 16 || strip16) && W % 8 == 0 && H * 8 <= 256 && H % 4 == 0 && kw == k && layout == 1) {
    // ---- row-group mode 3 (see Geometry::rg): item = one 8-pixel column group of one image, all rows ----
    Geometry t = g;
    t.rg = 3; t.G = 1; t.npass = W / 8; t.We = 8;
    t.Wp = 8 + kw - 1;                                   // the staged strip: 8 pixels + halo columns
    t.half_cols = H * 8; t.v_cnt = H * 8; t.col_stride = t.v_cnt; t.ncols_item = t.v_cnt;
    t.nseg = 1; t.seg_n[0] = t.v_cnt; t.seg_off[0] = 0; t.seg_px[0] = 0;
    t.R = H + 2 * t.pad;
    t.img_pix = (t.R * t.Wp + 7) / 8 * 8;
    t.box_bytes = t.R * t.Wp * t.row_bytes;
    t.plane_bytes = ((t.img_pix + 16) * t.row_bytes + 1023) / 1024 * 1024;   // + the last tap's reach past the last row
    bool ok_rg = false;
    for (t.nbuf = 2; t.nbuf >= 1; --t.nbuf) {
      const size_t act = (size_t)t.nbuf * 2 * t.plane_bytes;
      if (act + 2 * (size_t)t.stage_bytes + SMEM_SLACK > SMEM_LIMIT) continue;
      t.nstage = (int)std::min<size_t>(MAX_STAGE, (SMEM_LIMIT - SMEM_SLACK - act) / t.stage_bytes);
      ok_rg = true;
      break;
    }
    if (ok_rg && t.R <= 256) {
      t.nacc = (2 * t.ncols_item <= 512) ? 2 : 1;
      t.ngroups = B;
      t.nitems = t.n_mt * t.ngroups * t.npass;
      *out = t;
      return true;
    }
  }
  if (rg_env && W == 8 && kw == k && layout == 1 && H >= 2 && H * 8 <= 128) {
    // ---- row-group mode (see Geometry::rg) ----
    const int pitch = H + g.pad;
    int G = 1;
    for (int c = 2; c <= 4; ++c) {
      const int rows = (c - 1) * pitch + H;
      if (rows * 8 <= 256 && rows % 2 == 0) G = c;
    }
    if ((H % 2) == 0 || G > 1) {
      const int rows = (G - 1) * pitch + H;
      if (rows % 2 == 0) {
        g.rg = 1; g.G = G; g.npass = 1; g.We = W;
        g.v_cnt = (H * 8 + 31) / 32 * 32;
        g.col_stride = pitch * 8; g.ncols_item = rows * 8;
        g.nseg = 1; g.seg_n[0] = g.ncols_item; g.seg_off[0] = 0; g.seg_px[0] = 0;
        g.R = H + 2 * g.pad;
        g.img_pix = pitch * g.Wp;
        g.box_bytes = g.R * g.Wp * g.row_bytes;
        g.plane_bytes = (((G - 1) * g.img_pix + g.R * g.Wp) * g.row_bytes + 1023) / 1024 * 1024;
        bool ok_rg = false;
        for (g.nbuf = 2; g.nbuf >= 1; --g.nbuf) {
          const size_t act = (size_t)g.nbuf * 2 * g.plane_bytes;
          if (act + 2 * (size_t)g.stage_bytes + SMEM_SLACK > SMEM_LIMIT) continue;
          g.nstage = (int)std::min<size_t>(MAX_STAGE, (SMEM_LIMIT - SMEM_SLACK - act) / g.stage_bytes);
          ok_rg = true;
          break;
        }
        if (ok_rg) {
          g.nacc = (2 * g.ncols_item <= 512) ? 2 : 1;
          g.ngroups = (B + g.G - 1) / g.G;
          g.nitems = g.n_mt * g.ngroups * g.npass;
          *out = g;
          return true;
        }
        g.rg = 0; g.We = g.Wp;
      }
    }
  }
  if (V <= 128) {
    g.v_cnt = (V + 31) / 32 * 32;
    g.G = std::max(1, std::min(256 / g.v_cnt, 3));
    g.npass = 1;
  } else {
    g.G = 1;                                            // <= 256 columns per item: two TMEM accumulator sets
    const int np0 = (V + 255) / 256;
    int best_waste = 1 << 30;
    for (int np = np0; np <= np0 + 2; ++np) {           // least padded columns, then fewest passes
      const int v = ((V + np - 1) / np + 31) / 32 * 32;
      if (v > 256) continue;
      if (np * v - V < best_waste) { best_waste = np * v - V; g.npass = np; g.v_cnt = v; }
    }
  }
  // MMA column segments: n <= 256, multiple of 16
  g.nseg = 0;
  int left = g.v_cnt, off = 0;
  const int nsplit = (g.v_cnt + 255) / 256;
  const int base = ((g.v_cnt / nsplit) + 15) / 16 * 16;
  while (left > 0) {
    const int n = std::min(base, left);
    if (n % 16 || g.nseg >= MAX_SEG) return false;
    g.seg_n[g.nseg] = n; g.seg_off[g.nseg] = off; g.seg_px[g.nseg] = off; ++g.nseg;
    off += n; left -= n;
  }
  if (!box_rows(g)) return false;
  // buffers: prefer double-buffered activations + as many weight stages as fit (>= 2)
  bool ok = false;
  for (; g.G >= 1 && !ok; --g.G) {
    g.plane_bytes = (g.G * g.img_pix * g.row_bytes + 1023) / 1024 * 1024;
    for (g.nbuf = 2; g.nbuf >= 1; --g.nbuf) {
      const size_t act = (size_t)g.nbuf * 2 * g.plane_bytes;
      if (act + 2 * (size_t)g.stage_bytes + SMEM_SLACK > SMEM_LIMIT) continue;
      g.nstage = (int)std::min<size_t>(MAX_STAGE, (SMEM_LIMIT - SMEM_SLACK - act) / g.stage_bytes);
      ok = true;
      break;
    }
    if (ok) break;
  }
  if (!ok || g.G * g.nseg > MAX_UNIT) return false;
  g.nacc = (2 * g.G * g.v_cnt <= 512) ? 2 : 1;
  g.ngroups = (B + g.G - 1) / g.G;
  g.nitems = g.n_mt * g.ngroups * g.npass;
  *out = g;
  return true;
}

// >= 116 KB so that exactly one CTA is resident per SM: every CTA allocates all 512 TMEM columns
size_t smem_bytes(const Geometry& g) {
  return std::max((size_t)g.nbuf * 2 * g.plane_bytes + (size_t)g.nstage * g.stage_bytes + SMEM_SLACK + stg_bytes(g),
                  (size_t)116 * 1024);
}

}  // namespace

bool mma_conv_supported(int k, int cin, int cout, int H, int W) {
  if (k != 3 && k != 5) return false;
  int layout, bo;
  current_mode(&layout, &bo);
  Geometry g;
  return plan_geometry(layout, bo, k, k, k, cin, cout, H, W, 1, 3, &g);
}

// Row-group mode 3: split the strips of the last, partial round of the persistent grid in two half strips when that lets the
// round finish in about half the time (800 strips on 148 CTAs: 5 whole rounds + 60 strips -> 120 half strips, 5.6 instead of 6
// rounds).  Only when a half strip keeps the epilogue's chunk -> warp assignment (4*H % 64 == 0), so that the fused statistics
// are bit-identical with and without the split; not with producer-side finalisation (its arrival counts assume whole items).
void apply_tail_split(Geometry& g, int sms, bool fused_finalize) {
  static const bool on = !(getenv("VF_TAIL_SPLIT") && atoi(getenv("VF_TAIL_SPLIT")) == 0);
  g.tsplit = 0; g.n_full = g.nitems;
  if (!on || g.rg != 3 || fused_finalize || ((g.v_cnt >> 1) % 64) != 0 || sms <= 0) return;
  const int grid = std::min(g.nitems, sms);
  const int n_full = (g.nitems / grid) * grid, tail = g.nitems - n_full;
  if (tail <= 0 || 2 * tail > grid) return;
  g.tsplit = 1; g.n_full = n_full; g.nitems = n_full + 2 * tail;
}

// partial slots per (sample, channel) the fused instance-norm statistics of this layer shape use (0: not fused); sizes the
// engine's partial-sum scratch
int mma_conv_stats_slots(int k, int kw, int cin, int cout, int H, int W) {
  int layout, bo;
  current_mode(&layout, &bo);
  if (layout == 2 && !(cin % 64 == 0 && cout >= 128)) layout = 1;
  Geometry g;
  if (!plan_geometry(layout, bo, k, kw, k, cin, cout, H, W, 1, 3, &g)) return 0;
  return g.swap == 2 ? g.npass : (g.swap ? 0 : (g.rg == 3 ? 4 : 2) * g.npass);
}

// Host-only description of the tiling plan_geometry() picks for a layer (no device needed): lets the CPU test suite check the
// resource invariants (shared memory, TMEM columns, MMA N, box coverage) for every layer shape of the supported specs.
bool mma_conv_describe(int k, int kw, int cin, int cout, int H, int W, int B, int passes, int out[24]) {
  int layout, bo;
  current_mode(&layout, &bo);
  if (layout == 2 && !(cin % 64 == 0 && cout >= 128)) layout = 1;
  Geometry g;
  if (!plan_geometry(layout, bo, k, kw, k, cin, cout, H, W, B, passes == 3 ? 3 : 1, &g)) return false;
  apply_tail_split(g, 148, false);
  const int acc_cols = g.swap ? g.units * g.ncols * (1 + g.stk) : (g.rg ? g.ncols_item : g.G * g.v_cnt);
  int nmax = 0;                                       // largest MMA N
  if (g.swap) nmax = g.ncols * (1 + g.stk); else for (int i = 0; i < g.nseg; ++i) nmax = std::max(nmax, g.seg_n[i]);
  // last shared-memory pixel row any MMA of an item reads (relative to the item's plane), and the rows the plane holds
  int last_read, plane_rows = g.plane_bytes / g.row_bytes;
  const int tap_reach = (g.k - 1) * g.Wp + (g.kw - 1);
  int off = 0;                                        // largest offset of an item's first virtual pixel inside its first box row
  for (int ps = 0; ps < g.npass; ++ps) off = std::max(off, (ps * g.v_cnt) % g.Wp);
  if (g.swap == 2) last_read = off + (g.units - 1) * g.ustride + 127 + (g.k - 1) * g.Wp;
  else if (g.swap) last_read = off + g.units * 128 - 1 + tap_reach;
  else if (g.rg == 1) last_read = ((g.G - 1) * (g.H + g.pad) + g.H - 1) * g.Wp + 7 + tap_reach;
  else if (g.rg == 2) last_read = (g.H - 1) * g.Wp + 8 + 7 + tap_reach;
  else if (g.rg == 3) last_read = (g.H - 1) * g.Wp + 7 + tap_reach;
  else last_read = (g.G - 1) * g.img_pix + off + g.v_cnt - 1 + tap_reach;
  const int v[24] = {g.swap, g.rg, g.G, g.npass, g.v_cnt, g.units, g.ncols, nmax, acc_cols, g.nacc, g.nbuf, g.nstage,
                     g.stage_bytes, g.plane_bytes, (int)smem_bytes(g), g.nitems, g.R, g.Wp, g.img_pix, g.box_bytes,
                     last_read, plane_rows, g.nchunk, g.n_mt};
  for (int i = 0; i < 24; ++i) out[i] = v[i];
  return true;
}

int mma_conv_prepare_weights(const float* w_sp, int k, int kw, int kcl, int cin, int cout, MmaConvWeights* out,
                             std::vector<void*>* allocs, std::string* err) {
  int layout, bo;
  current_mode(&layout, &bo);
  if (layout == 2 && !(cin % 64 == 0 && cout >= 128)) layout = 1;      // SWIZZLE_128B only for the wide gate convolutions
  const int ch = layout == 2 ? 64 : 32;
  if (cin % 8) { if (err) *err = "cin % 8"; return -1; }
  const bool swap = cout <= thin_max_cout();
  const int np = (cout + 15) / 16 * 16;
  const bool stacked = !thin_pertap() && swap && kw == k && k * np <= 256 && 2 * kcl * kcl * np <= 2304;   // one stage = a filter ROW: rows = (dx, output channel); must agree with plan_geometry
  const int rows = stacked ? k * np : (swap ? np : MT);               // operand tile rows
  const int kk = stacked ? k : k * kw;                                // stages per channel chunk
  const int nchunk = (cin + ch - 1) / ch, n_mt = swap ? 1 : (cout + MT - 1) / MT, kc = ch / 8, rb = ch * 2;
  const int swz = layout == 2 ? 7 : 3;
  float amax = 0.f;
  for (size_t i = 0; i < (size_t)k * kw * cin * cout; ++i) amax = std::max(amax, fabsf(w_sp[i]));
  int sl = 0;
  if (amax > 0.f) sl = (int)floorf(log2f(16384.0f / amax));           // max |w| * 2^sl in [8192, 16384]
  sl = std::max(-24, std::min(sl, 24));
  const float scale = ldexpf(1.0f, sl);
  const size_t half_elems = (size_t)rows * ch;
  const size_t total = (size_t)n_mt * nchunk * kk * 2 * half_elems;
  std::vector<__half> packed(total);
  for (int mt = 0; mt < n_mt; ++mt)
    for (int c = 0; c < nchunk; ++c)
      for (int t = 0; t < kk; ++t) {
        const size_t blk = (((size_t)mt * nchunk + c) * kk + t) * 2 * half_elems;
        for (int kc8 = 0; kc8 < kc; ++kc8)
          for (int r = 0; r < rows; ++r)
            for (int e = 0; e < 8; ++e) {
              const int ci = c * ch + kc8 * 8 + e;
              const int n = stacked ? r % np : mt * rows + r;            // output channel
              const int tap = stacked ? t * k + r / np : t;              // stacked: stage t = filter row dy, r / np = dx
              const float v = (ci < cin && n < cout) ? w_sp[((size_t)tap * cin + ci) * cout + n] * scale : 0.f;
              const __half h = __float2half_rn(v);
              const __half l = __float2half_rn(v - __half2float(h));
              size_t pos;                                              // element index inside the 128-row operand tile
              pos = ((size_t)r * rb + (size_t)((kc8 ^ ((r * rb >> 7) & swz)) << 4)) / 2 + e;
              packed[blk + pos] = h;
              packed[blk + half_elems + pos] = l;
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
  out->k = k; out->kw = kw; out->kcl = kcl; out->cin = cin; out->cout = cout; out->scale_log2 = sl; out->ready = true;
  return 0;
}

// ---- tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link dependency) -------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct MapKey {
  const void* base;
  long long sample_stride, lo_off;
  int C, pix_stride, H, W, B, ch, Wp, R, layout;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapEntry { MapKey key; CUtensorMap map; };
std::vector<MapEntry> g_maps;      // a handful of (buffer, layer geometry) pairs per engine; engines are not thread-safe

// 5-D view of a split-half activation tensor: (channel, x, y, sample, plane); box = ch x Wp x R x 1 x 1
int activation_map(const View& v, const Geometry& g, int B, CUtensorMap* out) {
  if (!v.lo_off || !v.p) return -10;                                  // the tensor-core path needs split-half storage
  if ((v.pix_stride % 8) || (v.ch_off % 8) || (v.sample_stride % 8) || (v.lo_off % 8) || (v.C % 8)) return -11;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = reinterpret_cast<const __half*>(v.p) + v.ch_off;
  key.sample_stride = v.sample_stride; key.lo_off = v.lo_off; key.C = v.C; key.pix_stride = v.pix_stride;
  key.H = g.H; key.W = g.W; key.B = B; key.ch = g.ch; key.Wp = g.Wp; key.R = g.R; key.layout = g.layout;
  for (const MapEntry& e : g_maps)
    if (e.key == key) { *out = e.map; return 0; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return -12;
  const cuuint64_t gdim[5] = {(cuuint64_t)v.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)B, 2};
  const cuuint64_t gstride[4] = {(cuuint64_t)v.pix_stride * 2, (cuuint64_t)g.W * v.pix_stride * 2,
                                 (cuuint64_t)(B > 1 ? v.sample_stride : (long long)g.H * g.W * v.pix_stride) * 2, (cuuint64_t)v.lo_off * 2};
  const cuuint32_t box[5] = {(cuuint32_t)g.ch, (cuuint32_t)g.Wp, (cuuint32_t)g.R, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  MapEntry e;
  e.key = key;
  const CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(key.base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, g.layout == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -13;
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.push_back(e);
  *out = e.map;
  return 0;
}

int mma_conv_launch(const MmaConvWeights& w, const MmaConvCall& c, int B, cudaStream_t s) {
  Params P;
  int layout, bo;
  current_mode(&layout, &bo);
  if (layout == 2 && !(w.cin % 64 == 0 && w.cout >= 128)) layout = 1;   // must agree with mma_conv_prepare_weights
  if (c.src.C + c.src1.C != w.cin) return -5;
  if (!plan_geometry(layout, bo, w.k, w.kw, w.kcl, w.cin, w.cout, c.H, c.W, B, c.passes == 3 ? 3 : 1, &P.g)) return -1;
  // CTA-pair kernel (cta_group::2) for the 16x16 wide layers with whole 256-channel tiles (k_conv_pair)
  static const bool pair_env = !(getenv("VF_CTA_PAIR") && atoi(getenv("VF_CTA_PAIR")) == 0);      // default on; VF_CTA_PAIR=0 = single-CTA kernel
  const bool pair = pair_env && P.g.rg == 2 && P.g.layout == 1 && w.cout % 256 == 0 && c.src1.C == 0 && !c.out.lo_off &&
                    c.act == ACT_NONE && c.passes == 3;
  if (pair) {                                             // every CTA stages ONE 8-pixel column group of all rows (+ halo)
    Geometry& g = P.g;
    g.Wp = 8 + g.kw - 1;
    g.R = g.H + 2 * g.pad;
    g.box_bytes = g.R * g.Wp * g.row_bytes;
    g.img_pix = (g.R * g.Wp + 7) / 8 * 8;
    g.plane_bytes = ((g.img_pix + 16) * g.row_bytes + 1023) / 1024 * 1024;
    g.nbuf = 2;
    const size_t fixed = 1024 + 2 * 25 * MT * 4 + 512 + (size_t)2 * 2 * g.plane_bytes;
    g.nstage = (int)std::min<size_t>(MAX_STAGE, (SMEM_LIMIT - fixed) / g.stage_bytes);
    if (g.nstage < 2) return -8;
  }
  P.g.out_scale = ldexpf(1.0f, -w.scale_log2);
  P.g.m_Wp = fd_magic(P.g.Wp); P.g.m_npass = fd_magic(P.g.npass); P.g.m_np = fd_magic(P.g.np > 0 ? P.g.np : 1);
  static const int dual_env = getenv("VF_DUAL_ISSUE") ? atoi(getenv("VF_DUAL_ISSUE")) : 1;
  P.g.dual = dual_env;
  P.out = c.out; P.sabias = c.sabias; P.bias = c.bias; P.w = w.w_hi; P.B = B; P.act = c.act;
  P.stats_partial = nullptr; P.stats_S = 0;
  P.stats_fin = nullptr; P.stats_cnt = nullptr; P.stats_npix = c.H * c.W; P.stats_eps = c.stats_eps;
  if (c.stats_partial && !P.g.swap) { P.stats_partial = c.stats_partial; P.stats_S = P.g.npass * (P.g.rg == 3 ? 4 : 2); }

  if (!P.g.swap && c.out.lo_off) return -6;                           // the wide epilogue writes float32 (pre-norm) outputs
  int r = activation_map(c.src, P.g, B, &P.tmap0);
  if (r) return r;
  P.tmap1 = P.tmap0;
  if (c.src1.C > 0) {
    if (c.src.C % P.g.ch) return -7;                                  // the second source starts on a chunk boundary
    P.g.nchunk0 = c.src.C / P.g.ch;
    if ((r = activation_map(c.src1, P.g, B, &P.tmap1))) return r;
  }
  const size_t smem = smem_bytes(P.g);
  static bool attr_set = false;
  static int epi_groups = 4;          // epilogue warp groups of the 3x3 row-stacked layers (VF_EPI_GROUPS=2|3|4: A/B switch)
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_conv_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT) != cudaSuccess) return -3;
    if (cudaFuncSetAttribute(k_conv_mma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT) != cudaSuccess) return -3;
    if (cudaFuncSetAttribute(k_conv_mma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT) != cudaSuccess) return -3;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    const char* e = getenv("VF_EPI_GROUPS");
    if (e && atoi(e) >= 2 && atoi(e) <= 4) epi_groups = atoi(e);
    attr_set = true;
  }
  apply_tail_split(P.g, g_num_sms > 0 ? g_num_sms : 148, c.stats_fin != nullptr && c.stats_cnt != nullptr);
  const int grid = std::min(P.g.nitems, g_num_sms > 0 ? g_num_sms : 148);
  {
    const int ng = P.g.stk ? 3 : ((P.g.swap == 2 && P.g.k == 3) ? epi_groups : 2), ncb = P.g.np >> 4;
    if (c.stats_partial && P.g.swap == 2 && thin_epilogue_tiled(P.g, c.out, c.act) && (ng == 2 || ng % ncb == 0)) {
      P.stats_partial = c.stats_partial;
      P.stats_S = P.g.npass;
    }
  }
  if (c.stats_slots) *c.stats_slots = P.stats_S;
  if (pair) {
    // the packed weights as a 2-D tensor (rows of 512 bytes; one 16 KB stage = 32 rows) for the pair's tensor-TMA weight loads
    {
      struct WMap { const void* p; cuuint64_t rows; CUtensorMap m; };
      static std::vector<WMap> wmaps;                       // keyed by (buffer, extent): a freed buffer's address may be reused
      const cuuint64_t rows = (cuuint64_t)P.g.n_mt * P.g.nchunk * P.g.nst * (P.g.stage_bytes / 512);
      bool found = false;
      for (auto& e : wmaps) if (e.p == (const void*)w.w_hi && e.rows == rows) { P.tmap1 = e.m; found = true; }
      if (!found) {
        EncodeTiledFn enc = encode_fn();
        if (!enc) return -12;
        const cuuint64_t gdim[2] = {256, rows};
        const cuuint64_t gstride[1] = {512};
        const cuuint32_t box[2] = {256, (cuuint32_t)(P.g.stage_bytes / 512)};
        const cuuint32_t estr[2] = {1, 1};
        CUtensorMap m;
        if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(w.w_hi), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -13;
        if (wmaps.size() > 256) wmaps.clear();
        wmaps.push_back(WMap{(const void*)w.w_hi, rows, m});
        P.tmap1 = m;
      }
    }
    static bool pair_attr = false;
    if (!pair_attr) {
      if (cudaFuncSetAttribute(k_conv_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT) != cudaSuccess) return -3;
      pair_attr = true;
    }
    if (c.stats_finalized) *c.stats_finalized = false;
    const int nitems_pair = (w.cout / 256) * B;
    const int pairs = std::min(nitems_pair, (g_num_sms > 0 ? g_num_sms : 148) / 2);
    const size_t psmem = 1024 + 2 * 25 * MT * 4 + 512 + (size_t)2 * 2 * P.g.plane_bytes + (size_t)P.g.nstage * P.g.stage_bytes;
    ++g_launch_counter;
    return launch_k(k_conv_pair, dim3(2 * pairs), dim3(PAIR_THREADS), psmem, s, P) == cudaSuccess ? 0 : -4;
  }
  if (P.stats_partial && c.stats_fin && c.stats_cnt && (P.g.n_mt * 4 <= VF_STAT_CNT_STRIDE)) { P.stats_fin = c.stats_fin; P.stats_cnt = c.stats_cnt; }
  if (c.stats_finalized) *c.stats_finalized = P.stats_fin != nullptr;
  ++g_launch_counter;
  if (P.g.stk)                                                          // stacked weight halves: the 128-register instance (3 epilogue groups)
    return launch_k(k_conv_mma<3>, dim3(grid), dim3(128 + 128 * 3), smem, s, P) == cudaSuccess ? 0 : -4;
  if (epi_groups == 4 && P.g.swap == 2 && P.g.k == 3)
    return launch_k(k_conv_mma<4>, dim3(grid), dim3(128 + 128 * 4), smem, s, P) == cudaSuccess ? 0 : -4;
  if (epi_groups == 3 && P.g.swap == 2 && P.g.k == 3)
    return launch_k(k_conv_mma<3>, dim3(grid), dim3(128 + 128 * 3), smem, s, P) == cudaSuccess ? 0 : -4;
  return launch_k(k_conv_mma<2>, dim3(grid), dim3(NTHREADS), smem, s, P) == cudaSuccess ? 0 : -4;
}

}  // namespace vf

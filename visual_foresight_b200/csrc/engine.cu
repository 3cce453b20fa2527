// engine.cu — the vfengine handle: weight preprocessing, the conv-LSTM/CDNA rollout schedule, costs,
// the device-resident CEM loop, and the extern "C" boundary declared in include/vfengine.h.
//
// Data layout in HBM (per handle, B = max_samples):
//   ctx_frames  f32 [C][ncam][H][W][3]        context, shared by all samples (sample_stride 0 views)
//   gen_images  f32 [B][P][ncam][H][W][3]     every predicted frame, reference output layout
//   gen_distrib f32 [B][P][ncam][H][W][nd]    (vpred_model_interface.py:75-88)
//   per view / rnn layer: lstm_in [B][h][w][2F]  = [x | h_prev] conv input, c [B][h][w][F]
//   raw         f32 [B][H*W*Cout max]         conv output scratch (pre-norm)
//   mask_h      [B][H][W][ngf]                h_masks (first source of the mask-logit conv)
//   layers      [B][H][W][cl = 3*(nt+3) -> 8]  T_0..T_nt-1 | prev | first | scratch | 0   (second source, and the composite's layers)
// On the tensor-core path every convolution INPUT buffer (lstm_in, pack0, dec_in, act_*, scr_h, mask_h, layers) is kept
// in split-half storage (two fp16 planes, vf_common.cuh) so the convolution can fetch it with TMA; conv OUTPUTS (raw,
// logits), the cell state and the predicted frames stay float32.
#include <math.h>
#include <stdarg.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/vfengine.h"
#include "vf_common.cuh"
#include "conv_mma.cuh"

using namespace vf;

namespace {

// exchange window layout: [VF_MAX_WORLD x u32 arrival counters][u32 status] padded to 256 bytes, then the f64 score matrix
constexpr size_t COMM_HEADER_BYTES = 256;
inline double* comm_scores(void* window) { return reinterpret_cast<double*>(static_cast<char*>(window) + COMM_HEADER_BYTES); }
inline unsigned* comm_status_word(void* window) { return reinterpret_cast<unsigned*>(window) + VF_MAX_WORLD; }

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct ConvLayer {
  std::string name;
  int k = 0, H = 0, W = 0, cin_sp = 0, cin_const = 0, cout = 0;
  int cin_w = 0;            // spatial channels present in the weight tensor (cin_sp may be zero-padded beyond it)
  bool lstm = false;
  bool fold = false;        // tensor-core path: the k dx taps are folded into the input channels by the producer (k x 1 conv over
                            // k*cin_sp channels) — used for the first encoder conv, whose 8 input channels would waste the K=16 MMAs
  bool s2d = false;         // tensor-core path, first encoder conv: conv k x k + 2x2 average pool evaluated as ONE 3x3 conv over the
                            // space-to-depth input (2x2 pixel blocks -> 4*cin_sp channels) with pre-averaged weights; the conv
                            // output is already pooled (see prepare_conv)
  int kcls = 0;             // kernel size of the border classes of sabias (0 = k; 3 for s2d)
  float* w_sp = nullptr;    // [k*k][cin_sp][cout]
  float* wcls = nullptr;    // [k*k][A][cout]
  float* bias = nullptr;
  float *gamma = nullptr, *beta = nullptr;          // conv norm  (or gates norm for lstm)
  float *cgamma = nullptr, *cbeta = nullptr;        // cell norm (lstm)
  float* sabias = nullptr;  // [B][k*k][cout], or [S-1][B][k*k][cout] when the engine hoists it out of the step loop
  long long sab_step = 0;   // elements between consecutive cell steps (hoisted form)
  MmaConvWeights mma;       // tensor-core operand copies (precision != SIMT)
};

struct RnnState {
  float* h_s2d = nullptr;     // [B][h/2][w/2][4F]: space-to-depth copy of h for the next (pool-fused) encoder conv, or null
  float* snap_c = nullptr;    // [h][w][F]    sample 0's cell state after the shared-prefix steps (per-plan prefix cache)
  float* snap_in = nullptr;   // [h][w][2F]   sample 0's lstm_in row (split-half: hi row then lo row, same bytes)
  float* lstm_in = nullptr;   // [B][h][w][2F]
  float* c = nullptr;         // [B][h][w][F]
  int h = 0, w = 0, F = 0;
};

struct ViewNet {
  std::vector<ConvLayer> enc_conv, dec_conv;
  std::vector<ConvLayer> enc_lstm, dec_lstm;     // entries with k == 0 for non-rnn layers
  std::vector<RnnState> enc_rnn, dec_rnn;
  ConvLayer scratch0, scratch1, masks0, masks1;
  ConvLayer heads0;            // scratch.conv0 and masks.conv0 as ONE 3x3 conv ngf -> 2*ngf over the last decoder map (opt.merge_heads)
  float *cdna_w = nullptr, *cdna_b = nullptr;
  float *w_state = nullptr, *b_state = nullptr;   // per-view state head (IndepMultiSAVP: independent weight sets)
  float *w_z = nullptr, *b_z = nullptr;           // dense LSTM over the latent (use_rnn_z)
  float* zstate = nullptr;                        // [B][2*nz] its (c, h)
  float* state_cur = nullptr;                      // [B][sdim] this view's predicted state
};

struct DebugEntry {
  View v;
  int H, W;
};

}  // namespace

struct vf_engine {
  vf_config cfg;
  char err[512];
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t side = nullptr;           // second stream of the CDNA-head branch (opt.side_cdna)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<void*> allocs;
  std::map<std::string, HostTensor> host_w;
  bool weights_ready = false, context_set = false, distrib_set = false, predicted = false;
  int last_M = 0;

  // derived
  int B, H, W, ncam, nd, adim, sdim, nz, A, S, C, P, ngf, nt, kc, nm, n_enc, cm;
  std::vector<ViewNet> views;

  // shared scratch
  float *raw = nullptr, *dec_in = nullptr, *stats = nullptr, *cstats = nullptr;
  double *stats_partial = nullptr, *cstats_partial = nullptr;
  int* stat_cnt = nullptr;      // [B][VF_STAT_CNT_STRIDE] arrival counters of the fused statistics finalisation (zero between kernels)
  std::vector<float*> act_enc, act_dec;
  float* pack0 = nullptr;    // [B][H][W][8] packed (image, first) input of enc0 (tensor-core path)
  float *scr_h = nullptr, *mask_h = nullptr, *layers = nullptr, *logits = nullptr, *kern = nullptr, *partial = nullptr;
  float* cdna_part = nullptr;   // split-K partial products of the CDNA dense head
  int nblk = 0, cl = 0;
  bool split = false;           // conv inputs in split-half storage (tensor-core path)
  int conv_error = 0;           // first failed tensor-core conv launch (reported by rollout)
  char conv_error_layer[64] = {0};

  // context / outputs
  uint8_t* ctx_u8 = nullptr;
  float *ctx_frames = nullptr, *ctx_distrib = nullptr, *ctx_states = nullptr, *ctx_actions = nullptr;
  int n_ctx_actions = 0;
  int* desig_pix_dev = nullptr;
  float *gen_images = nullptr, *gen_distrib = nullptr, *gen_states = nullptr;
  float *sa = nullptr, *state_cur = nullptr, *zs = nullptr;
  float* sa_all = nullptr;    // [ncam][S-1][B][A]: every step's tiled action/state vector (opt.hoist_sa)
  float* actions = nullptr;   // [B][Tcap][adim]
  int Tcap = 0, T = 0;
  float* cost = nullptr;      // [B][P][ntask]
  double *goal_dev = nullptr, *taskw_dev = nullptr;
  float* goal_img = nullptr;
  double* scores_tmp = nullptr;   // [B]
  int* fetch_idx = nullptr;
  float* fetch_buf = nullptr;

  // CEM
  vf_cem_params cem;
  bool cem_active = false;
  int cem_D = 0, cem_T = 0, cem_Dmax = 0, cem_act_cap = 0;
  double *cem_mean = nullptr, *cem_factor = nullptr, *cem_std0 = nullptr, *cem_cov = nullptr;
  double *cem_scores = nullptr;     // [iters][Mg] (engine-owned or bound by vf_cem_bind_scores)
  double *cem_scores_own = nullptr, *cem_scores_ext = nullptr;
  size_t cem_scores_cap = 0;
  float* cem_noise = nullptr;
  size_t cem_noise_cap = 0;
  bool cem_has_noise = false;
  int* cem_elite_idx = nullptr;
  double *cem_elites_nr = nullptr, *cem_best64 = nullptr, *cem_local_nr = nullptr, *cem_actions64 = nullptr;
  int cem_Kf = 1;                   // futures per action sequence (stochastic planning)
  int* cem_fut_idx = nullptr;       // [B] global action index of every rollout sample
  double* cem_fut_scores = nullptr; // [B] per-rollout scores before the mean over futures
  double* topk_keys = nullptr;
  int* topk_idx = nullptr;
  size_t topk_cap = 0;
  int cem_Mg_cap = 0;               // capacity (global samples) of the elite buffers
  float* cem_goal_host_copy = nullptr;

  // peer-memory score exchange (vf_comm_*): this handle's window = [VF_MAX_WORLD arrival counters | status | score matrix]
  struct Comm {
    void* window = nullptr;           // cudaMalloc'd (IPC-exportable)
    size_t window_bytes = 0;
    int cap_iters = 0, cap_global = 0;
    bool connected = false;
    int rank = 0, world = 1;
    void* peer_base[VF_MAX_WORLD] = {nullptr};   // mapped windows in rank order (own entry = window)
    bool peer_ipc[VF_MAX_WORLD] = {false};       // opened with cudaIpcOpenMemHandle (to be closed)
    unsigned epoch = 0;
    unsigned long long timeout_ns = 30000000000ull;
  } comm;
  double* cem_scores_final = nullptr;           // while connected: private copy of the rows vf_cem_iter_select consumed

  std::map<std::string, DebugEntry> debug[4];

  // CUDA graph of one rollout (S-1 cell steps): captured on the second call with a given key, replayed afterwards
  struct GraphSlot { cudaGraphExec_t exec = nullptr; long long launches = 0; bool seen = false; };
  std::map<long long, GraphSlot> graphs;      // rollout graphs by (prefix mode, shared steps, M, T, n_ctx_actions)
  long long ctx_version = 0;                  // bumped by everything the shared-prefix state depends on (context, weights)
  long long snap_version = -1;                // ctx_version the prefix snapshot was taken at (-1: none)
  bool use_graph = true;
  // A/B switches, read ONCE in vf_create (VF_* environment variables; defaults in brackets)
  struct Opts {
    bool shared_prefix = true;   // VF_SHARED_PREFIX [1]: context-only cell steps run once on one sample
    bool prefix_cache = true;    // VF_PREFIX_CACHE [1]: CEM iterations 1.. restore the prefix state saved by iteration 0
    bool stats_fin = true;       // VF_STATS_FIN [1]: separate k_stats_finalize launches (1) or consumers finalise on the fly (0; measured 4 % slower)
    bool epi_stats = true;       // VF_EPI_STATS [1]: instance-norm statistics of the thin convolutions come from their epilogue
    bool merge_heads = true;     // VF_MERGE_HEADS [1]: scratch.conv0 + masks.conv0 as one convolution
    bool side_cdna = false;      // VF_SIDE_CDNA [0] (measured +1.4 ms: the side kernels delay the decoder's launches more than they hide): the CDNA head (dense -> kernels -> apply) runs on a second stream beside the decoder
                                 //   (fork after the last encoder conv-LSTM, join before the scratch-image conv; a branch of the graph)
    bool lstm_fused = false;     // VF_LSTM_FUSED [0]: conv-LSTM pointwise (both instance norms) as one cluster kernel per layer
                                 //   (measured SLOWER: 130 us vs 26 + 18 + 4 us at 32x32x32 — one 214-register CTA per SM, phases serialised)
    bool hoist_sa = true;        // VF_HOIST_SA [1]: the action/state vectors and border-class biases of ALL cell steps are built
                                 //   by two launches at the start of a rollout (they do not depend on the predicted frames)
    bool fuse_fin = false;       // VF_FUSE_FIN [0]: the producer's last-arriving warp/block finalises the statistics (no finalize launch);
                                 //   measured SLOWER (thin convs 80 -> 148 us): the last arriver's serial walk over up to 34 slots sits on
                                 //   an epilogue warp's critical path, the separate 4.5 us finalize launch is cheaper
  } opt;
  int cur_tau = 0;             // cell step being launched (selects the step's slice of the hoisted sabias tables)
  int cur_M = 0;               // samples of the rollout being launched (convs on fewer samples = shared-prefix steps)

  // profiling (vf_profile_*)
  bool prof_on = false;
  struct ProfRec { cudaEvent_t a, b; int cls; double flops; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_pool;
};

namespace {

int fail(vf_engine* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

// sticky device faults: a kernel of this library traps (mbar_wait in conv_mma.cu) instead of hanging when an mbarrier arrival is lost
const char* cuda_hint(cudaError_t e) {
  if (e == cudaErrorLaunchFailure || e == cudaErrorIllegalInstruction || e == cudaErrorIllegalAddress || e == cudaErrorAssert)
    return " [device-side trap or fault in a vfengine kernel; the CUDA context is lost: vf_destroy the handle and restart the process]";
  return "";
}
#define CU(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return fail(h, VF_ERR_CUDA, "%s:%d %s -> %s%s", __FILE__, __LINE__, #call, cudaGetErrorString(e__), cuda_hint(e__)); \
  } while (0)

template <typename T>
int dalloc(vf_engine* h, T** p, size_t n) {
  void* q = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) return fail(h, VF_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = (T*)q;
  return VF_OK;
}
// cudaFree a buffer that dalloc() registered in the handle (grow-only buffers release their previous copy)
void dfree(vf_engine* h, void* p) {
  if (!p) return;
  auto it = std::find(h->allocs.begin(), h->allocs.end(), p);
  if (it != h->allocs.end()) h->allocs.erase(it);
  cudaFree(p);
}
void drop_graphs(vf_engine* h) {
  for (auto& kv : h->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->graphs.clear();
}
// device scratch of one C-ABI call, released when the call returns (after the stream has been synchronised)
struct Scratch {
  std::vector<void*> p;
  ~Scratch() { for (void* q : p) cudaFree(q); }
  template <typename T>
  bool get(T** out, size_t n) {
    void* q = nullptr;
    if (cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return false;
    p.push_back(q);
    *out = (T*)q;
    return true;
  }
};
#define DA(p, n)                          \
  do {                                    \
    int r__ = dalloc(h, &(p), (size_t)(n)); \
    if (r__) return r__;                  \
  } while (0)

const HostTensor* find_w(vf_engine* h, int view, const std::string& name) {
  char pre[32];
  snprintf(pre, sizeof(pre), "view%d.", view);
  auto it = h->host_w.find(std::string(pre) + name);
  if (it != h->host_w.end()) return &it->second;
  if (view == 0) {
    it = h->host_w.find(name);
    if (it != h->host_w.end()) return &it->second;
  }
  return nullptr;
}

int upload(vf_engine* h, float** dst, const std::vector<float>& v) {
  DA(*dst, v.size());
  CU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return VF_OK;
}

// Split an HWIO conv weight into the spatial part [k*k][cin_sp][cout] and the border-class sums of the
// tiled action/state channels wcls[cls][a][cout] (see conv_simt.cu header).
int prepare_conv(vf_engine* h, int view, ConvLayer& L) {
  const int A = L.cin_const, k = L.k, kk = k * k;
  const HostTensor* w = find_w(h, view, L.name + ".w");
  if (!w) return fail(h, VF_ERR_STATE, "missing weight view%d.%s.w", view, L.name.c_str());
  if (!L.cin_w) L.cin_w = L.cin_sp;
  const int cin_total = L.cin_w + A;
  if (w->shape.size() != 4 || w->shape[0] != k || w->shape[1] != k || w->shape[2] != cin_total || w->shape[3] != L.cout)
    return fail(h, VF_ERR_INVALID, "weight %s.w has wrong shape (want %d,%d,%d,%d)", L.name.c_str(), k, k, cin_total, L.cout);
  // channel order: conv [spatial, sa]; lstm [x, sa, h]
  const int xch = L.lstm ? L.cin_w / 2 : L.cin_w;
  std::vector<float> wsp((size_t)kk * L.cin_sp * L.cout, 0.f);
  for (int t = 0; t < kk; ++t)
    for (int c = 0; c < L.cin_w; ++c) {
      const int src_c = (c < xch) ? c : c + A;
      memcpy(&wsp[((size_t)t * L.cin_sp + c) * L.cout], &w->data[((size_t)t * cin_total + src_c) * L.cout], L.cout * sizeof(float));
    }
  int r = upload(h, &L.w_sp, wsp);
  if (r) return r;
  if (A > 0) {
    const int pad = k / 2;
    std::vector<float> wc((size_t)kk * A * L.cout);
    for (int cy = 0; cy < k; ++cy)
      for (int cx = 0; cx < k; ++cx) {
        const int yr = cy < pad ? cy : (cy > pad ? L.H - 1 - (k - 1 - cy) : pad);
        const int xr = cx < pad ? cx : (cx > pad ? L.W - 1 - (k - 1 - cx) : pad);
        for (int a = 0; a < A; ++a)
          for (int n = 0; n < L.cout; ++n) {
            double s = 0.0;
            for (int dy = 0; dy < k; ++dy) {
              const int yy = yr + dy - pad;
              if (yy < 0 || yy >= L.H) continue;
              for (int dx = 0; dx < k; ++dx) {
                const int xx = xr + dx - pad;
                if (xx < 0 || xx >= L.W) continue;
                s += (double)w->data[((size_t)(dy * k + dx) * cin_total + xch + a) * L.cout + n];
              }
            }
            wc[((size_t)(cy * k + cx) * A + a) * L.cout + n] = (float)s;
          }
      }
    r = upload(h, &L.wcls, wc);
    if (r) return r;
    L.sab_step = (long long)h->B * kk * L.cout;
    DA(L.sabias, (size_t)L.sab_step * (h->opt.hoist_sa ? (size_t)(h->S - 1) : 1));
  }
  auto opt = [&](const char* suffix, float** dst, int n) -> int {
    const HostTensor* t = find_w(h, view, L.name + suffix);
    if (!t) return fail(h, VF_ERR_STATE, "missing weight view%d.%s%s", view, L.name.c_str(), suffix);
    if ((int)t->data.size() != n) return fail(h, VF_ERR_INVALID, "weight %s%s has %zu elements, want %d", L.name.c_str(), suffix, t->data.size(), n);
    return upload(h, dst, t->data);
  };
  if (L.lstm) {
    if ((r = opt(".gates_gamma", &L.gamma, L.cout))) return r;
    if ((r = opt(".gates_beta", &L.beta, L.cout))) return r;
    if ((r = opt(".cell_gamma", &L.cgamma, L.cout / 4))) return r;
    if ((r = opt(".cell_beta", &L.cbeta, L.cout / 4))) return r;
  } else {
    if ((r = opt(".b", &L.bias, L.cout))) return r;
    if (find_w(h, view, L.name + ".gamma")) {
      if ((r = opt(".gamma", &L.gamma, L.cout))) return r;
      if ((r = opt(".beta", &L.beta, L.cout))) return r;
    }
  }
  if (h->cfg.precision != VF_PREC_FP32_SIMT && L.s2d) {
    // pooled(Y,X) = 1/4 sum_{py,px in {0,1}} sum_{dy,dx} w[dy][dx][c] in(2Y+py+dy-pad, 2X+px+dx-pad, c).  With input pixel
    // (2(Y+by-1)+sy, 2(X+bx-1)+sx) = block (Y+by-1, X+bx-1), sub-position (sy,sx):  dy = 2by+sy-py+pad-2, dx likewise, so
    //   w'[by][bx][(sy*2+sx)*cin_sp + c][n] = 1/4 sum_{py,px} w[2by+sy-py+pad-2][2bx+sx-px+pad-2][c][n]   (taps outside 0..k-1 drop).
    // SAME zero padding of pad <= 2 pixels is exactly the one zero BLOCK of the 3x3 block conv (H, W even).
    const int pad = k / 2, C4 = 4 * L.cin_sp;
    std::vector<float> w2((size_t)9 * C4 * L.cout, 0.f);
    for (int by = 0; by < 3; ++by)
      for (int bx = 0; bx < 3; ++bx)
        for (int sy = 0; sy < 2; ++sy)
          for (int sx = 0; sx < 2; ++sx)
            for (int c = 0; c < L.cin_sp; ++c)
              for (int n = 0; n < L.cout; ++n) {
                double acc = 0.0;
                for (int py = 0; py < 2; ++py)
                  for (int px_ = 0; px_ < 2; ++px_) {
                    const int dy = 2 * by + sy - py + pad - 2, dx = 2 * bx + sx - px_ + pad - 2;
                    if (dy < 0 || dy >= k || dx < 0 || dx >= k) continue;
                    acc += (double)wsp[((size_t)(dy * k + dx) * L.cin_sp + c) * L.cout + n];
                  }
                w2[((size_t)(by * 3 + bx) * C4 + (sy * 2 + sx) * L.cin_sp + c) * L.cout + n] = (float)(0.25 * acc);
              }
    if (A > 0) {
      // border classes of the pooled output: pooled row 0 averages pixel rows 0,1 (classes 0,1 of the k x k conv), the last
      // pooled row the classes k-2,k-1, every other row the interior class (needs H/2 >= 3 rows: checked in build_net)
      std::vector<float> wc_full((size_t)kk * A * L.cout), wc9((size_t)9 * A * L.cout);
      CU(cudaMemcpy(wc_full.data(), L.wcls, wc_full.size() * sizeof(float), cudaMemcpyDeviceToHost));
      auto cls_of = [&](int c3, int p) { return c3 == 0 ? p : (c3 == 2 ? k - 2 + p : pad); };
      for (int cy = 0; cy < 3; ++cy)
        for (int cx = 0; cx < 3; ++cx)
          for (size_t i = 0; i < (size_t)A * L.cout; ++i) {
            double acc = 0.0;
            for (int py = 0; py < 2; ++py)
              for (int px_ = 0; px_ < 2; ++px_) acc += (double)wc_full[((size_t)(cls_of(cy, py) * k + cls_of(cx, px_))) * A * L.cout + i];
            wc9[(size_t)(cy * 3 + cx) * A * L.cout + i] = (float)(0.25 * acc);
          }
      CU(cudaMemcpy(L.wcls, wc9.data(), wc9.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    L.kcls = 3;
    std::string e;
    if (mma_conv_prepare_weights(w2.data(), 3, 3, 3, C4, L.cout, &L.mma, &h->allocs, &e))
      return fail(h, VF_ERR_CUDA, "mma weight prep %s: %s", L.name.c_str(), e.c_str());
  } else if (h->cfg.precision != VF_PREC_FP32_SIMT && L.fold) {
    // folded[dy][dx*cin_sp + c][n] = w[dy][dx][c][n]
    std::vector<float> wf((size_t)k * k * L.cin_sp * L.cout);
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx)
        memcpy(&wf[((size_t)dy * k * L.cin_sp + (size_t)dx * L.cin_sp) * L.cout], &wsp[((size_t)(dy * k + dx) * L.cin_sp) * L.cout],
               (size_t)L.cin_sp * L.cout * sizeof(float));
    std::string e;
    if (mma_conv_prepare_weights(wf.data(), k, 1, k, k * L.cin_sp, L.cout, &L.mma, &h->allocs, &e))
      return fail(h, VF_ERR_CUDA, "mma weight prep %s: %s", L.name.c_str(), e.c_str());
  } else if (h->cfg.precision != VF_PREC_FP32_SIMT && mma_conv_supported(L.k, L.cin_sp, L.cout, L.H, L.W)) {
    std::string e;
    if (mma_conv_prepare_weights(wsp.data(), L.k, L.k, L.k, L.cin_sp, L.cout, &L.mma, &h->allocs, &e))
      return fail(h, VF_ERR_CUDA, "mma weight prep %s: %s", L.name.c_str(), e.c_str());
  }
  return VF_OK;
}

ConvLayer mk(const std::string& name, int k, int H, int W, int cin_sp, int cin_const, int cout, bool lstm) {
  ConvLayer L;
  L.name = name; L.k = k; L.H = H; L.W = W; L.cin_sp = cin_sp; L.cin_const = cin_const; L.cout = cout; L.lstm = lstm;
  return L;
}

int build_net(vf_engine* h) {
  const vf_config& c = h->cfg;
  const int n = c.n_enc, A = h->A, B = h->B;
  h->views.resize(h->ncam);
  size_t raw_max = 0, decin_max = 0, cmax = 0, fmax_ = 0, stat_max = 0;
  auto upd = [&](const ConvLayer& L) {
    raw_max = std::max(raw_max, (size_t)L.H * L.W * L.cout);
    cmax = std::max(cmax, (size_t)L.cout);
    // partial-sum slots per (sample, channel): k_plane_stats / k_lstm_gates use <= 16, a conv epilogue one or two per pass
    int slots = 16;
    if (c.precision != VF_PREC_FP32_SIMT) {
      const bool sd = L.s2d;
      slots = std::max(slots, mma_conv_stats_slots(sd ? 3 : L.k, L.fold ? 1 : (sd ? 3 : L.k), L.fold ? L.k * L.cin_sp : (sd ? 4 * L.cin_sp : L.cin_sp),
                                                   L.cout, sd ? L.H / 2 : L.H, sd ? L.W / 2 : L.W));
    }
    stat_max = std::max(stat_max, (size_t)L.cout * slots);
  };
  for (int v = 0; v < h->ncam; ++v) {
    ViewNet& net = h->views[v];
    int hh = h->H, ww = h->W, cprev = 6;
    std::vector<int> enc_out;
    net.enc_lstm.resize(n); net.enc_rnn.resize(n); net.dec_lstm.resize(n); net.dec_rnn.resize(n);
    for (int i = 0; i < n; ++i) {
      const int oc = c.enc_channels[i];
      char nm[64];
      snprintf(nm, sizeof(nm), "enc%d.conv", i);
      net.enc_conv.push_back(mk(nm, i == 0 ? 5 : 3, hh, ww, cprev, A, oc, false));
      if (i == 0 && c.precision != VF_PREC_FP32_SIMT) {      // (image, first) packed to 8 channels for the tensor-core path
        net.enc_conv.back().cin_sp = 8;
        net.enc_conv.back().cin_w = 6;
        // default: conv + pool as one 3x3 conv over 2x2 pixel blocks (s2d); VF_ENC0=fold keeps the full-resolution k x 1 conv
        // over the dx-folded input (A/B switch and fallback for shapes the block conv does not cover)
        const char* e0 = getenv("VF_ENC0");
        const bool want_fold = e0 && !strcmp(e0, "fold");
        if (!want_fold && hh % 2 == 0 && ww % 2 == 0 && hh >= 6 && ww >= 6 && mma_conv_supported(3, 32, oc, hh / 2, ww / 2))
          net.enc_conv.back().s2d = true;
        else
          net.enc_conv.back().fold = true;
      }
      if (i > 0 && c.precision != VF_PREC_FP32_SIMT && c.enc_rnn[i - 1]) {
        // conv 3x3 + 2x2 pool as ONE 3x3 conv over 2x2 pixel blocks of the previous conv-LSTM's h, which k_lstm_out also
        // writes in space-to-depth layout.  Opt-in (VF_ENC_S2D=1): measured 0.3 ms of 102 per plan (the MACs are the same, only
        // the pooled reads go away) and the 128-px / 15-step distribution error moves from 0.9e-5 to 1.1e-5.
        const char* e1 = getenv("VF_ENC_S2D");
        if (e1 && atoi(e1) == 1 && hh % 2 == 0 && ww % 2 == 0 && hh >= 6 && ww >= 6 && cprev % 8 == 0 &&
            mma_conv_supported(3, 4 * cprev, oc, hh / 2, ww / 2)) {
          net.enc_conv.back().s2d = true;
          DA(net.enc_rnn[i - 1].h_s2d, (size_t)B * hh * ww * cprev);
        }
      }
      upd(net.enc_conv.back());
      hh /= 2; ww /= 2;
      if (c.enc_rnn[i]) {
        snprintf(nm, sizeof(nm), "enc%d.lstm", i);
        net.enc_lstm[i] = mk(nm, c.lstm_ksize, hh, ww, 2 * oc, A, 4 * oc, true);
        upd(net.enc_lstm[i]);
        RnnState& r = net.enc_rnn[i];
        r.h = hh; r.w = ww; r.F = oc;
        DA(r.lstm_in, (size_t)B * hh * ww * 2 * oc);
        DA(r.c, (size_t)B * hh * ww * oc);
        DA(r.snap_in, (size_t)hh * ww * 2 * oc);
        DA(r.snap_c, (size_t)hh * ww * oc);
        fmax_ = std::max(fmax_, (size_t)oc);
      }
      enc_out.push_back(oc);
      cprev = oc;
    }
    for (int i = 0; i < n; ++i) {
      const int oc = c.dec_channels[i];
      const int cin = cprev + (i > 0 ? enc_out[n - 1 - i] : 0);
      hh *= 2; ww *= 2;
      decin_max = std::max(decin_max, (size_t)hh * ww * cin);
      char nm[64];
      snprintf(nm, sizeof(nm), "dec%d.conv", i);
      net.dec_conv.push_back(mk(nm, 3, hh, ww, cin, A, oc, false));
      upd(net.dec_conv.back());
      if (c.dec_rnn[i]) {
        snprintf(nm, sizeof(nm), "dec%d.lstm", i);
        net.dec_lstm[i] = mk(nm, c.lstm_ksize, hh, ww, 2 * oc, A, 4 * oc, true);
        upd(net.dec_lstm[i]);
        RnnState& r = net.dec_rnn[i];
        r.h = hh; r.w = ww; r.F = oc;
        DA(r.lstm_in, (size_t)B * hh * ww * 2 * oc);
        DA(r.c, (size_t)B * hh * ww * oc);
        DA(r.snap_in, (size_t)hh * ww * 2 * oc);
        DA(r.snap_c, (size_t)hh * ww * oc);
        fmax_ = std::max(fmax_, (size_t)oc);
      }
      cprev = oc;
    }
    const int g = h->ngf;
    net.scratch0 = mk("scratch.conv0", 3, h->H, h->W, g, 0, g, false);
    net.scratch1 = mk("scratch.conv1", 3, h->H, h->W, g, 0, 3, false);
    net.masks0 = mk("masks.conv0", 3, h->H, h->W, g, 0, g, false);
    net.masks1 = mk("masks.conv1", 3, h->H, h->W, h->cm, 0, h->nm, false);
    net.masks1.cin_w = g + 3 * h->nm;        // the layers buffer is padded to a multiple of 8 channels (16-byte units)
    upd(net.scratch0);
    if (h->opt.merge_heads) {
      net.heads0 = mk("heads0", 3, h->H, h->W, g, 0, 2 * g, false);
      upd(net.heads0);
    }
  }
  if (h->opt.hoist_sa) {                                   // all-steps tables: (S-1) x the per-step size (c4 at M = 4096 on one GPU: 8 GiB of 180); cap 12 GiB
    size_t per_step = 0;
    for (auto& net : h->views) {
      auto add = [&](const ConvLayer& L) { if (L.k && L.cin_const > 0) per_step += (size_t)B * L.k * L.k * L.cout * sizeof(float); };
      for (auto& L : net.enc_conv) add(L);
      for (auto& L : net.dec_conv) add(L);
      for (auto& L : net.enc_lstm) add(L);
      for (auto& L : net.dec_lstm) add(L);
    }
    if (per_step * (size_t)(h->S - 1) > ((size_t)12 << 30)) h->opt.hoist_sa = false;
  }
  // shared scratch (first view's shapes == all views' shapes)
  DA(h->raw, (size_t)B * raw_max);
  DA(h->dec_in, (size_t)B * decin_max);
  DA(h->stats, (size_t)B * cmax * 2);
  DA(h->stats_partial, std::max(plane_stats_partial_doubles(B, (int)cmax), (size_t)B * stat_max * 2));
  DA(h->cstats_partial, plane_stats_partial_doubles(B, (int)std::max(fmax_, (size_t)1)));
  DA(h->cstats, (size_t)B * std::max(fmax_, (size_t)1) * 2);
  DA(h->stat_cnt, (size_t)B * VF_STAT_CNT_STRIDE);
  if (cudaMemset(h->stat_cnt, 0, (size_t)B * VF_STAT_CNT_STRIDE * sizeof(int)) != cudaSuccess) return fail(h, VF_ERR_CUDA, "memset stat_cnt");
  h->act_enc.assign(n, nullptr);
  h->act_dec.assign(n, nullptr);
  {
    int hh = h->H, ww = h->W;
    for (int i = 0; i < n; ++i) {
      hh /= 2; ww /= 2;
      if (!c.enc_rnn[i]) DA(h->act_enc[i], (size_t)B * hh * ww * c.enc_channels[i]);
    }
    for (int i = 0; i < n; ++i) {
      hh *= 2; ww *= 2;
      if (!c.dec_rnn[i]) DA(h->act_dec[i], (size_t)B * hh * ww * c.dec_channels[i]);
    }
  }
  const size_t px = (size_t)h->H * h->W;
  DA(h->scr_h, (size_t)B * px * h->ngf);           // two dense buffers even when ONE conv produces both (opt.merge_heads): each
  DA(h->mask_h, (size_t)B * px * h->ngf);          // consumer's TMA then reads whole 128-byte lines
  if (c.precision != VF_PREC_FP32_SIMT) DA(h->pack0, (size_t)B * px * 8 * 5);   // (image, first) x 5 dx taps, 8 channels each
  DA(h->layers, (size_t)B * px * h->cl);
  if (cudaMemset(h->layers, 0, (size_t)B * px * h->cl * sizeof(float)) != cudaSuccess) return fail(h, VF_ERR_CUDA, "memset layers");
  DA(h->logits, (size_t)B * px * h->nm);
  DA(h->kern, (size_t)B * h->nt * h->kc * h->kc);
  DA(h->cdna_part, cdna_partial_floats((h->H >> c.n_enc) * (h->W >> c.n_enc) * c.enc_channels[c.n_enc - 1], B));
  h->nblk = composite_blocks(h->H, h->W);
  DA(h->partial, (size_t)B * h->nd * h->nblk);
  return VF_OK;
}

int finalize_weights(vf_engine* h) {
  if (h->weights_ready) return VF_OK;
  for (int v = 0; v < h->ncam; ++v) {
    ViewNet& net = h->views[v];
    int r;
    for (auto& L : net.enc_conv) if ((r = prepare_conv(h, v, L))) return r;
    for (auto& L : net.dec_conv) if ((r = prepare_conv(h, v, L))) return r;
    for (auto& L : net.enc_lstm) if (L.k && (r = prepare_conv(h, v, L))) return r;
    for (auto& L : net.dec_lstm) if (L.k && (r = prepare_conv(h, v, L))) return r;
    if (h->opt.merge_heads) {
      // heads0 = [scratch.conv0 | masks.conv0] along the output channels (same input, same 3x3 geometry, both instance-normalised)
      const char* parts[4] = {".w", ".b", ".gamma", ".beta"};
      for (const char* suf : parts) {
        const HostTensor* a = find_w(h, v, std::string("scratch.conv0") + suf);
        const HostTensor* b = find_w(h, v, std::string("masks.conv0") + suf);
        if (!a || !b) return fail(h, VF_ERR_STATE, "missing weight view%d.{scratch,masks}.conv0%s", v, suf);
        if (a->shape != b->shape) return fail(h, VF_ERR_INVALID, "scratch.conv0%s and masks.conv0%s differ in shape", suf, suf);
        const size_t inner = (size_t)a->shape.back(), outer = a->data.size() / inner;
        HostTensor m;
        m.shape = a->shape;
        m.shape.back() = 2 * (int64_t)inner;
        m.data.resize(2 * a->data.size());
        for (size_t o = 0; o < outer; ++o) {
          memcpy(&m.data[o * 2 * inner], &a->data[o * inner], inner * sizeof(float));
          memcpy(&m.data[o * 2 * inner + inner], &b->data[o * inner], inner * sizeof(float));
        }
        char nm[64];
        snprintf(nm, sizeof(nm), "view%d.heads0%s", v, suf);
        h->host_w[nm] = std::move(m);
      }
      if ((r = prepare_conv(h, v, net.heads0))) return r;
    } else {
      if ((r = prepare_conv(h, v, net.scratch0))) return r;
      if ((r = prepare_conv(h, v, net.masks0))) return r;
    }
    if ((r = prepare_conv(h, v, net.scratch1))) return r;
    if ((r = prepare_conv(h, v, net.masks1))) return r;
    const HostTensor* cw = find_w(h, v, "cdna.dense.w");
    const HostTensor* cb = find_w(h, v, "cdna.dense.b");
    if (!cw || !cb) return fail(h, VF_ERR_STATE, "missing weight view%d.cdna.dense.{w,b}", v);
    const int nout = h->kc * h->kc * h->nt;
    const RnnState& sm = net.enc_rnn[h->n_enc - 1];
    const size_t feat = (size_t)(h->H >> h->n_enc) * (h->W >> h->n_enc) * h->cfg.enc_channels[h->n_enc - 1];
    (void)sm;
    if (cw->data.size() != feat * nout || (int)cb->data.size() != nout)
      return fail(h, VF_ERR_INVALID, "cdna.dense has wrong shape (want %zu x %d)", feat, nout);
    if ((r = upload(h, &net.cdna_w, cw->data))) return r;
    if ((r = upload(h, &net.cdna_b, cb->data))) return r;
  }
  for (int v = 0; v < h->ncam && h->sdim > 0; ++v) {
    const HostTensor* sw = find_w(h, v, "state.dense.w");
    const HostTensor* sb = find_w(h, v, "state.dense.b");
    if (!sw || !sb) return fail(h, VF_ERR_STATE, "missing weight view%d.state.dense.{w,b}", v);
    if ((int)sw->data.size() != (h->adim + h->sdim) * h->sdim || (int)sb->data.size() != h->sdim)
      return fail(h, VF_ERR_INVALID, "state.dense has wrong shape");
    int r;
    if ((r = upload(h, &h->views[v].w_state, sw->data))) return r;
    if ((r = upload(h, &h->views[v].b_state, sb->data))) return r;
    DA(h->views[v].state_cur, (size_t)h->B * h->sdim);
  }
  for (int v = 0; v < h->ncam && h->cfg.rnn_z; ++v) {
    const HostTensor* zw = find_w(h, v, "zrnn.w");
    const HostTensor* zb = find_w(h, v, "zrnn.b");
    if (!zw || !zb) return fail(h, VF_ERR_STATE, "missing weight view%d.zrnn.{w,b}", v);
    if ((int)zw->data.size() != 2 * h->nz * 4 * h->nz || (int)zb->data.size() != 4 * h->nz)
      return fail(h, VF_ERR_INVALID, "zrnn has wrong shape");
    int r;
    if ((r = upload(h, &h->views[v].w_z, zw->data))) return r;
    if ((r = upload(h, &h->views[v].b_z, zb->data))) return r;
    DA(h->views[v].zstate, (size_t)h->B * 2 * h->nz);
  }
  h->host_w.clear();
  h->weights_ready = true;
  return VF_OK;
}

// finalise the S partial sums of n planes into (mean, rstd) pairs and hand both to the consumer
StatsRef fin_stats(vf_engine* h, const double* partial, int S, int n, int npix, float* dst, bool finalized = false) {
  if (finalized) return stats_ref(partial, S, npix, h->cfg.norm_eps, dst);               // the producer kernel wrote (mean, rstd) already
  if (!h->opt.stats_fin) return stats_ref(partial, S, npix, h->cfg.norm_eps, nullptr);   // the consumer finalises in its prologue
  launch_stats_finalize(partial, n, S, npix, h->cfg.norm_eps, dst, h->stream);
  return stats_ref(partial, S, npix, h->cfg.norm_eps, dst);
}
View dense_view(float* p, int hw, int C) { return make_view(p, (long long)hw * C, C, 0, C); }
// view of a convolution-input buffer [B][hw][ps] (split-half storage on the tensor-core path: lo plane after B*hw*ps halfs)
View cview(const vf_engine* h, float* p, int hw, int ps, int ch_off, int C) {
  return make_view(p, (long long)hw * ps, ps, ch_off, C, h->split ? (long long)h->B * hw * ps : 0);
}

void run_conv_impl(vf_engine* h, const ConvLayer& L, View s0, View s1, View out, int B, int act, double* stats_partial, int* stats_slots,
                   bool* finalized);

cudaEvent_t prof_event(vf_engine* h) {
  if (!h->prof_pool.empty()) { cudaEvent_t e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// stats_partial / stats_slots: fused instance-norm partial sums (slots = 0: not fused, run k_plane_stats); finalized: the launch
// also wrote the (mean, rstd) pairs of its output planes into h->stats
void run_conv(vf_engine* h, const ConvLayer& L, View s0, View s1, View out, int B, int act = ACT_NONE, double* stats_partial = nullptr,
              int* stats_slots = nullptr, bool* finalized = nullptr) {
  if (stats_slots) *stats_slots = 0;
  if (finalized) *finalized = false;
  if (!h->prof_on) { run_conv_impl(h, L, s0, s1, out, B, act, stats_partial, stats_slots, finalized); return; }
  vf_engine::ProfRec r;
  r.a = prof_event(h); r.b = prof_event(h);
  r.cls = (B < h->cur_M) ? 2 : (L.lstm ? 0 : 1);           // class 2: shared-prefix steps (run on one sample)
  r.flops = 2.0 * B * L.H * L.W * L.k * L.k * (double)L.cin_sp * L.cout;
  cudaEventRecord(r.a, h->stream);
  run_conv_impl(h, L, s0, s1, out, B, act, stats_partial, stats_slots, finalized);
  cudaEventRecord(r.b, h->stream);
  h->prof.push_back(r);
}

void run_conv_impl(vf_engine* h, const ConvLayer& L, View s0, View s1, View out, int B, int act, double* stats_partial, int* stats_slots,
                   bool* finalized) {
  if (h->cfg.precision != VF_PREC_FP32_SIMT && L.mma.ready) {
    MmaConvCall c;
    c.src = s0; c.src1 = s1; c.out = out; c.bias = L.bias;
    c.sabias = L.sabias ? L.sabias + (h->opt.hoist_sa ? (long long)h->cur_tau * L.sab_step : 0) : nullptr;
    c.H = L.s2d ? L.H / 2 : L.H; c.W = L.s2d ? L.W / 2 : L.W;
    c.passes = (h->cfg.precision == VF_PREC_F16X3) ? 3 : 1;
    c.act = act;
    c.stats_partial = stats_partial;
    c.stats_slots = stats_slots;
    if (stats_partial && finalized && h->opt.fuse_fin) { c.stats_fin = h->stats; c.stats_cnt = h->stat_cnt; c.stats_finalized = finalized; }
    c.stats_eps = h->cfg.norm_eps;
    const int rc = mma_conv_launch(L.mma, c, B, h->stream);
    if (rc && !h->conv_error) {
      h->conv_error = rc;
      snprintf(h->conv_error_layer, sizeof(h->conv_error_layer), "%s", L.name.c_str());
    }
    return;
  }
  ConvArgs a;
  a.src0 = s0; a.src1 = s1; a.w = L.w_sp; a.bias = L.bias; a.out = out;
  a.sabias = L.sabias ? L.sabias + (h->opt.hoist_sa ? (long long)h->cur_tau * L.sab_step : 0) : nullptr;
  a.H = L.H; a.W = L.W; a.Cin = L.cin_sp; a.Cout = L.cout; a.k = L.k; a.act = act;
  launch_conv_simt(a, B, h->stream);
}

void run_lstm(vf_engine* h, int v, const ConvLayer& L, RnnState& r, int B, const std::string& dbg) {
  const int hw = r.h * r.w, F = r.F;
  View in = cview(h, r.lstm_in, hw, 2 * F, 0, 2 * F);
  View gates = dense_view(h->raw, hw, 4 * F);
  View none = make_view(nullptr, 0, 0, 0, 0);
  const float eps = h->cfg.norm_eps;
  int slots = 0;
  bool gfin = false, cfin = false;
  float* ffin = h->opt.fuse_fin ? h->stats : nullptr;
  float* cffin = h->opt.fuse_fin ? h->cstats : nullptr;
  run_conv(h, L, in, none, gates, B, ACT_NONE, h->stats_partial, &slots, &gfin);     // gate statistics fused into the epilogue
  if (slots == 0) {
    slots = launch_plane_stats(gates, B, r.h, r.w, 0, h->stats_partial, h->stream, ffin, h->stat_cnt, eps);
    gfin = ffin != nullptr && (4 * F + 31) / 32 <= VF_STAT_CNT_STRIDE;
  }
  const StatsRef gsr = fin_stats(h, h->stats_partial, slots, B * 4 * F, hw, h->stats, gfin);
  View hv = cview(h, r.lstm_in, hw, 2 * F, F, F);
  View h2 = r.h_s2d ? cview(h, r.h_s2d, hw / 4, 4 * F, 0, 4 * F) : make_view(nullptr, 0, 0, 0, 0);
  h->debug[v][dbg + ".h"] = DebugEntry{hv, r.h, r.w};
  h->debug[v][dbg + ".c"] = DebugEntry{dense_view(r.c, hw, F), r.h, r.w};
  if (h->opt.lstm_fused && gsr.fin &&
      launch_lstm_fused(gates, B, hw, F, gsr.fin, L.gamma, L.beta, h->cfg.forget_bias, L.cgamma, L.cbeta, eps, r.c, hv, h->stream, h2, r.w))
    return;
  int cslots;
  if (256 % F == 0) {     // cell-state statistics fused into the pointwise kernel
    cslots = launch_lstm_gates(gates, B, hw, F, gsr, L.gamma, L.beta, h->cfg.forget_bias, r.c, h->cstats_partial, h->stream, cffin,
                               h->stat_cnt, eps);
    cfin = cffin != nullptr;
  } else {
    if (!gsr.fin) launch_stats_finalize(h->stats_partial, B * 4 * F, slots, hw, eps, h->stats, h->stream);   // the generic kernel reads finalised pairs
    launch_lstm_gates_generic(gates, B, hw, F, h->stats, L.gamma, L.beta, h->cfg.forget_bias, r.c, h->stream);
    cslots = launch_plane_stats(dense_view(r.c, hw, F), B, r.h, r.w, 0, h->cstats_partial, h->stream, cffin, h->stat_cnt, eps);
    cfin = cffin != nullptr && (F + 31) / 32 <= VF_STAT_CNT_STRIDE;
  }
  launch_lstm_out(gates, B, hw, F, gsr, L.gamma, L.beta, fin_stats(h, h->cstats_partial, cslots, B * F, hw, h->cstats, cfin), L.cgamma,
                  L.cbeta, r.c, hv, h->stream, h2, r.w);
}

// sabias[b][cls][cout] = bias + sa[b] . wcls[cls] for every layer of view v; nsteps > 1: all cell steps in one launch
void launch_step_sabias(vf_engine* h, int v, int B, int nsteps, const float* sa, long long sa_step) {
  ViewNet& net = h->views[v];
  SabiasBatch sbb;
  sbb.n = 0; sbb.A = h->A; sbb.B = B; sbb.sa = sa; sbb.nsteps = nsteps; sbb.sa_step = sa_step;
  auto sab = [&](ConvLayer& L) {
    if (!(L.k && L.wcls)) return;
    if (sbb.n == 24) { launch_sabias_batch(sbb, h->stream); sbb.n = 0; }
    SabiasBatch::Layer& e = sbb.L[sbb.n++];
    const int kc_ = L.kcls ? L.kcls : L.k;
    e.wcls = L.wcls; e.bias = L.bias; e.out = L.sabias; e.ncls = kc_ * kc_; e.Cout = L.cout; e.out_step = L.sab_step;
  };
  for (int i = 0; i < h->n_enc; ++i) { sab(net.enc_conv[i]); sab(net.enc_lstm[i]); sab(net.dec_conv[i]); sab(net.dec_lstm[i]); }
  launch_sabias_batch(sbb, h->stream);
}

// statistics pass over a conv output whose epilogue did not fuse them; *finalized: the pass wrote (mean, rstd) into h->stats too
int plane_stats_pass(vf_engine* h, View x, int B, int H, int W, int pool, bool* finalized) {
  float* fin = h->opt.fuse_fin ? h->stats : nullptr;
  *finalized = fin != nullptr && (x.C + 31) / 32 <= VF_STAT_CNT_STRIDE;
  return launch_plane_stats(x, B, H, W, pool, h->stats_partial, h->stream, fin, h->stat_cnt, h->cfg.norm_eps);
}

// one cell step of one view (spec P1-P8)
void run_step(vf_engine* h, int v, int tau, int B) {
  const vf_config& c = h->cfg;
  ViewNet& net = h->views[v];
  const int n = h->n_enc, H = h->H, W = h->W, nd = h->nd;
  const long long px = (long long)H * W;
  View none = make_view(nullptr, 0, 0, 0, 0);
  // P1 inputs
  View image, distrib;
  if (tau < h->C) {
    image = make_view(h->ctx_frames + ((long long)tau * h->ncam + v) * px * 3, 0, 3, 0, 3);
    distrib = make_view(h->ctx_distrib + ((long long)tau * h->ncam + v) * px * nd, 0, nd, 0, nd);
  } else {
    const int t = tau - h->C;
    image = make_view(h->gen_images + ((long long)t * h->ncam + v) * px * 3, (long long)h->P * h->ncam * px * 3, 3, 0, 3);
    distrib = make_view(h->gen_distrib + ((long long)t * h->ncam + v) * px * nd, (long long)h->P * h->ncam * px * nd, nd, 0, nd);
  }
  View first = make_view(h->ctx_frames + (long long)v * px * 3, 0, 3, 0, 3);
  View first_d = make_view(h->ctx_distrib + (long long)v * px * nd, 0, nd, 0, nd);

  // per-layer border-class bias of the tiled action/state vector (hoisted: built for every step at the start of the rollout)
  if (!h->opt.hoist_sa) launch_step_sabias(h, v, B, 1, h->sa, 0);

  // P2 encoder
  std::vector<View> enc_out(n);
  std::vector<int> enc_h(n), enc_w(n);
  View x0 = image, x1 = first;
  if (h->pack0 && net.enc_conv[0].s2d) {                    // 2x2 pixel blocks -> channels: [sub-position][image rgb, first rgb, 0, 0]
    x0 = cview(h, h->pack0, (H / 2) * (W / 2), 32, 0, 32);
    launch_pack_s2d(image, first, B, H, W, x0, h->stream);
    x1 = none;
  } else if (h->pack0) {
    const int kf = net.enc_conv[0].k;                       // dx taps folded into channels: [dx][image rgb, first rgb, 0, 0]
    x0 = cview(h, h->pack0, H * W, 8 * kf, 0, 8 * kf);
    launch_pack_fold(image, first, B, H, W, kf, x0, h->stream);
    x1 = none;
  }
  int hh = H, ww = W;
  for (int i = 0; i < n; ++i) {
    ConvLayer& L = net.enc_conv[i];
    const int oc = L.cout;
    const int pooled = L.s2d ? 0 : 1;                       // s2d: the conv output is the pooled map already
    View rawv = dense_view(h->raw, L.s2d ? (hh / 2) * (ww / 2) : hh * ww, oc);
    int S_e = 0;                                             // statistics of the POOLED map: only a pool-fused conv can supply them
    bool f_e = false;
    run_conv(h, L, x0, x1, rawv, B, ACT_NONE, (h->opt.epi_stats && !pooled) ? h->stats_partial : nullptr, &S_e, &f_e);
    hh /= 2; ww /= 2;
    if (!S_e) S_e = plane_stats_pass(h, rawv, B, hh, ww, pooled, &f_e);
    View dst = c.enc_rnn[i] ? cview(h, net.enc_rnn[i].lstm_in, hh * ww, 2 * oc, 0, oc)
                            : cview(h, h->act_enc[i], hh * ww, oc, 0, oc);
    launch_norm_act(rawv, B, hh, ww, pooled, fin_stats(h, h->stats_partial, S_e, B * oc, hh * ww, h->stats, f_e), L.gamma, L.beta, ACT_RELU, dst, h->stream);
    h->debug[v][L.name] = DebugEntry{dst, hh, ww};
    View out = dst;
    if (c.enc_rnn[i]) {
      run_lstm(h, v, net.enc_lstm[i], net.enc_rnn[i], B, net.enc_lstm[i].name);
      out = cview(h, net.enc_rnn[i].lstm_in, hh * ww, 2 * oc, oc, oc);
    }
    enc_out[i] = out; enc_h[i] = hh; enc_w[i] = ww;
    x0 = out; x1 = none;
    if (i + 1 < n && net.enc_conv[i + 1].s2d) x0 = cview(h, net.enc_rnn[i].h_s2d, (hh / 2) * (ww / 2), 4 * oc, 0, 4 * oc);
  }
  // P5/P6 on a side branch: the CDNA head only needs the last encoder state and the previous frame, and nothing before the
  // scratch-image conv needs its outputs — it overlaps the decoder (latency-shaped kernels beside the decoder's streaming ones)
  const bool emit = tau >= h->C - 1;                      // a warm-up step's prediction is never consumed
  const bool side = emit && h->opt.side_cdna && h->side;
  const int featK = enc_h[n - 1] * enc_w[n - 1] * enc_out[n - 1].C;
  View layers = cview(h, h->layers, (int)px, h->cl, 0, h->cl);
  if (side) {
    cudaEventRecord(h->ev_fork, h->stream);
    cudaStreamWaitEvent(h->side, h->ev_fork, 0);
    launch_cdna_kernels(enc_out[n - 1], enc_h[n - 1] * enc_w[n - 1], net.cdna_w, h->kc, h->nt, B, h->cdna_part, h->side);
    launch_cdna_apply(image, first, h->cdna_part, featK, net.cdna_b, h->kern, h->kc, h->nt, B, H, W, layers, h->side);
    cudaEventRecord(h->ev_join, h->side);
  }
  // P4 decoder
  View x = enc_out[n - 1];
  for (int i = 0; i < n; ++i) {
    ConvLayer& L = net.dec_conv[i];
    const int oc = L.cout;
    View skip = i > 0 ? enc_out[n - 1 - i] : none;
    View din = cview(h, h->dec_in, 4 * hh * ww, x.C + skip.C, 0, x.C + skip.C);
    launch_upsample2x(x, skip, B, hh, ww, din, h->stream);
    hh *= 2; ww *= 2;
    View rawv = dense_view(h->raw, hh * ww, oc);
    int S_d = 0;
    bool f_d = false;
    run_conv(h, L, din, none, rawv, B, ACT_NONE, h->opt.epi_stats ? h->stats_partial : nullptr, &S_d, &f_d);
    if (!S_d) S_d = plane_stats_pass(h, rawv, B, hh, ww, 0, &f_d);
    View dst = c.dec_rnn[i] ? cview(h, net.dec_rnn[i].lstm_in, hh * ww, 2 * oc, 0, oc)
                            : cview(h, h->act_dec[i], hh * ww, oc, 0, oc);
    launch_norm_act(rawv, B, hh, ww, 0, fin_stats(h, h->stats_partial, S_d, B * oc, hh * ww, h->stats, f_d), L.gamma, L.beta, ACT_RELU, dst, h->stream);
    h->debug[v][L.name] = DebugEntry{dst, hh, ww};
    x = dst;
    if (c.dec_rnn[i]) {
      run_lstm(h, v, net.dec_lstm[i], net.dec_rnn[i], B, net.dec_lstm[i].name);
      x = cview(h, net.dec_rnn[i].lstm_in, hh * ww, 2 * oc, oc, oc);
    }
  }
  if (!emit) return;
  const int t_out = tau - (h->C - 1);
  const int g = h->ngf, nm = h->nm, cl = h->cl;
  View h_last = x;
  if (!side) {                  // P5/P6 in line
    launch_cdna_kernels(enc_out[n - 1], enc_h[n - 1] * enc_w[n - 1], net.cdna_w, h->kc, h->nt, B, h->cdna_part, h->stream);
    launch_cdna_apply(image, first, h->cdna_part, featK, net.cdna_b, h->kern, h->kc, h->nt, B, H, W, layers, h->stream);
  }
  double* epi_sp = h->opt.epi_stats ? h->stats_partial : nullptr;
  View scr, hm;
  View scratch_out = cview(h, h->layers, (int)px, cl, 3 * (h->nt + 2), 3);
  if (h->opt.merge_heads) {
    // P7 + P8 first layers: scratch.conv0 and masks.conv0 read the same map -> ONE 3x3 conv ngf -> 2*ngf, one statistics set,
    // one normalise pass; the two hidden maps are the channel halves of heads_h
    View raw2 = dense_view(h->raw, (int)px, 2 * g);
    int S_h = 0;
    bool f_h = false;
    run_conv(h, net.heads0, h_last, none, raw2, B, ACT_NONE, epi_sp, &S_h, &f_h);
    if (!S_h) S_h = plane_stats_pass(h, raw2, B, H, W, 0, &f_h);
    scr = cview(h, h->scr_h, (int)px, g, 0, g);
    hm = cview(h, h->mask_h, (int)px, g, 0, g);
    launch_norm_act(raw2, B, H, W, 0, fin_stats(h, h->stats_partial, S_h, B * 2 * g, (int)px, h->stats, f_h), net.heads0.gamma, net.heads0.beta, ACT_RELU, scr, h->stream, hm);
    if (side) cudaStreamWaitEvent(h->stream, h->ev_join, 0);     // `layers` (written whole by the CDNA apply) before the scratch image lands in it
    run_conv(h, net.scratch1, scr, none, scratch_out, B, ACT_SIGMOID);
  } else {
    // P7 scratch image
    View rawg = dense_view(h->raw, (int)px, g);
    int S_s = 0;
    bool f_s = false;
    run_conv(h, net.scratch0, h_last, none, rawg, B, ACT_NONE, epi_sp, &S_s, &f_s);
    if (!S_s) S_s = plane_stats_pass(h, rawg, B, H, W, 0, &f_s);
    scr = cview(h, h->scr_h, (int)px, g, 0, g);
    launch_norm_act(rawg, B, H, W, 0, fin_stats(h, h->stats_partial, S_s, B * g, (int)px, h->stats, f_s), net.scratch0.gamma, net.scratch0.beta, ACT_RELU, scr, h->stream);
    if (side) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
    run_conv(h, net.scratch1, scr, none, scratch_out, B, ACT_SIGMOID);
    // P8 masks
    int S_m = 0;
    bool f_m = false;
    run_conv(h, net.masks0, h_last, none, rawg, B, ACT_NONE, epi_sp, &S_m, &f_m);
    if (!S_m) S_m = plane_stats_pass(h, rawg, B, H, W, 0, &f_m);
    hm = cview(h, h->mask_h, (int)px, g, 0, g);
    launch_norm_act(rawg, B, H, W, 0, fin_stats(h, h->stats_partial, S_m, B * g, (int)px, h->stats, f_m), net.masks0.gamma, net.masks0.beta, ACT_RELU, hm, h->stream);
  }
  View lg = dense_view(h->logits, (int)px, nm);
  run_conv(h, net.masks1, hm, layers, lg, B);
  CompositeArgs ca;
  ca.logits = lg;
  ca.layers = layers;
  ca.prev_d = distrib; ca.first_d = first_d; ca.kern = h->kern;
  ca.gen_image = make_view(h->gen_images + ((long long)t_out * h->ncam + v) * px * 3, (long long)h->P * h->ncam * px * 3, 3, 0, 3);
  ca.gen_distrib = make_view(h->gen_distrib + ((long long)t_out * h->ncam + v) * px * nd, (long long)h->P * h->ncam * px * nd, nd, 0, nd);
  ca.partial = h->partial; ca.nt = h->nt; ca.ksize = h->kc; ca.nd = nd; ca.H = H; ca.W = W;
  launch_composite(ca, B, h->stream);
  launch_distrib_normalize(ca.gen_distrib, h->partial, h->nblk, B, H, W, nd, h->stream);
  h->debug[v]["scratch"] = DebugEntry{scratch_out, H, W};
  h->debug[v]["mask_logits"] = DebugEntry{lg, H, W};
  h->debug[v]["gen_image"] = DebugEntry{ca.gen_image, H, W};
  h->debug[v]["gen_distrib"] = DebugEntry{ca.gen_distrib, H, W};
  h->debug[v]["cdna.kernels"] = DebugEntry{make_view(h->kern, (long long)h->nt * h->kc * h->kc, h->nt * h->kc * h->kc, 0, h->nt * h->kc * h->kc), 1, 1};
  h->debug[v]["dec_last"] = DebugEntry{h_last, H, W};
}

int shared_prefix_steps(const vf_engine* h, int M) {
  if (!h->opt.shared_prefix) return 0;
  return (h->nz == 0 && M > 1) ? std::max(0, std::min(h->n_ctx_actions, h->C - 1)) : 0;
}

// prefix modes of a rollout: the recurrent state after the shared-prefix steps is the same for every rollout of one plan
// (the context does not change between CEM iterations), so iteration 0 saves sample 0's state (PREFIX_SAVE) and the later
// iterations start from it (PREFIX_RESTORE) instead of recomputing the prefix; vf_predict uses PREFIX_NONE.
enum { PREFIX_NONE = 0, PREFIX_SAVE = 1, PREFIX_RESTORE = 2 };

// rolls S-1 cell steps for M samples whose actions are in h->actions [M][T][adim]
int rollout_body(vf_engine* h, int M, int T, int mode) {
  const int n_shared = shared_prefix_steps(h, M);
  const bool restore = mode == PREFIX_RESTORE && n_shared > 0;
  for (auto& net : h->views) {
    if (net.zstate) CU(cudaMemsetAsync(net.zstate, 0, (size_t)M * 2 * h->nz * sizeof(float), h->stream));
    if (restore) continue;                           // every state row is overwritten from the snapshot
    // lstm_in holds two B-strided fp16 planes on the tensor-core path: clear the whole buffer
    const size_t Bz = h->split ? (size_t)h->B : (size_t)M;
    for (auto& r : net.enc_rnn) if (r.c) {
      CU(cudaMemsetAsync(r.c, 0, (size_t)M * r.h * r.w * r.F * sizeof(float), h->stream));
      CU(cudaMemsetAsync(r.lstm_in, 0, Bz * r.h * r.w * 2 * r.F * sizeof(float), h->stream));
    }
    for (auto& r : net.dec_rnn) if (r.c) {
      CU(cudaMemsetAsync(r.c, 0, (size_t)M * r.h * r.w * r.F * sizeof(float), h->stream));
      CU(cudaMemsetAsync(r.lstm_in, 0, Bz * r.h * r.w * 2 * r.F * sizeof(float), h->stream));
    }
  }
  SaArgs sa;
  sa.actions = h->actions; sa.T = T; sa.adim = h->adim; sa.sdim = h->sdim; sa.nz = h->nz;
  sa.n_ctx_actions = h->n_ctx_actions; sa.C = h->C; sa.ctx_actions = h->ctx_actions; sa.ctx_states = h->ctx_states;
  sa.zs = h->nz ? h->zs : nullptr; sa.sa = h->sa; sa.P = h->P;
  // Shared prefix: while the cell is still fed context frames, context states AND context actions (tau < n_ctx_actions,
  // tau < C-1 so that no prediction is emitted) and there is no per-sample latent, every one of the M samples computes
  // exactly the same thing from the same zero state (the reference tiles the context per tower and recomputes it,
  // setup_predictor.py:40-44).  Those steps run once, on sample 0, and the recurrent state is replicated to the other
  // samples before the first per-sample step.  VF_SHARED_PREFIX=0 switches it off (A/B, tests).
  h->cur_M = M;
  // the recurrent-state rows of one view as (buffer, row bytes[, snapshot]) entries
  auto state_rows = [&](int v, bool from_snapshot, BroadcastBatch& bb, auto&& flush) {
    auto add = [&](RnnState& r) {
      if (!r.c) return;
      const long long hw = (long long)r.h * r.w;
      const long long cb = hw * r.F * (long long)sizeof(float);
      bb.b[bb.n++] = {r.c, cb, from_snapshot ? r.snap_c : nullptr};
      if (h->split) {                                   // two fp16 planes, the lo plane h->B samples after the hi plane
        const long long pb = hw * 2 * r.F * 2;
        char* snap = reinterpret_cast<char*>(r.snap_in);
        bb.b[bb.n++] = {r.lstm_in, pb, from_snapshot ? snap : nullptr};
        bb.b[bb.n++] = {reinterpret_cast<__half*>(r.lstm_in) + (long long)h->B * hw * 2 * r.F, pb, from_snapshot ? snap + pb : nullptr};
      } else {
        bb.b[bb.n++] = {r.lstm_in, hw * 2 * r.F * (long long)sizeof(float), from_snapshot ? r.snap_in : nullptr};
      }
      if (bb.n > 20) flush();
    };
    for (auto& r : h->views[v].enc_rnn) add(r);
    for (auto& r : h->views[v].dec_rnn) add(r);
    flush();
  };
  const int nsteps = h->S - 1;
  const int first_tau = restore ? n_shared : 0;              // first cell step this rollout actually launches
  const long long sa_step = (long long)h->B * h->A;          // h->sa holds [S-1][B][A] (one view at a time when hoisted per view below)
  for (int tau = 0; tau < nsteps; ++tau) {
    const bool shared = tau < n_shared;
    h->cur_tau = tau;
    if (shared && restore) {                         // prefix cached by iteration 0 of this plan
      if (tau == n_shared - 1)
        for (int v = 0; v < h->ncam; ++v) {
          BroadcastBatch bb;
          bb.n = 0;
          state_rows(v, true, bb, [&]() { launch_broadcast_rows(bb, M, h->stream); bb.n = 0; });
        }
      continue;
    }
    for (int v = 0; v < h->ncam; ++v) {
      // every view runs its own state recurrence with its own state head; the states returned to the
      // caller are view 0's (vpred_model_interface.py:80-82 reads outputs['gen_states'] of the first model)
      sa.w_state = h->views[v].w_state; sa.b_state = h->views[v].b_state;
      sa.w_z = h->views[v].w_z; sa.b_z = h->views[v].b_z; sa.zstate = h->views[v].zstate;
      sa.state_cur = h->sdim ? h->views[v].state_cur : h->state_cur;
      sa.gen_states_all = (h->sdim && v == 0) ? h->gen_states : nullptr;
      if (!h->opt.hoist_sa) {
        launch_build_sa(sa, M, tau, h->stream);
      } else if (tau == first_tau) {
        // every step's vector and every layer's border-class bias for this view: two launches per rollout instead of two per step
        sa.sa = h->sa_all + (long long)v * nsteps * sa_step;
        launch_build_sa_all(sa, M, nsteps, sa_step, h->stream);
        launch_step_sabias(h, v, M, nsteps, sa.sa, sa_step);
      }
      run_step(h, v, tau, shared ? 1 : M);
      if (shared && tau == n_shared - 1) {
        BroadcastBatch bb;
        bb.n = 0;
        if (mode == PREFIX_SAVE) {                     // keep sample 0's rows for the later iterations of this plan
          state_rows(v, true, bb, [&]() {
            for (int i = 0; i < bb.n; ++i)
              cudaMemcpyAsync(const_cast<void*>(bb.b[i].src), bb.b[i].p, (size_t)bb.b[i].row_bytes, cudaMemcpyDeviceToDevice, h->stream);
            bb.n = 0;
          });
        }
        state_rows(v, false, bb, [&]() { launch_broadcast_rows(bb, M, h->stream); bb.n = 0; });
      }
    }
  }
  return VF_OK;
}

// rolls S-1 cell steps for M samples whose actions are in h->actions [M][T][adim].  The launch sequence of a rollout is
// static for a given (prefix mode, shared steps, M, T, n_ctx_actions), so it is captured into a CUDA graph the second time
// a key is seen and replayed afterwards (~2100 kernel launches per plan collapse into 3 graph launches: one PREFIX_SAVE
// rollout and two PREFIX_RESTORE rollouts).
int rollout(vf_engine* h, int M, int T, int mode = PREFIX_NONE) {
  if (!h->weights_ready) { int r = finalize_weights(h); if (r) return r; }
  if (!h->context_set) return fail(h, VF_ERR_STATE, "vf_set_context must be called before predicting");
  if (!h->distrib_set) return fail(h, VF_ERR_STATE, "no designated-pixel distribution: pass pix_distrib to vf_set_context or call vf_set_desig");
  const int need = h->S - 1 - h->n_ctx_actions;
  if (T < need) return fail(h, VF_ERR_INVALID, "need %d actions per sample (S-1-n_ctx_actions), got T=%d", need, T);
  if (shared_prefix_steps(h, M) == 0) mode = PREFIX_NONE;
  if (mode == PREFIX_RESTORE && h->snap_version != h->ctx_version) mode = PREFIX_SAVE;     // no snapshot of this context yet
  const long long key = ((long long)shared_prefix_steps(h, M) << 56) | ((long long)mode << 52) | ((long long)M << 24) | ((long long)T << 8) |
                        (long long)h->n_ctx_actions;
  if (h->graphs.size() > 24 && !h->graphs.count(key)) {       // a handful of (mode, M, T) keys per controller: bound the cache
    CU(cudaStreamSynchronize(h->stream));
    drop_graphs(h);
  }
  vf_engine::GraphSlot& gs = h->graphs[key];
  int r = VF_OK;
  if (!h->use_graph || h->prof_on) {
    r = rollout_body(h, M, T, mode);
  } else if (gs.exec) {
    CU(cudaGraphLaunch(gs.exec, h->stream));
    g_launch_counter += gs.launches;
  } else if (!gs.seen) {
    gs.seen = true;                                // first sighting: run eagerly (one-time attribute setup, JIT-free warm-up)
    r = rollout_body(h, M, T, mode);
  } else {
    cudaGraph_t graph = nullptr;
    const long long before = g_launch_counter;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    r = rollout_body(h, M, T, mode);
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (r) { if (graph) cudaGraphDestroy(graph); return r; }
    if (e != cudaSuccess) return fail(h, VF_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    gs.launches = g_launch_counter - before;
    e = cudaGraphInstantiate(&gs.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { gs.exec = nullptr; return fail(h, VF_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
    CU(cudaGraphLaunch(gs.exec, h->stream));
  }
  if (r) return r;
  if (mode == PREFIX_SAVE) h->snap_version = h->ctx_version;
  if (h->conv_error) {
    const int ce = h->conv_error;
    h->conv_error = 0;
    return fail(h, VF_ERR_CUDA, "tensor-core convolution %s could not be launched (code %d)", h->conv_error_layer, ce);
  }
  CU(cudaGetLastError());
  h->predicted = true;
  h->last_M = M;
  return VF_OK;
}


int ensure_actions(vf_engine* h, int T) {
  if (T > h->Tcap) {
    // captured rollout graphs hold the old pointer by value (SaArgs): they must not survive the reallocation
    CU(cudaStreamSynchronize(h->stream));
    drop_graphs(h);
    dfree(h, h->actions);
    h->actions = nullptr;
    DA(h->actions, (size_t)h->B * T * h->adim);
    h->Tcap = T;
  }
  return VF_OK;
}

int score_device(vf_engine* h, int cost_kind, const float* goal, const float* task_weights, float finalweight,
                 const float* distrib, int M, int P, double* scores_dev) {
  const int ntask = h->ncam * h->nd;
  if (cost_kind == VF_COST_PIXEL_DISTANCE) {
    double g[VF_MAX_TASKS * 2], tw[VF_MAX_TASKS];
    for (int i = 0; i < ntask * 2; ++i) g[i] = (double)goal[i];
    for (int i = 0; i < ntask; ++i) tw[i] = task_weights ? (double)task_weights[i] : 1.0 / ntask;
    // pageable sources: cudaMemcpyAsync returns once they are staged, so the stack arrays may go out of scope (no stream sync)
    CU(cudaMemcpyAsync(h->goal_dev, g, sizeof(double) * ntask * 2, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->taskw_dev, tw, sizeof(double) * ntask, cudaMemcpyHostToDevice, h->stream));
    launch_pixel_cost(distrib, M, P, h->ncam, h->H, h->W, h->nd, h->goal_dev, h->cost, h->stream);
    launch_score_final(h->cost, M, P, ntask, h->taskw_dev, (double)finalweight, scores_dev, h->stream);
  } else if (cost_kind == VF_COST_GOAL_IMAGE) {
    const size_t n = (size_t)h->ncam * h->H * h->W * 3;
    CU(cudaMemcpyAsync(h->goal_img, goal, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    launch_goal_image_cost(h->gen_images, M, P, h->ncam, h->H, h->W, h->goal_img, scores_dev, h->stream);
  } else {
    return fail(h, VF_ERR_INVALID, "unknown cost kind %d", cost_kind);
  }
  CU(cudaGetLastError());
  return VF_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int vf_abi_version(void) { return VF_ABI_VERSION; }

const char* vf_last_error(const vf_engine* h) { return h ? h->err : "null handle"; }

int vf_create(const vf_config* cfg, vf_engine** out) {
  if (!cfg || !out) return VF_ERR_INVALID;
  *out = nullptr;
  vf_engine* h = new vf_engine();
  h->err[0] = 0;
  h->cfg = *cfg;
  *out = h;   // returned even on failure so the caller can read vf_last_error, then vf_destroy
  if (cfg->abi_version != VF_ABI_VERSION) return fail(h, VF_ERR_INVALID, "ABI version %d != %d", cfg->abi_version, VF_ABI_VERSION);
  h->B = cfg->max_samples; h->H = cfg->height; h->W = cfg->width; h->ncam = cfg->ncam; h->nd = cfg->ndesig;
  h->adim = cfg->adim; h->sdim = cfg->sdim; h->nz = cfg->nz; h->A = cfg->adim + cfg->sdim + cfg->nz;
  h->S = cfg->seq_len; h->C = cfg->context_frames; h->P = h->S - h->C; h->ngf = cfg->ngf;
  h->nt = cfg->num_transformed; h->kc = cfg->cdna_ksize; h->nm = h->nt + 3; h->n_enc = cfg->n_enc;
  h->cl = (3 * h->nm + 7) / 8 * 8;
  h->cm = cfg->ngf + h->cl;
  h->split = cfg->precision != VF_PREC_FP32_SIMT;
  { const char* e = getenv("VF_NO_GRAPH"); h->use_graph = !(e && e[0] == '1'); }
  {
    auto flag = [](const char* name, bool dflt) { const char* e = getenv(name); return e && e[0] ? atoi(e) != 0 : dflt; };
    h->opt.shared_prefix = flag("VF_SHARED_PREFIX", true);
    h->opt.prefix_cache = flag("VF_PREFIX_CACHE", true);
    h->opt.stats_fin = flag("VF_STATS_FIN", true);
    h->opt.epi_stats = flag("VF_EPI_STATS", true);
    h->opt.merge_heads = flag("VF_MERGE_HEADS", true) && cfg->precision != VF_PREC_FP32_SIMT;
    h->opt.fuse_fin = flag("VF_FUSE_FIN", false) && h->opt.stats_fin;
    h->opt.hoist_sa = flag("VF_HOIST_SA", true);
    h->opt.side_cdna = flag("VF_SIDE_CDNA", false);
    h->opt.lstm_fused = flag("VF_LSTM_FUSED", false) && h->opt.stats_fin;
  }
  // programmatic dependent launch: measured SLOWER on B200 inside the replayed graph (131.4 vs 124.7 ms per plan), so opt-in
  { const char* e = getenv("VF_PDL"); g_use_pdl = e && e[0] == '1'; }
  if (h->B < 1 || h->H < 8 || h->W < 8 || h->ncam < 1 || h->ncam > 4 || h->nd < 1 || h->nd > 4 || h->ncam * h->nd > VF_MAX_TASKS)
    return fail(h, VF_ERR_INVALID, "bad sizes (max_samples %d, %dx%d, ncam %d, ndesig %d)", h->B, h->H, h->W, h->ncam, h->nd);
  if (h->adim < 1 || h->adim > 8 || h->sdim < 0 || h->sdim > 16 || h->nz < 0 || h->A > 24)
    return fail(h, VF_ERR_INVALID, "bad adim/sdim/nz (adim <= 8, sdim <= 16, adim + sdim + nz <= 24)");
  if (h->C < 1 || h->S <= h->C) return fail(h, VF_ERR_INVALID, "need seq_len > context_frames >= 1");
  if (cfg->n_enc < 1 || cfg->n_enc > VF_MAX_LAYERS || cfg->n_dec != cfg->n_enc) return fail(h, VF_ERR_INVALID, "n_enc/n_dec");
  if ((h->H % (1 << cfg->n_enc)) || (h->W % (1 << cfg->n_enc))) return fail(h, VF_ERR_INVALID, "H, W must be divisible by 2^n_enc");
  if ((h->H >> cfg->n_enc) < 4 || (h->W >> cfg->n_enc) < 4) return fail(h, VF_ERR_INVALID, "coarsest feature map must be at least 4x4");
  if (cfg->dec_channels[cfg->n_dec - 1] != cfg->ngf) return fail(h, VF_ERR_INVALID, "last decoder layer must have ngf channels");
  for (int i = 0; i < cfg->n_enc; ++i)
    if ((cfg->enc_channels[i] % 4) || (cfg->dec_channels[i] % 4) || cfg->enc_channels[i] > 256 || cfg->dec_channels[i] > 256 || (cfg->ngf % 4))
      return fail(h, VF_ERR_INVALID, "layer widths must be multiples of 4 and at most 256 (vectorised statistics / conv-LSTM kernels)");
  if (!cfg->enc_rnn[cfg->n_enc - 1]) return fail(h, VF_ERR_INVALID, "the last encoder layer must be recurrent (CDNA feature)");
  if (h->nt < 1 || h->nt > 8 || h->kc * h->kc * h->nt > 128 || (h->kc % 2) == 0 || (cfg->lstm_ksize != 5 && cfg->lstm_ksize != 3))
    return fail(h, VF_ERR_INVALID, "unsupported CDNA/LSTM kernel configuration");
  if (cfg->precision < 0 || cfg->precision > VF_PREC_F16X1) return fail(h, VF_ERR_INVALID, "bad precision");
  if (cfg->rnn_z && (h->nz < 1 || h->nz > 16)) return fail(h, VF_ERR_INVALID, "rnn_z needs 1 <= nz <= 16");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(h, VF_ERR_CUDA, "no CUDA device available (%s): vfengine has no CPU path", cudaGetErrorString(e));
  CU(cudaSetDevice(cfg->device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  CU(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  int r = build_net(h);
  if (r) return r;
  const size_t px = (size_t)h->H * h->W;
  DA(h->ctx_u8, (size_t)h->C * h->ncam * px * 3);
  DA(h->ctx_frames, (size_t)h->C * h->ncam * px * 3);
  DA(h->ctx_distrib, (size_t)h->C * h->ncam * px * h->nd);
  DA(h->ctx_states, (size_t)h->C * std::max(h->sdim, 1));
  DA(h->ctx_actions, (size_t)std::max(h->S, 1) * h->adim);
  DA(h->desig_pix_dev, (size_t)h->ncam * h->nd * 2);
  DA(h->gen_images, (size_t)h->B * h->P * h->ncam * px * 3);
  DA(h->gen_distrib, (size_t)h->B * h->P * h->ncam * px * h->nd);
  DA(h->gen_states, (size_t)h->B * h->P * std::max(h->sdim, 1));
  DA(h->sa, (size_t)h->B * h->A);
  if (h->opt.hoist_sa) DA(h->sa_all, (size_t)h->ncam * (h->S - 1) * h->B * h->A);
  DA(h->state_cur, (size_t)h->B * std::max(h->sdim, 1));
  if (h->nz) DA(h->zs, (size_t)h->B * (h->S - 1) * h->nz);
  DA(h->cost, (size_t)h->B * h->P * h->ncam * h->nd);
  DA(h->goal_dev, VF_MAX_TASKS * 2);
  DA(h->taskw_dev, VF_MAX_TASKS);
  DA(h->goal_img, (size_t)h->ncam * px * 3);
  DA(h->scores_tmp, (size_t)h->B);
  DA(h->fetch_idx, (size_t)h->B);
  return VF_OK;
}

int vf_destroy(vf_engine* h) {
  if (!h) return VF_OK;
  if (h->stream) cudaStreamSynchronize(h->stream);
  vf_comm_close(h);
  for (auto& kv : h->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : h->allocs) cudaFree(p);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
  return VF_OK;
}

int vf_set_stream(vf_engine* h, void* s) {
  if (!h) return VF_ERR_INVALID;
  if (h->stream) CU(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (s) { h->stream = (cudaStream_t)s; h->own_stream = false; }
  else { CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  return VF_OK;
}

int vf_synchronize(vf_engine* h) {
  if (!h) return VF_ERR_INVALID;
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_load_weights(vf_engine* h, const vf_tensor* t, int32_t n) {
  if (!h || (!t && n > 0)) return VF_ERR_INVALID;
  if (h->weights_ready) return fail(h, VF_ERR_STATE, "weights already finalised for this handle");
  for (int i = 0; i < n; ++i) {
    if (!t[i].name || !t[i].data || t[i].dtype != VF_F32 || t[i].ndim < 1 || t[i].ndim > 6)
      return fail(h, VF_ERR_INVALID, "tensor %d is malformed", i);
    HostTensor ht;
    size_t cnt = 1;
    for (int d = 0; d < t[i].ndim; ++d) { ht.shape.push_back(t[i].shape[d]); cnt *= (size_t)t[i].shape[d]; }
    ht.data.assign((const float*)t[i].data, (const float*)t[i].data + cnt);
    h->host_w[t[i].name] = std::move(ht);
  }
  return VF_OK;
}

int vf_set_context(vf_engine* h, const uint8_t* frames, const float* states, const float* ctx_actions, int32_t nca,
                   const float* pix_distrib) {
  if (!h || !frames) return fail(h, VF_ERR_INVALID, "frames_u8 is required");
  if (h->sdim > 0 && !states) return fail(h, VF_ERR_INVALID, "states required when sdim > 0");
  if (nca < 0 || nca > h->S - 1 || (nca > 0 && !ctx_actions)) return fail(h, VF_ERR_INVALID, "bad n_ctx_actions %d", nca);
  const size_t px = (size_t)h->H * h->W, nf = (size_t)h->C * h->ncam * px * 3;
  CU(cudaMemcpyAsync(h->ctx_u8, frames, nf, cudaMemcpyHostToDevice, h->stream));
  launch_u8_to_f32(h->ctx_u8, h->ctx_frames, (long long)nf, 255.0f, h->stream);
  if (h->sdim > 0) CU(cudaMemcpyAsync(h->ctx_states, states, sizeof(float) * h->C * h->sdim, cudaMemcpyHostToDevice, h->stream));
  if (nca > 0) CU(cudaMemcpyAsync(h->ctx_actions, ctx_actions, sizeof(float) * nca * h->adim, cudaMemcpyHostToDevice, h->stream));
  h->n_ctx_actions = nca;
  if (pix_distrib) {
    CU(cudaMemcpyAsync(h->ctx_distrib, pix_distrib, sizeof(float) * h->C * h->ncam * px * h->nd, cudaMemcpyHostToDevice, h->stream));
    h->distrib_set = true;
  }
  CU(cudaStreamSynchronize(h->stream));   // caller buffers may be pageable / short-lived
  h->context_set = true;
  ++h->ctx_version;                                  // invalidates the shared-prefix snapshot
  return VF_OK;
}

int vf_set_desig(vf_engine* h, const float* desig) {
  if (!h || !desig) return VF_ERR_INVALID;
  int pix[VF_MAX_TASKS * 2];
  for (int i = 0; i < h->ncam * h->nd; ++i) {
    // np.clip(desig, 0, [H-1, W-1]).astype(int): truncation toward zero of the clipped float
    float y = desig[2 * i], x = desig[2 * i + 1];
    y = fminf(fmaxf(y, 0.f), (float)(h->H - 1));
    x = fminf(fmaxf(x, 0.f), (float)(h->W - 1));
    pix[2 * i] = (int)y;
    pix[2 * i + 1] = (int)x;
  }
  const size_t n = (size_t)h->C * h->ncam * h->H * h->W * h->nd;
  CU(cudaMemsetAsync(h->ctx_distrib, 0, n * sizeof(float), h->stream));
  CU(cudaMemcpyAsync(h->desig_pix_dev, pix, sizeof(int) * h->ncam * h->nd * 2, cudaMemcpyHostToDevice, h->stream));
  launch_onehot(h->ctx_distrib, h->C, h->ncam, h->H, h->W, h->nd, h->desig_pix_dev, h->stream);
  CU(cudaStreamSynchronize(h->stream));
  h->distrib_set = true;
  return VF_OK;
}

int vf_predict(vf_engine* h, const float* actions, int32_t M, int32_t T, const float* zs, float* of, float* od, float* os) {
  if (!h || !actions) return fail(h, VF_ERR_INVALID, "actions is required");
  if (M < 1 || M > h->B) return fail(h, VF_ERR_INVALID, "M=%d outside [1, max_samples=%d]", M, h->B);
  if (h->nz > 0 && !zs) return fail(h, VF_ERR_INVALID, "zs required when nz > 0");
  int r = ensure_actions(h, T);
  if (r) return r;
  CU(cudaMemcpyAsync(h->actions, actions, sizeof(float) * (size_t)M * T * h->adim, cudaMemcpyHostToDevice, h->stream));
  if (h->nz) CU(cudaMemcpyAsync(h->zs, zs, sizeof(float) * (size_t)M * (h->S - 1) * h->nz, cudaMemcpyHostToDevice, h->stream));
  h->T = T;
  r = rollout(h, M, T);
  if (r) return r;
  const size_t px = (size_t)h->H * h->W;
  if (of) CU(cudaMemcpyAsync(of, h->gen_images, sizeof(float) * (size_t)M * h->P * h->ncam * px * 3, cudaMemcpyDeviceToHost, h->stream));
  if (od) CU(cudaMemcpyAsync(od, h->gen_distrib, sizeof(float) * (size_t)M * h->P * h->ncam * px * h->nd, cudaMemcpyDeviceToHost, h->stream));
  if (os && h->sdim) CU(cudaMemcpyAsync(os, h->gen_states, sizeof(float) * (size_t)M * h->P * h->sdim, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_score(vf_engine* h, int32_t kind, const float* goal, const float* tw, float fw, double* out) {
  if (!h || !goal || !out) return fail(h, VF_ERR_INVALID, "goal and out_scores are required");
  if (!h->predicted) return fail(h, VF_ERR_STATE, "vf_score before vf_predict");
  int r = score_device(h, kind, goal, tw, fw, h->gen_distrib, h->last_M, h->P, h->scores_tmp);
  if (r) return r;
  CU(cudaMemcpyAsync(out, h->scores_tmp, sizeof(double) * h->last_M, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_score_external(vf_engine* h, const float* distrib, int32_t M, int32_t P, const float* goal, const float* tw,
                      float fw, double* out) {
  if (!h || !distrib || !goal || !out) return fail(h, VF_ERR_INVALID, "null argument");
  if (M < 1 || P < 1 || M > h->B || (long long)M * P > (long long)h->B * h->P) return fail(h, VF_ERR_INVALID, "M*P exceeds capacity");
  const size_t n = (size_t)M * P * h->ncam * h->H * h->W * h->nd;
  CU(cudaMemcpyAsync(h->gen_distrib, distrib, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  h->predicted = false;   // gen_distrib no longer holds a rollout
  double* sc = h->scores_tmp;
  int r = score_device(h, VF_COST_PIXEL_DISTANCE, goal, tw, fw, h->gen_distrib, M, P, sc);
  if (r) return r;
  CU(cudaMemcpyAsync(out, sc, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_fetch(vf_engine* h, const int32_t* idx, int32_t n, float* of, float* od) {
  if (!h || !idx || n < 1) return VF_ERR_INVALID;
  if (!h->predicted) return fail(h, VF_ERR_STATE, "vf_fetch before a rollout");
  const size_t px = (size_t)h->H * h->W;
  const size_t rf = (size_t)h->P * h->ncam * px * 3, rd = (size_t)h->P * h->ncam * px * h->nd;
  for (int i = 0; i < n; ++i) {
    if (idx[i] < 0 || idx[i] >= h->last_M) return fail(h, VF_ERR_INVALID, "index %d out of range", idx[i]);
    if (of) CU(cudaMemcpyAsync(of + (size_t)i * rf, h->gen_images + (size_t)idx[i] * rf, rf * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (od) CU(cudaMemcpyAsync(od + (size_t)i * rd, h->gen_distrib + (size_t)idx[i] * rd, rd * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

// ---- CEM -----------------------------------------------------------------------------------------

static void fill_sample_args(vf_engine* h, SampleArgs& a, int iteration) {
  const vf_cem_params& p = h->cem;
  a.D = h->cem_D; a.nactions = p.nactions; a.adim = h->adim - p.n_append; a.repeat = p.repeat;
  a.adim_out = h->adim;
  for (int i = 0; i < 8; ++i) { a.append[i] = i < p.n_append ? p.append_action[i] : 0.0; a.bias[i] = p.mean_bias[i]; }
  a.discrete_mask = p.discrete_mask;
  a.kind = p.sampler; a.beta0 = p.beta0; a.beta1 = p.beta1;
  a.K = (iteration == 0 || p.sampler == VF_SAMPLER_CORRELATED) ? 0 : p.num_elites;
  a.mean = h->cem_mean; a.factor = h->cem_factor; a.std0 = h->cem_std0;
  a.noise = h->cem_has_noise ? h->cem_noise + (size_t)iteration * p.global_samples * h->cem_Dmax : nullptr;
  a.noise_stride = h->cem_Dmax;
  a.indices = nullptr; a.offset = p.sample_offset;
  for (int i = 0; i < 8; ++i) {
    a.clip_lo[i] = p.action_bound ? (double)p.clip_lo[i] : -INFINITY;
    a.clip_hi[i] = p.action_bound ? (double)p.clip_hi[i] : INFINITY;
  }
  a.seed = p.seed; a.plan_index = p.plan_index; a.iteration = (uint32_t)iteration;
  a.out_nr = nullptr; a.out_actions = nullptr; a.out_actions64 = nullptr;
}

int vf_cem_begin(vf_engine* h, const vf_cem_params* p, const float* goal, const float* noise) {
  if (!h || !p || !goal) return fail(h, VF_ERR_INVALID, "params and goal are required");
  const int M = p->num_samples, Mg = p->global_samples, K = p->num_elites;
  const int Kf = p->k_futures > 1 ? p->k_futures : 1;
  if (M < 1 || (long long)M * Kf > h->B)
    return fail(h, VF_ERR_INVALID, "num_samples=%d x k_futures=%d outside [1, max_samples=%d]", M, Kf, h->B);
  if (Mg < M || p->sample_offset < 0 || p->sample_offset + M > Mg) return fail(h, VF_ERR_INVALID, "bad shard (offset %d, M %d, global %d)", p->sample_offset, M, Mg);
  if (K < 1 || K > Mg) return fail(h, VF_ERR_INVALID, "num_elites=%d outside [1, %d]", K, Mg);
  if (p->iterations < 1 || p->nactions < 1 || p->repeat < 1) return fail(h, VF_ERR_INVALID, "bad iterations/nactions/repeat");
  if (p->n_append < 0 || p->n_append >= h->adim) return fail(h, VF_ERR_INVALID, "n_append=%d outside [0, adim=%d)", p->n_append, h->adim);
  if (p->sampler != VF_SAMPLER_GAUSSIAN && p->sampler != VF_SAMPLER_CORRELATED) return fail(h, VF_ERR_INVALID, "unknown sampler %d", p->sampler);
  if (p->sampler == VF_SAMPLER_CORRELATED && (p->repeat != 1 || p->use_mean0))
    return fail(h, VF_ERR_INVALID, "the correlated-noise sampler has repeat = 1 and no warm start (correlated_noise.py:37-45)");
  const int sdims = h->adim - p->n_append;            // dims the sampler draws per step
  const int D = p->nactions * sdims;
  if (D > 128) return fail(h, VF_ERR_INVALID, "nactions*(adim - n_append)=%d > 128", D);
  if (p->n_ctx_actions != h->n_ctx_actions) return fail(h, VF_ERR_INVALID, "n_ctx_actions (%d) differs from vf_set_context (%d)", p->n_ctx_actions, h->n_ctx_actions);
  h->cem = *p;
  h->cem_D = D; h->cem_T = p->nactions * p->repeat; h->cem_Dmax = p->sampler == VF_SAMPLER_CORRELATED ? D : std::max(D, K);
  int r = ensure_actions(h, h->cem_T);
  if (r) return r;
  if (!h->cem_mean) {
    DA(h->cem_mean, 128); DA(h->cem_std0, 128); DA(h->cem_cov, 128 * 128);
  }
  // per-call sized buffers (grow only)
  const size_t need_scores = (size_t)p->iterations * Mg;
  if (need_scores > h->cem_scores_cap) {
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, h->cem_scores_own);
    h->cem_scores_own = nullptr;
    DA(h->cem_scores_own, need_scores);
    h->cem_scores_cap = need_scores;
  }
  h->cem_scores = h->cem_scores_ext ? h->cem_scores_ext : h->cem_scores_own;
  h->cem_scores_final = h->cem_scores;
  if (h->comm.connected && Mg > M) {
    // sharded plan over the engine's peer exchange: rollouts write into the window, vf_cem_iter_select keeps a private copy of
    // every row it consumed (a fast peer may already be storing the NEXT plan's scores into the window when vf_cem_finish reads)
    if (p->iterations > h->comm.cap_iters || (size_t)p->iterations * Mg > (size_t)h->comm.cap_iters * h->comm.cap_global)
      return fail(h, VF_ERR_INVALID, "plan (%d iterations x %d samples) exceeds the exchange window (%d x %d)", p->iterations, Mg,
                  h->comm.cap_iters, h->comm.cap_global);
    h->cem_scores = comm_scores(h->comm.window);
    h->cem_scores_final = h->cem_scores_own;
  }
  const size_t npad = (size_t)topk_padded(Mg);
  if (npad > h->topk_cap) {
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, h->topk_keys); dfree(h, h->topk_idx);
    h->topk_keys = nullptr; h->topk_idx = nullptr;
    DA(h->topk_keys, npad); DA(h->topk_idx, npad); h->topk_cap = npad;
  }
  if (Mg > h->cem_Mg_cap) {                          // elite buffers hold up to K <= Mg rows: sized by Mg itself, not by its padding
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, h->cem_elite_idx); dfree(h, h->cem_factor); dfree(h, h->cem_elites_nr); dfree(h, h->cem_best64);
    h->cem_elite_idx = nullptr; h->cem_factor = nullptr; h->cem_elites_nr = nullptr; h->cem_best64 = nullptr;
    DA(h->cem_elite_idx, Mg);
    DA(h->cem_factor, (size_t)128 * Mg); DA(h->cem_elites_nr, (size_t)Mg * 128);
    DA(h->cem_best64, (size_t)Mg * 128 * 8);
    h->cem_Mg_cap = Mg;
  }
  if (!h->cem_local_nr) { DA(h->cem_local_nr, (size_t)h->B * 128); }
  if (h->cem_T > h->cem_act_cap) {
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, h->cem_actions64);
    h->cem_actions64 = nullptr;
    DA(h->cem_actions64, (size_t)h->B * h->cem_T * h->adim);
    h->cem_act_cap = h->cem_T;
  }
  if ((size_t)K * h->cem_T * h->adim > (size_t)h->cem_Mg_cap * 128 * 8) return fail(h, VF_ERR_INVALID, "elite buffer too small");
  h->cem_Kf = Kf;
  if (!h->cem_fut_idx) { DA(h->cem_fut_idx, (size_t)h->B); DA(h->cem_fut_scores, (size_t)h->B); }
  if (Kf > 1) {                                     // rollout sample r evaluates action sequence offset + r / Kf
    std::vector<int> fi((size_t)M * Kf);
    for (int r = 0; r < M * Kf; ++r) fi[r] = p->sample_offset + r / Kf;
    CU(cudaMemcpyAsync(h->cem_fut_idx, fi.data(), sizeof(int) * fi.size(), cudaMemcpyHostToDevice, h->stream));   // pageable: staged on return
  }
  h->cem_has_noise = noise != nullptr;
  if (noise) {
    const size_t nn = (size_t)p->iterations * Mg * h->cem_Dmax;
    if (nn > h->cem_noise_cap) {
      CU(cudaStreamSynchronize(h->stream));
      dfree(h, h->cem_noise);
      h->cem_noise = nullptr;
      DA(h->cem_noise, nn);
      h->cem_noise_cap = nn;
    }
    CU(cudaMemcpyAsync(h->cem_noise, noise, nn * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  double mean[128], std0[128];
  for (int d = 0; d < D; ++d) {
    mean[d] = p->use_mean0 ? (double)p->mean0[d] : 0.0;
    double s = (double)p->initial_std[d % sdims];
    // construct_initial_sigma scales the VARIANCE of all but the last action block (controller_utils.py:76-81)
    if (p->reduce_std_scale != 1.0 && d < (p->nactions - 1) * sdims) s *= sqrt(p->reduce_std_scale);
    std0[d] = s;
  }
  CU(cudaMemcpyAsync(h->cem_mean, mean, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->cem_std0, std0, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream));
  // goal / task weights
  const int ntask = h->ncam * h->nd;
  if (p->cost_kind == VF_COST_PIXEL_DISTANCE) {
    double g[VF_MAX_TASKS * 2], tw[VF_MAX_TASKS];
    double wsum = 0;
    for (int i = 0; i < ntask; ++i) wsum += p->task_weights[i];
    for (int i = 0; i < ntask * 2; ++i) g[i] = (double)goal[i];
    for (int i = 0; i < ntask; ++i) tw[i] = wsum > 0 ? (double)p->task_weights[i] : 1.0 / ntask;
    CU(cudaMemcpyAsync(h->goal_dev, g, sizeof(double) * ntask * 2, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->taskw_dev, tw, sizeof(double) * ntask, cudaMemcpyHostToDevice, h->stream));
  } else if (p->cost_kind == VF_COST_GOAL_IMAGE) {
    CU(cudaMemcpyAsync(h->goal_img, goal, sizeof(float) * (size_t)h->ncam * h->H * h->W * 3, cudaMemcpyHostToDevice, h->stream));
  } else {
    return fail(h, VF_ERR_INVALID, "unknown cost kind");
  }
  // no stream synchronisation: every source above is pageable host memory, which cudaMemcpyAsync has staged when it returns
  h->cem_active = true;
  return VF_OK;
}

int vf_cem_iter_rollout(vf_engine* h, int32_t it) {
  if (!h || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  const vf_cem_params& p = h->cem;
  if (it < 0 || it >= p.iterations) return fail(h, VF_ERR_INVALID, "iteration %d out of range", it);
  const int Kf = h->cem_Kf, nroll = p.num_samples * Kf;
  SampleArgs a;
  fill_sample_args(h, a, it);
  if (Kf > 1) a.indices = h->cem_fut_idx;           // consecutive copies of every action sequence (np.repeat order)
  a.out_nr = h->cem_local_nr; a.out_actions = h->actions; a.out_actions64 = h->cem_actions64;
  launch_sample_actions(a, nroll, h->stream);
  if (h->nz > 0)                                    // latents of the stochastic predictor: Philox, keyed by the global rollout index
    launch_sample_latents(h->zs, nroll, h->S - 1, h->nz, p.sample_offset * Kf, p.seed, p.plan_index, (uint32_t)it, h->stream);
  h->T = h->cem_T;
  const bool cache_env = h->opt.prefix_cache;
  int r = rollout(h, nroll, h->cem_T, !cache_env ? PREFIX_NONE : (it == 0 ? PREFIX_SAVE : PREFIX_RESTORE));
  if (r) return r;
  double* sc = h->cem_scores + (size_t)it * p.global_samples + p.sample_offset;
  double* raw_sc = Kf > 1 ? h->cem_fut_scores : sc;
  const int ntask = h->ncam * h->nd;
  if (p.cost_kind == VF_COST_PIXEL_DISTANCE) {
    launch_pixel_cost(h->gen_distrib, nroll, h->P, h->ncam, h->H, h->W, h->nd, h->goal_dev, h->cost, h->stream);
    launch_score_final(h->cost, nroll, h->P, ntask, h->taskw_dev, (double)p.finalweight, raw_sc, h->stream);
  } else {
    launch_goal_image_cost(h->gen_images, nroll, h->P, h->ncam, h->H, h->W, h->goal_img, raw_sc, h->stream);
  }
  if (Kf > 1) launch_reduce_futures(raw_sc, p.num_samples, Kf, (double)p.lambda_variance, sc, h->stream);
  CU(cudaGetLastError());
  return VF_OK;
}

int vf_cem_iter_select(vf_engine* h, int32_t it) {
  if (!h || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  const vf_cem_params& p = h->cem;
  if (it < 0 || it >= p.iterations) return fail(h, VF_ERR_INVALID, "iteration %d out of range", it);
  const int K = p.num_elites;
  launch_topk(h->cem_scores + (size_t)it * p.global_samples, p.global_samples, K, h->cem_elite_idx, h->topk_keys, h->topk_idx, h->stream);
  if (h->cem_scores_final != h->cem_scores)
    CU(cudaMemcpyAsync(h->cem_scores_final + (size_t)it * p.global_samples, h->cem_scores + (size_t)it * p.global_samples,
                       sizeof(double) * p.global_samples, cudaMemcpyDeviceToDevice, h->stream));
  // regenerate the elites' action rows from their global indices (no exchange of actions between ranks)
  SampleArgs a;
  fill_sample_args(h, a, it);
  a.indices = h->cem_elite_idx;
  a.out_nr = h->cem_elites_nr; a.out_actions64 = h->cem_best64;
  launch_sample_actions(a, K, h->stream);
  if (it < p.iterations - 1) {
    if (p.sampler == VF_SAMPLER_CORRELATED)
      launch_refit_correlated(h->cem_elites_nr, h->cem_scores + (size_t)it * p.global_samples, h->cem_elite_idx, K, h->cem_D, p.kappa,
                              h->cem_mean, h->stream);
    else
      launch_refit(h->cem_elites_nr, K, h->cem_D, h->cem_mean, h->cem_factor, nullptr, h->stream);
  }
  CU(cudaGetLastError());
  return VF_OK;
}

int vf_cem_finish(vf_engine* h, double* best, int32_t* eidx, double* scores) {
  if (!h || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  const vf_cem_params& p = h->cem;
  if (best) CU(cudaMemcpyAsync(best, h->cem_best64, sizeof(double) * (size_t)p.num_elites * h->cem_T * h->adim, cudaMemcpyDeviceToHost, h->stream));
  if (eidx) CU(cudaMemcpyAsync(eidx, h->cem_elite_idx, sizeof(int) * p.num_elites, cudaMemcpyDeviceToHost, h->stream));
  if (scores) CU(cudaMemcpyAsync(scores, h->cem_scores_final, sizeof(double) * (size_t)p.iterations * p.global_samples, cudaMemcpyDeviceToHost, h->stream));
  unsigned comm_status = 0;
  if (h->comm.connected) CU(cudaMemcpyAsync(&comm_status, comm_status_word(h->comm.window), sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (comm_status) {
    cudaMemsetAsync(comm_status_word(h->comm.window), 0, sizeof(unsigned), h->stream);
    return fail(h, VF_ERR_CUDA, "score exchange timed out after %.1f s waiting for a peer rank (rank %d of %d): the plan's elites are invalid",
                h->comm.timeout_ns * 1e-9, h->comm.rank, h->comm.world);
  }
  return VF_OK;
}

// ---- peer-memory score exchange --------------------------------------------------------------------
namespace {
struct PeerDesc {                      // VF_PEER_DESC_BYTES, plain data
  uint32_t magic, version;
  int64_t pid;
  int32_t device, cap_iters, cap_global, pad;
  uint64_t ptr, bytes;
  cudaIpcMemHandle_t ipc;              // 64 bytes
  unsigned char reserved[VF_PEER_DESC_BYTES - 48 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(PeerDesc) == VF_PEER_DESC_BYTES, "peer descriptor layout");
constexpr uint32_t PEER_MAGIC = 0x76665043u;   // 'vfPC'
}  // namespace

int vf_comm_close(vf_engine* h) {
  if (!h) return VF_ERR_INVALID;
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int r = 0; r < VF_MAX_WORLD; ++r) {
    if (h->comm.peer_ipc[r] && h->comm.peer_base[r]) cudaIpcCloseMemHandle(h->comm.peer_base[r]);
    h->comm.peer_base[r] = nullptr;
    h->comm.peer_ipc[r] = false;
  }
  if (h->comm.window) cudaFree(h->comm.window);
  h->comm = vf_engine::Comm();
  h->cem_active = false;
  return VF_OK;
}

int vf_comm_export(vf_engine* h, int32_t max_iterations, int32_t max_global, void* out_desc) {
  if (!h || !out_desc || max_iterations < 1 || max_global < 1) return fail(h, VF_ERR_INVALID, "bad exchange window size");
  vf_comm_close(h);
  const size_t bytes = COMM_HEADER_BYTES + sizeof(double) * (size_t)max_iterations * max_global;
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMalloc(&h->comm.window, bytes));
  CU(cudaMemset(h->comm.window, 0, bytes));
  h->comm.window_bytes = bytes; h->comm.cap_iters = max_iterations; h->comm.cap_global = max_global;
  { const char* e = getenv("VF_COMM_TIMEOUT_MS"); if (e && atof(e) > 0) h->comm.timeout_ns = (unsigned long long)(atof(e) * 1e6); }
  PeerDesc d;
  memset(&d, 0, sizeof(d));
  d.magic = PEER_MAGIC; d.version = VF_ABI_VERSION; d.pid = (int64_t)getpid(); d.device = h->cfg.device;
  d.cap_iters = max_iterations; d.cap_global = max_global; d.ptr = (uint64_t)(uintptr_t)h->comm.window; d.bytes = bytes;
  CU(cudaIpcGetMemHandle(&d.ipc, h->comm.window));
  memcpy(out_desc, &d, sizeof(d));
  return VF_OK;
}

int vf_comm_connect(vf_engine* h, int32_t rank, int32_t world, const void* descs) {
  if (!h || !descs || world < 1 || world > VF_MAX_WORLD || rank < 0 || rank >= world) return fail(h, VF_ERR_INVALID, "bad rank/world (world <= %d)", VF_MAX_WORLD);
  if (!h->comm.window) return fail(h, VF_ERR_STATE, "vf_comm_export first");
  const PeerDesc* d = reinterpret_cast<const PeerDesc*>(descs);
  CU(cudaSetDevice(h->cfg.device));
  for (int r = 0; r < world; ++r) {
    if (d[r].magic != PEER_MAGIC || d[r].version != VF_ABI_VERSION) return fail(h, VF_ERR_INVALID, "descriptor %d is not a vfengine peer descriptor", r);
    if (d[r].cap_iters != h->comm.cap_iters || d[r].cap_global != h->comm.cap_global)
      return fail(h, VF_ERR_INVALID, "rank %d exported a %d x %d window, this rank %d x %d", r, d[r].cap_iters, d[r].cap_global, h->comm.cap_iters, h->comm.cap_global);
    if (r == rank) {
      if ((uint64_t)(uintptr_t)h->comm.window != d[r].ptr || d[r].pid != (int64_t)getpid()) return fail(h, VF_ERR_INVALID, "descriptor %d is not this handle's", r);
      h->comm.peer_base[r] = h->comm.window;
    } else if (d[r].pid == (int64_t)getpid()) {                 // another handle of this process: plain (peer) pointer
      if (d[r].device != h->cfg.device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, h->cfg.device, d[r].device));
        if (!can) return fail(h, VF_ERR_UNSUPPORTED, "device %d cannot access device %d", h->cfg.device, d[r].device);
        cudaError_t e = cudaDeviceEnablePeerAccess(d[r].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, VF_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
      }
      h->comm.peer_base[r] = (void*)(uintptr_t)d[r].ptr;
    } else {                                                     // another process: map its window (peer access over NVLink)
      void* q = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&q, d[r].ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return fail(h, VF_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
      h->comm.peer_base[r] = q;
      h->comm.peer_ipc[r] = true;
    }
  }
  h->comm.rank = rank; h->comm.world = world; h->comm.connected = world > 1; h->comm.epoch = 0;
  h->cem_active = false;
  return VF_OK;
}

int vf_cem_exchange(vf_engine* h, int32_t it) {
  if (!h || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  const vf_cem_params& p = h->cem;
  if (it < 0 || it >= p.iterations) return fail(h, VF_ERR_INVALID, "iteration %d out of range", it);
  if (p.global_samples == p.num_samples) return VF_OK;            // nothing to exchange
  if (!h->comm.connected) return fail(h, VF_ERR_STATE, "vf_comm_connect first");
  ExchangeArgs a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < h->comm.world; ++r) {
    a.scores[r] = comm_scores(h->comm.peer_base[r]);
    a.flags[r] = reinterpret_cast<unsigned*>(h->comm.peer_base[r]);
  }
  a.status = comm_status_word(h->comm.window);
  a.world = h->comm.world; a.rank = h->comm.rank;
  a.row_off = (long long)it * p.global_samples; a.offset = p.sample_offset; a.local = p.num_samples;
  a.epoch = ++h->comm.epoch;
  a.timeout_ns = h->comm.timeout_ns;
  launch_score_exchange(a, h->stream);
  CU(cudaGetLastError());
  return VF_OK;
}

int vf_cem_scores_dev(vf_engine* h, void** dev) {
  if (!h || !dev || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  *dev = h->cem_scores;
  return VF_OK;
}

int vf_cem_bind_scores(vf_engine* h, void* dev) {
  if (!h) return VF_ERR_INVALID;
  h->cem_scores_ext = (double*)dev;
  h->cem_active = false;            // the binding takes effect at the next vf_cem_begin
  return VF_OK;
}

int vf_cem_scores_read(vf_engine* h, int32_t it, int32_t off, int32_t n, double* out) {
  if (!h || !out || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  if (it < 0 || it >= h->cem.iterations || off < 0 || n < 1 || off + n > h->cem.global_samples) return fail(h, VF_ERR_INVALID, "bad score segment");
  CU(cudaMemcpyAsync(out, h->cem_scores + (size_t)it * h->cem.global_samples + off, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_cem_scores_write(vf_engine* h, int32_t it, int32_t off, int32_t n, const double* in) {
  if (!h || !in || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  if (it < 0 || it >= h->cem.iterations || off < 0 || n < 1 || off + n > h->cem.global_samples) return fail(h, VF_ERR_INVALID, "bad score segment");
  CU(cudaMemcpyAsync(h->cem_scores + (size_t)it * h->cem.global_samples + off, in, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_cem_actions(vf_engine* h, double* out) {
  if (!h || !out || !h->cem_active) return fail(h, VF_ERR_STATE, "vf_cem_begin first");
  // with k_futures > 1 every action row appears k_futures times (rollout order); row r belongs to action r / k_futures
  CU(cudaMemcpyAsync(out, h->cem_actions64, sizeof(double) * (size_t)h->cem.num_samples * h->cem_Kf * h->cem_T * h->adim,
                     cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_cem_plan(vf_engine* h, const vf_cem_params* p, const float* goal, const float* noise, double* best, int32_t* eidx,
                double* scores) {
  int r = vf_cem_begin(h, p, goal, noise);
  if (r) return r;
  if (p->global_samples != p->num_samples)
    return fail(h, VF_ERR_INVALID, "vf_cem_plan is the single-rank form; sharded plans use begin/iter_rollout/<all-gather>/iter_select/finish");
  for (int it = 0; it < p->iterations; ++it) {
    if ((r = vf_cem_iter_rollout(h, it))) return r;
    if ((r = vf_cem_iter_select(h, it))) return r;
  }
  return vf_cem_finish(h, best, eidx, scores);
}

int vf_topk(vf_engine* h, const double* scores, int32_t n, int32_t k, int32_t* out) {
  if (!h || !scores || !out || n < 1 || k < 1 || k > n) return fail(h, VF_ERR_INVALID, "bad topk arguments");
  double *ds, *keys;
  int *idx, *oi;
  const int npad = topk_padded(n);
  Scratch sc;
  if (!sc.get(&ds, n) || !sc.get(&keys, npad) || !sc.get(&idx, npad) || !sc.get(&oi, k)) return fail(h, VF_ERR_NOMEM, "topk scratch");
  CU(cudaMemcpyAsync(ds, scores, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  launch_topk(ds, n, k, oi, keys, idx, h->stream);
  CU(cudaMemcpyAsync(out, oi, sizeof(int) * k, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_refit(vf_engine* h, const double* elites, int32_t K, int32_t nactions, int32_t repeat, int32_t adim, double* om,
             double* oc, double* of) {
  if (!h || !elites || K < 1 || nactions < 1 || repeat < 1 || adim < 1) return fail(h, VF_ERR_INVALID, "bad refit arguments");
  const int D = nactions * adim, T = nactions * repeat;
  if (D > 128) return fail(h, VF_ERR_INVALID, "D > 128");
  // _fit_gaussians keeps the LAST action of every repeat group (gaussian_sampler.py:97-99)
  std::vector<double> nr((size_t)K * D);
  for (int k = 0; k < K; ++k)
    for (int a = 0; a < nactions; ++a)
      for (int d = 0; d < adim; ++d) nr[((size_t)k * nactions + a) * adim + d] = elites[((size_t)k * T + a * repeat + repeat - 1) * adim + d];
  double *dx, *dm, *df, *dc;
  Scratch sc;
  if (!sc.get(&dx, (size_t)K * D) || !sc.get(&dm, D) || !sc.get(&df, (size_t)D * K) || !sc.get(&dc, (size_t)D * D))
    return fail(h, VF_ERR_NOMEM, "refit scratch");
  CU(cudaMemcpyAsync(dx, nr.data(), sizeof(double) * K * D, cudaMemcpyHostToDevice, h->stream));
  launch_refit(dx, K, D, dm, df, dc, h->stream);
  if (om) CU(cudaMemcpyAsync(om, dm, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream));
  if (oc) CU(cudaMemcpyAsync(oc, dc, sizeof(double) * D * D, cudaMemcpyDeviceToHost, h->stream));
  if (of) CU(cudaMemcpyAsync(of, df, sizeof(double) * D * K, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

// ---- debug -----------------------------------------------------------------------------------------

int vf_debug_conv2d(vf_engine* h, int32_t impl, const float* x, const float* w, const float* bias, int32_t B, int32_t H,
                    int32_t W, int32_t Cin, int32_t Cout, int32_t k, float* y) {
  if (!h || !x || !w || !y || (k != 3 && k != 5)) return fail(h, VF_ERR_INVALID, "bad conv arguments");
  float *dx, *dw, *db = nullptr, *dy;
  const size_t nx = (size_t)B * H * W * Cin, nw = (size_t)k * k * Cin * Cout, ny = (size_t)B * H * W * Cout;
  Scratch sc;                                                          // every device buffer of this call is released on return
  if (!sc.get(&dx, nx) || !sc.get(&dw, nw) || !sc.get(&dy, ny)) return fail(h, VF_ERR_NOMEM, "conv scratch");
  CU(cudaMemcpyAsync(dx, x, nx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(dw, w, nw * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  if (bias) {
    if (!sc.get(&db, Cout)) return fail(h, VF_ERR_NOMEM, "conv scratch");
    CU(cudaMemcpyAsync(db, bias, Cout * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  View vx = make_view(dx, (long long)H * W * Cin, Cin, 0, Cin), vy = make_view(dy, (long long)H * W * Cout, Cout, 0, Cout);
  if (impl == VF_PREC_FP32_SIMT) {
    ConvArgs a;
    a.src0 = vx; a.src1 = make_view(nullptr, 0, 0, 0, 0); a.w = dw; a.bias = db; a.sabias = nullptr; a.out = vy;
    a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.k = k; a.act = ACT_NONE;
    launch_conv_simt(a, B, h->stream);
  } else {
    if (!mma_conv_supported(k, Cin, Cout, H, W)) return fail(h, VF_ERR_UNSUPPORTED, "shape not supported by the tcgen05 conv");
    MmaConvWeights mw;
    std::string e;
    std::vector<float> wh(w, w + nw);
    if (mma_conv_prepare_weights(wh.data(), k, k, k, Cin, Cout, &mw, &sc.p, &e)) return fail(h, VF_ERR_CUDA, "mma weight prep: %s", e.c_str());
    float* dxs;                                                        // split-half copy of x (what a producer kernel would write)
    if (!sc.get(&dxs, nx)) return fail(h, VF_ERR_NOMEM, "conv scratch");
    View vxs = make_view(dxs, (long long)H * W * Cin, Cin, 0, Cin, (long long)B * H * W * Cin);
    launch_dense_to_view(dx, B, H * W, vxs, h->stream);
    MmaConvCall c;
    c.src = vxs; c.out = vy; c.sabias = nullptr; c.bias = db; c.H = H; c.W = W; c.passes = impl == VF_PREC_F16X3 ? 3 : 1;
    const int rc = mma_conv_launch(mw, c, B, h->stream);
    if (rc) return fail(h, VF_ERR_CUDA, "mma conv launch failed (code %d)", rc);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(y, dy, ny * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return VF_OK;
}

int vf_debug_conv_plan(int32_t k, int32_t kw, int32_t cin, int32_t cout, int32_t H, int32_t W, int32_t B, int32_t passes,
                       int32_t* out24) {
  if (!out24 || (k != 3 && k != 5) || (kw != k && kw != 1)) return VF_ERR_INVALID;
  int v[24];
  if (!mma_conv_describe(k, kw, cin, cout, H, W, B, passes, v)) return VF_ERR_UNSUPPORTED;
  for (int i = 0; i < 24; ++i) out24[i] = v[i];
  return VF_OK;
}

int vf_debug_conv_time(vf_engine* h, int32_t impl, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k,
                       int32_t reps, double* out_ms) {
  if (!h || !out_ms || (k != 3 && k != 5) || B < 1 || reps < 1) return fail(h, VF_ERR_INVALID, "bad conv arguments");
  const size_t nx = (size_t)B * H * W * Cin, nw = (size_t)k * k * Cin * Cout, ny = (size_t)B * H * W * Cout;
  // device buffers of this tool are freed on return (a sweep over many shapes must not accumulate them in the handle)
  float *dx = nullptr, *dy = nullptr, *dw = nullptr;
  auto release = [&]() { cudaFree(dx); cudaFree(dy); cudaFree(dw); };
  if (cudaMalloc(&dx, nx * sizeof(float)) != cudaSuccess || cudaMalloc(&dy, ny * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&dw, nw * sizeof(float)) != cudaSuccess) { release(); return fail(h, VF_ERR_NOMEM, "conv timing scratch"); }
  std::vector<float> wh(nw);
  uint32_t st = 12345u;
  const float wscale = 1.0f / sqrtf((float)(k * k * Cin));
  for (size_t i = 0; i < nw; ++i) { st = st * 1664525u + 1013904223u; wh[i] = ((float)(st >> 8) / 8388608.0f - 1.0f) * wscale; }
  cudaMemcpyAsync(dw, wh.data(), nw * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  launch_fill(dx, (long long)nx, 0.25f, h->stream);
  View vy = make_view(dy, (long long)H * W * Cout, Cout, 0, Cout);
  MmaConvWeights mw;
  std::vector<void*> wallocs;
  View vx = make_view(dx, (long long)H * W * Cin, Cin, 0, Cin);
  if (impl != VF_PREC_FP32_SIMT) {
    if (!mma_conv_supported(k, Cin, Cout, H, W)) { release(); return fail(h, VF_ERR_UNSUPPORTED, "shape not supported by the tcgen05 conv"); }
    std::string e;
    if (mma_conv_prepare_weights(wh.data(), k, k, k, Cin, Cout, &mw, &wallocs, &e)) { release(); return fail(h, VF_ERR_CUDA, "mma weight prep: %s", e.c_str()); }
    // split-half view over the same bytes: hi plane = first half of dx, lo plane B*H*W*Cin halfs later (values are irrelevant)
    vx = make_view(dx, (long long)H * W * Cin, Cin, 0, Cin, (long long)B * H * W * Cin);
  }
  auto one = [&]() -> int {
    if (impl == VF_PREC_FP32_SIMT) {
      ConvArgs a;
      a.src0 = vx; a.src1 = make_view(nullptr, 0, 0, 0, 0); a.w = dw; a.bias = nullptr; a.sabias = nullptr; a.out = vy;
      a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.k = k; a.act = ACT_NONE;
      launch_conv_simt(a, B, h->stream);
      return 0;
    }
    MmaConvCall c;
    c.src = vx; c.out = vy; c.sabias = nullptr; c.bias = nullptr; c.H = H; c.W = W; c.passes = impl == VF_PREC_F16X3 ? 3 : 1;
    return mma_conv_launch(mw, c, B, h->stream);
  };
  int rc = one();                                                      // warm-up (attribute setup, tensor-map cache)
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, h->stream);
  for (int i = 0; i < reps && !rc; ++i) rc = one();
  cudaEventRecord(e1, h->stream);
  cudaError_t ce = cudaStreamSynchronize(h->stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (void* q : wallocs) cudaFree(q);
  release();
  if (rc) return fail(h, VF_ERR_CUDA, "conv launch failed (code %d)", rc);
  if (ce != cudaSuccess) return fail(h, VF_ERR_CUDA, "conv timing: %s", cudaGetErrorString(ce));
  *out_ms = (double)ms / reps;
  return VF_OK;
}

int64_t vf_debug_fetch(vf_engine* h, const char* name, int32_t view, float* out, int64_t cap) {
  if (!h || !name || view < 0 || view >= h->ncam) return VF_ERR_INVALID;
  auto it = h->debug[view].find(name);
  if (it == h->debug[view].end()) return fail(h, VF_ERR_INVALID, "no debug tensor '%s'", name);
  const DebugEntry& d = it->second;
  const int M = h->last_M;
  const int64_t n = (int64_t)M * d.H * d.W * d.v.C;
  if (!out) return n;
  if (cap < n) return fail(h, VF_ERR_INVALID, "capacity %lld < %lld", (long long)cap, (long long)n);
  float* tmp;                                                          // gather through a dense float32 copy (either storage format)
  {
    void* q = nullptr;
    if (cudaMalloc(&q, (size_t)n * sizeof(float)) != cudaSuccess) return fail(h, VF_ERR_NOMEM, "debug fetch scratch");
    tmp = (float*)q;
  }
  launch_view_to_dense(d.v, M, d.H * d.W, tmp, h->stream);
  cudaError_t e = cudaMemcpyAsync(out, tmp, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(h, VF_ERR_CUDA, "debug fetch: %s", cudaGetErrorString(e));
  CU(cudaStreamSynchronize(h->stream));
  return n;
}

int vf_profile_enable(vf_engine* h, int32_t on) {
  if (!h) return VF_ERR_INVALID;
  CU(cudaStreamSynchronize(h->stream));
  for (auto& r : h->prof) { h->prof_pool.push_back(r.a); h->prof_pool.push_back(r.b); }
  h->prof.clear();
  h->prof_on = on != 0;
  return VF_OK;
}

int vf_profile_read(vf_engine* h, double* ms, double* flops, int64_t* launches, int32_t nclass) {
  if (!h || !ms || !flops || !launches || nclass < 2) return VF_ERR_INVALID;
  CU(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < nclass; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; }
  for (auto& r : h->prof) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.cls >= nclass) continue;
    ms[r.cls] += t; flops[r.cls] += r.flops; launches[r.cls] += 1;
  }
  return VF_OK;
}

int64_t vf_launch_count(const vf_engine* h) { (void)h; return g_launch_counter; }

}  // extern "C"

/* vfengine.h — C-ABI of libvfengine.so, the B200 (sm_100a) visual-MPC planning engine.
 *
 * The reference (SudeepDasari/visual_foresight) has no FFI; its hot path is Python calling a TF1
 * session.  Each entry point below names the reference interface it replaces (paths relative to
 * /root/reference/visual_mpc).  INTEGRATION.md shows the ctypes stub a reference maintainer adds.
 *
 * Conventions
 *   - C linkage, no exceptions cross the boundary.  Every call returns 0 on success, <0 on error;
 *     vf_last_error(h) returns a NUL-terminated description (valid until the next call on h).
 *   - Handles are opaque and NOT thread-safe (the reference is single-threaded per process,
 *     sim/run.py:140-156).  One CUDA stream per handle (vf_set_stream to adopt the caller's).
 *   - All buffers are caller-owned.  Pointers are HOST pointers unless the parameter name ends in
 *     _dev.  Calls that return host data synchronise the handle's stream; nothing else does.
 *   - Layouts are the reference's: frames (.., ncam, H, W, 3), distributions (.., ncam, H, W, ndesig),
 *     actions (M, T, adim), row-major, pixel coordinates (row, col).
 */
#ifndef VFENGINE_H_
#define VFENGINE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VF_API __attribute__((visibility("default")))
#else
#define VF_API
#endif

#define VF_ABI_VERSION 2
#define VF_MAX_LAYERS 8
#define VF_MAX_TASKS 16 /* ncam * ndesig */

typedef struct vf_engine vf_engine;

enum vf_status {
  VF_OK = 0,
  VF_ERR_INVALID = -1,   /* bad argument / shape */
  VF_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  VF_ERR_STATE = -3,     /* call order (weights not loaded, context not set, ...) */
  VF_ERR_NOMEM = -4,
  VF_ERR_UNSUPPORTED = -5
};

enum vf_dtype { VF_F32 = 0 };

/* arithmetic of the dense convolutions */
enum vf_precision {
  VF_PREC_FP32_SIMT = 0,   /* fp32 FFMA kernels (checker path for the tensor-core kernels) */
  VF_PREC_F16X3 = 1,       /* tcgen05 kind::f16, fp16 hi/lo split of both operands, 3 MMA passes, fp32 TMEM accumulate (fp32-grade) */
  VF_PREC_F16X1 = 2        /* tcgen05 kind::f16 single pass (fast; ~1e-3 relative per layer) */
};

enum vf_sampler {
  VF_SAMPLER_GAUSSIAN = 0,    /* samplers/gaussian_sampler.py: N(mu, Sigma), refit = mean + unbiased covariance of the elites */
  VF_SAMPLER_CORRELATED = 1   /* samplers/correlated_noise.py:17-66: AR(1)-smoothed noise around a softmax-weighted elite mean */
};

enum vf_cost_kind {
  VF_COST_PIXEL_DISTANCE = 0, /* pixel_cost_controller.py:135-187 expected distance of the designated-pixel distribution */
  VF_COST_GOAL_IMAGE = 1      /* goal_im_controller.py:87-93 MSE of the final predicted frame vs goal image (normalised to [0,1]) */
};

/* Replaces the model-construction half of video_prediction/setup_predictor.py:61-162 and the
 * model_hparams.json ingestion of vpred_model_interface.py:20-58: the layer table is data. */
typedef struct vf_config {
  int32_t abi_version;              /* = VF_ABI_VERSION */
  int32_t height, width;            /* conf['orig_size'] */
  int32_t ncam, ndesig;             /* conf['ncam'], conf['ndesig'] */
  int32_t adim, sdim, nz;           /* sdim = 0 -> use_state False */
  int32_t seq_len, context_frames;  /* model_hparams sequence_length / context_frames */
  int32_t ngf;
  int32_t n_enc;
  int32_t enc_channels[VF_MAX_LAYERS];
  int32_t enc_rnn[VF_MAX_LAYERS];
  int32_t n_dec;
  int32_t dec_channels[VF_MAX_LAYERS];
  int32_t dec_rnn[VF_MAX_LAYERS];
  int32_t num_transformed, cdna_ksize, lstm_ksize;
  float norm_eps, forget_bias;
  int32_t max_samples;              /* capacity in action samples held by THIS handle (per rank) */
  int32_t device;                   /* CUDA ordinal (policy ctor gpu_id, sim/simulator.py:19-21) */
  int32_t precision;                /* enum vf_precision */
  int32_t rnn_z;                    /* use_rnn_z: latent z -> dense LSTM(nz) -> tiled (weights zrnn.w [(z,h), 4nz] i,j,f,o; zrnn.b) */
  int32_t reserved[7];
} vf_config;

typedef struct vf_tensor {
  const char* name;                 /* "view{v}.<spec name>" or "<spec name>" for view 0 */
  int32_t dtype;                    /* enum vf_dtype */
  int32_t ndim;
  int64_t shape[6];
  const void* data;                 /* host pointer, C-contiguous */
} vf_tensor;

/* CEM hyper-parameters — cem_base_controller.py:42-64 + samplers/gaussian_sampler.py:51-71 */
typedef struct vf_cem_params {
  int32_t num_samples;              /* M on THIS handle (the local shard when world_size > 1) */
  int32_t global_samples;           /* M over all ranks (== num_samples when single GPU) */
  int32_t sample_offset;            /* global index of local sample 0 (setup_predictor.py:34-39 contiguous split) */
  int32_t iterations;
  int32_t num_elites;               /* K = max(int(selection_frac*M), minimum_selection), cem_base_controller.py:89-91 */
  int32_t nactions, repeat;         /* T = nactions*repeat */
  int32_t action_bound;             /* truncate_movement, controller_utils.py:6-44 */
  int32_t use_mean0;
  int32_t cost_kind;                /* enum vf_cost_kind */
  int32_t n_ctx_actions;            /* context actions prepended to the plan (C-1 for PixelCostController, 0 legacy) */
  int32_t pad0;
  /* float64 like the reference's sampler state (gaussian_sampler.py works in float64 throughout) */
  double initial_std[8];            /* per action dim: sqrt of construct_initial_sigma diag, controller_utils.py:47-84 */
  double clip_lo[8], clip_hi[8];    /* per action dim bounds (+-inf = unbounded) */
  double mean0[128];                /* initial mean (nactions*adim), zeros unless reuse_mean */
  double reduce_std_scale;          /* multiplies the VARIANCE of all but the last action block when t>=2 */
  double finalweight;               /* pixel_cost_controller.py:63,175-176 */
  double task_weights[VF_MAX_TASKS];/* per (cam, desig); 1/n reproduces np.mean, pixel_cost_controller.py:153 */
  uint64_t seed;                    /* Philox key */
  uint32_t plan_index;              /* Philox counter word: MPC step */
  /* Stochastic planning (nz > 0, BASELINE config c5): every action sequence is rolled k_futures times with independent
   * latents z ~ N(0, I) (consecutive copies, the order of np.repeat(actions, K, 0) in samplers/gaussian_sampler.py:139-141)
   * and scored by mean_k + lambda_variance * var_k (variants/ensemble_vidpred.py:56-58).  0 or 1 = one future.
   * num_samples * k_futures <= max_samples.  Latents are Philox draws keyed by the GLOBAL rollout index, so sharded plans
   * stay bit-identical. */
  int32_t k_futures;
  float lambda_variance;
  int32_t reserved[6];
  /* ---- ABI 2: sampler family and sampler options on the device path ---- */
  int32_t sampler;                  /* enum vf_sampler */
  int32_t n_append;                 /* append_action (cem_base_controller.py:94-96): the LAST n_append action dims of the model are
                                       constants; the sampler works on adim - n_append dims (initial_std, clip_*, mean0 index those) */
  uint32_t discrete_mask;           /* bit a: sampled dim a is floor()ed and clipped to [0, 4] before the movement clip
                                       (discrete_ind, controller_utils.py:107-117; gaussian_sampler.py:87-88) */
  int32_t pad1;
  double append_action[8];
  /* VF_SAMPLER_CORRELATED (repeat must be 1): noise_i = z_i * initial_std + mean_bias; a_i = beta0 * noise_i + beta1 * a_{i-1}
   * (i = 0 wraps to the un-smoothed LAST step, the reference's quirk); next mean = sum_k S_k elite_k / (sum S + 1e-4),
   * S_k = exp(kappa * (r_k - max r)), r = -score.  The external noise tensor then supplies nactions*(adim-n_append) normals
   * per sample in EVERY iteration. */
  double beta0, beta1, kappa;
  double mean_bias[8];
} vf_cem_params;

/* ---- lifecycle ---------------------------------------------------------------------------- */

/* setup_predictor.py:61-128 (graph + session + towers) */
VF_API int vf_create(const vf_config* cfg, vf_engine** out);
/* no counterpart (the reference never closes its session) */
VF_API int vf_destroy(vf_engine* h);
VF_API const char* vf_last_error(const vf_engine* h);
VF_API int vf_abi_version(void);
/* adopt a caller-owned cudaStream_t (torch.cuda.current_stream().cuda_stream); NULL -> own stream */
VF_API int vf_set_stream(vf_engine* h, void* cuda_stream);
VF_API int vf_synchronize(vf_engine* h);

/* setup_predictor.py:130-145 + checkpoint_matcher.py:4-38 (restore): name-addressed tensors */
VF_API int vf_load_weights(vf_engine* h, const vf_tensor* tensors, int32_t n);

/* ---- predictor ---------------------------------------------------------------------------- */

/* pred_util.py:4-13 get_context + the images/states/pix_distrib feeds of setup_predictor.py:171-198.
 * frames_u8 (C,ncam,H,W,3) uint8; states (C,sdim) or NULL; ctx_actions (n_ctx_actions,adim) or NULL;
 * pix_distrib (C,ncam,H,W,ndesig) f32 or NULL (one-hot is then built from desig_pix by vf_set_desig). */
VF_API int vf_set_context(vf_engine* h, const uint8_t* frames_u8, const float* states,
                   const float* ctx_actions, int32_t n_ctx_actions, const float* pix_distrib);
/* pixel_cost_controller.py:206-215 _switch_on_pix: desig (ncam,ndesig,2) (row,col), clipped, cast to int */
VF_API int vf_set_desig(vf_engine* h, const float* desig_pix);

/* predictor_func, setup_predictor.py:164-200 / VPredEvaluation.__call__ (pixel_cost_controller.py:83-84).
 * actions (M,T,adim) f32 host; zs (M,S-1,nz) or NULL.  Rolls S-1 cell steps; results stay on device.
 * Any out pointer may be NULL: out_frames (M,P,ncam,H,W,3), out_distrib (M,P,ncam,H,W,ndesig),
 * out_states (M,P,sdim). */
VF_API int vf_predict(vf_engine* h, const float* actions, int32_t M, int32_t T, const float* zs,
               float* out_frames, float* out_distrib, float* out_states);

/* _eval_pixel_cost / _expected_distance / _get_distancegrid, pixel_cost_controller.py:135-197, on the
 * device-resident result of the last vf_predict.  goal: (ncam,ndesig,2) pixels for PIXEL_DISTANCE,
 * (ncam,H,W,3) f32 in [0,1] for GOAL_IMAGE.  task_weights (ncam*ndesig) or NULL (= mean). */
VF_API int vf_score(vf_engine* h, int32_t cost_kind, const float* goal, const float* task_weights,
             float finalweight, double* out_scores);
/* same cost on caller-supplied distributions (M,P,ncam,H,W,ndesig) — lets a foreign predictor_class
 * (pixel_cost_controller.py:54) reuse the device cost kernel */
VF_API int vf_score_external(vf_engine* h, const float* distrib, int32_t M, int32_t P, const float* goal_pix,
                      const float* task_weights, float finalweight, double* out_scores);
/* fetch predictions of selected samples (verbose top-10 export pixel_cost_controller.py:88-131,
 * predictor_propagation :161-165). */
VF_API int vf_fetch(vf_engine* h, const int32_t* indices, int32_t n, float* out_frames, float* out_distrib);

/* ---- CEM ---------------------------------------------------------------------------------- */

/* perform_CEM, cem_base_controller.py:85-116, entirely on device.
 * noise: optional standard-normal draws replacing Philox (parity tests), shape
 * (iterations, global_samples, Dmax) f32 with Dmax = max(nactions*adim, num_elites): iteration 0
 * consumes the first D entries of a sample's row, iterations > 0 the first K; NULL in production.
 * Scores and sampled actions are float64 like the reference's (plan_stat['scores_itr*'],
 * np.random.multivariate_normal): out_best_actions (K,T,adim) f64, out_elite_idx (K) int32 global
 * indices in ascending cost, out_scores (iterations, global_samples) f64. */
VF_API int vf_cem_plan(vf_engine* h, const vf_cem_params* p, const float* goal, const float* noise,
                double* out_best_actions, int32_t* out_elite_idx, double* out_scores);

/* split form used when scores must be exchanged between ranks (one exchange per iteration):
 *   begin(it)  : sample (it>0: from the refit) -> rollout -> local scores into scores_dev[it][offset..]
 *   <caller all-gathers scores_dev rows on the same stream (NCCL) or the engine's peer path>
 *   end(it)    : top-K over global scores -> elites -> refit.  */
VF_API int vf_cem_begin(vf_engine* h, const vf_cem_params* p, const float* goal, const float* noise);
VF_API int vf_cem_iter_rollout(vf_engine* h, int32_t iteration);
VF_API int vf_cem_iter_select(vf_engine* h, int32_t iteration);
VF_API int vf_cem_finish(vf_engine* h, double* out_best_actions, int32_t* out_elite_idx, double* out_scores);
/* device pointer to the (iterations, global_samples) f64 score matrix, for the collective */
VF_API int vf_cem_scores_dev(vf_engine* h, void** scores_dev);
/* adopt a caller-owned DEVICE buffer of at least iterations*global_samples float64 as the score matrix of the following
 * vf_cem_begin calls (e.g. a torch tensor the collective library already knows); NULL -> back to the engine-owned one */
VF_API int vf_cem_bind_scores(vf_engine* h, void* scores_dev);
/* host access to one row segment of the score matrix: scores[iteration][offset : offset+n] (used by
 * the host-staged exchange when no device collective is available, and by the shard tests) */
VF_API int vf_cem_scores_read(vf_engine* h, int32_t iteration, int32_t offset, int32_t n, double* out);
VF_API int vf_cem_scores_write(vf_engine* h, int32_t iteration, int32_t offset, int32_t n, const double* in);
/* fetch the action tensor sampled in the last vf_cem_iter_rollout for the local samples: (M,T,adim) f64 */
VF_API int vf_cem_actions(vf_engine* h, double* out_actions);

/* ---- multi-GPU: score exchange over peer memory (NVLink / NVSwitch), owned by the engine ------------------------------
 * Replaces the reference's in-graph towers: contiguous action slice per GPU, outputs concatenated in rank order
 * (video_prediction/setup_predictor.py:34-44,117-123,155-162; ngpu comes from the policy ctor, sim/simulator.py:21).
 * One handle per GPU (one process per GPU, or several handles in one process).  Every handle owns an exchange WINDOW in its
 * device memory: the (iterations, global_samples) f64 score matrix plus one arrival counter per rank.
 *   vf_comm_export : allocate the window, write a 128-byte descriptor (process id, device, cudaIpcMemHandle, pointer)
 *   <the caller moves the descriptors between ranks: any channel — torch.distributed, MPI, a pipe, a file>
 *   vf_comm_connect: descs = world x 128 bytes in rank order; maps every peer window (cudaIpcOpenMemHandle across
 *                    processes, plain peer access inside one process)
 *   vf_cem_exchange: ONE kernel on the handle's stream: stores the local segment of score row `iteration` into every peer's
 *                    window (P2P stores), publishes this rank's arrival counter (release, system scope) and waits for the
 *                    peers' counters (acquire; bounded by VF_COMM_TIMEOUT_MS, default 30000 — a timeout is reported by
 *                    the next vf_cem_finish, never a hang).  No host synchronisation, no collective library.
 * Call order per plan on every rank: vf_cem_begin, then per iteration vf_cem_iter_rollout, vf_cem_exchange,
 * vf_cem_iter_select, then vf_cem_finish.  While connected, vf_cem_begin uses the window as the score matrix. */
#define VF_PEER_DESC_BYTES 128
VF_API int vf_comm_export(vf_engine* h, int32_t max_iterations, int32_t max_global_samples, void* out_desc);
VF_API int vf_comm_connect(vf_engine* h, int32_t rank, int32_t world, const void* descs);
VF_API int vf_cem_exchange(vf_engine* h, int32_t iteration);
/* unmap the peers and release the window (also done by vf_destroy) */
VF_API int vf_comm_close(vf_engine* h);

/* unit-parity hooks: cem_base_controller.py:104 (argsort()[:K], stable, ties -> lower index) */
VF_API int vf_topk(vf_engine* h, const double* scores, int32_t n, int32_t k, int32_t* out_idx);
/* samplers/gaussian_sampler.py:96-107 _fit_gaussians on elites (K,T,adim) f64: mean (D), cov (D,D),
 * factor (D,K) with factor @ factor.T == cov  (D = nactions*adim) */
VF_API int vf_refit(vf_engine* h, const double* elites, int32_t K, int32_t nactions, int32_t repeat,
             int32_t adim, double* out_mean, double* out_cov, double* out_factor);

/* ---- debug / unit tests --------------------------------------------------------------------- */
/* raw NHWC convolution through the engine's conv kernels (SAME, stride 1): x (B,H,W,Cin),
 * w (k,k,Cin,Cout) HWIO, bias (Cout) or NULL -> y (B,H,W,Cout).  impl: enum vf_precision */
VF_API int vf_debug_conv2d(vf_engine* h, int32_t impl, const float* x, const float* w, const float* bias,
                    int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k, float* y);
/* Host-only (no device, no handle): the tiling the tcgen05 convolution picks for a k x kw layer (kw = k, or 1 for a
 * dx-folded input) with B samples, as 24 integers: {swap, rg, G, npass, v_cnt, units, ncols, max MMA N, TMEM columns per
 * accumulator set, accumulator sets, activation buffers, weight stages, stage bytes, plane bytes, dynamic shared memory
 * bytes, work items, TMA box rows, padded row pitch, image pitch in pixel rows, box bytes, last pixel row an item's MMAs
 * read, pixel rows a plane holds, channel chunks, Cout tiles}.  VF_ERR_UNSUPPORTED when the shape has no plan.  Used by
 * the CPU test suite to check the resource invariants of every layer shape. */
VF_API int vf_debug_conv_plan(int32_t k, int32_t kw, int32_t cin, int32_t cout, int32_t H, int32_t W, int32_t B,
                       int32_t passes, int32_t* out24);
/* tuning aid (profiles/conv_microbench.py): average milliseconds of `reps` back-to-back launches of ONE convolution of the
 * given shape on data resident in HBM, CUDA events on the handle's stream.  Not part of the reference surface. */
VF_API int vf_debug_conv_time(vf_engine* h, int32_t impl, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                       int32_t k, int32_t reps, double* out_ms);
/* copy a named internal activation of the LAST cell step to host (tests): returns element count */
VF_API int64_t vf_debug_fetch(vf_engine* h, const char* name, int32_t view, float* out, int64_t capacity);
/* per-kernel-class timing with CUDA events on the handle's stream (bench roofline): enable, run, read.
 * classes: 0 = conv-LSTM gate convolutions, 1 = all other convolutions, 2 (optional, nclass >= 3) = convolutions of the
 * shared-prefix cell steps that run once on one sample instead of on all M.  ms / flops / launches are
 * accumulated since the last enable; flops are the EXECUTED 2*MAC of the launches (sa channels folded). */
VF_API int vf_profile_enable(vf_engine* h, int32_t on);
VF_API int vf_profile_read(vf_engine* h, double* out_ms, double* out_flops, int64_t* out_launches, int32_t nclass);
/* number of kernels this handle launched since creation (bench gpu_launches) */
VF_API int64_t vf_launch_count(const vf_engine* h);

#ifdef __cplusplus
}
#endif
#endif /* VFENGINE_H_ */

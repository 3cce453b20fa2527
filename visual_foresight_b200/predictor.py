"""Engine-backed predictor objects for the reference's two predictor plugin hooks.

(i)  ``predictor_class`` (reference ``pixel_cost_controller.py:29-36,54,83-84,175``):
     ``obj = cls(model_path, {'designated_pixel_count', 'run_batch_size'}, n_gpus=, first_gpu=)``,
     ``obj.restore()``, ``obj.n_context``, ``obj.sequence_length``,
     ``obj(context_dict, {'actions': (M,T,adim)}) -> {'predicted_frames', 'predicted_pixel_distributions'}``
     -> ``B200VPredEvaluation``.
(ii) ``netconf['setup_predictor'](ag_params, netconf, gpu_id, ngpu, logger) -> predictor_func``
     (reference ``video_prediction/setup_predictor.py:61-202``, call site ``goal_im_controller.py:25-27``)
     -> ``setup_predictor``.

Weights come from an ``.npz`` (``<view>.<name>`` keys, see ``weights_io``) or, with an empty
``model_path``, from the seeded synthetic initialisation of ``spec.init_weights`` (there are no
published checkpoints to load: SURVEY.md 8c)."""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Any, Dict, Optional

import numpy as np

from . import spec as specmod
from .engine import (COST_GOAL_IMAGE, COST_PIXEL_DISTANCE, SAMPLER_CORRELATED, SAMPLER_GAUSSIAN, Engine, VfCemParams)
from .spec import PredictorSpec


# ---------------------------------------------------------------------------------------------------
def save_weights(path: str, spec: PredictorSpec, weights_per_view) -> None:
    flat = {}
    for v, wd in enumerate(weights_per_view):
        for k, a in wd.items():
            flat["view%d.%s" % (v, k)] = np.asarray(a, np.float32)
    flat["__spec__"] = np.frombuffer(json.dumps(dataclasses.asdict(spec)).encode(), dtype=np.uint8)
    np.savez(path, **flat)


def load_weights(path: str):
    z = np.load(path)
    d = json.loads(bytes(z["__spec__"]).decode())
    d["encoder"] = tuple((int(a), bool(b)) for a, b in d["encoder"])
    d["decoder"] = tuple((int(a), bool(b)) for a, b in d["decoder"])
    spec = PredictorSpec(**d)
    views = [dict() for _ in range(spec.ncam)]
    for k in z.files:
        if k.startswith("view"):
            v, name = k.split(".", 1)
            views[int(v[4:])][name] = z[k]
    return spec, views


def cem_params(spec: PredictorSpec, *, num_samples, iterations, num_elites, nactions, repeat, std, clip=None,
               mean0=None, reduce_std_scale=1.0, cost_kind=COST_PIXEL_DISTANCE, finalweight=10.0, task_weights=None,
               n_ctx_actions=0, seed=0, plan_index=0, global_samples=None, sample_offset=0, k_futures=1,
               lambda_variance=0.0, sampler="gaussian", append_action=None, discrete_ind=None, beta0=0.5, beta1=0.5, kappa=1.0,
               mean_bias=None) -> VfCemParams:
    """std / clip / mean0 are indexed by the SAMPLED action dims (adim - len(append_action))."""
    p = VfCemParams()
    p.num_samples = int(num_samples)
    p.global_samples = int(global_samples if global_samples is not None else num_samples)
    p.sample_offset = int(sample_offset)
    p.iterations, p.num_elites, p.nactions, p.repeat = int(iterations), int(num_elites), int(nactions), int(repeat)
    app = [] if append_action is None else [float(v) for v in np.asarray(append_action, np.float64).reshape(-1)]
    sdims = spec.adim - len(app)
    assert sdims >= 1, "append_action leaves no sampled action dimension"
    p.n_append = len(app)
    for i, v in enumerate(app):
        p.append_action[i] = v
    std = np.asarray(std, dtype=np.float64).reshape(-1)
    assert std.shape[0] >= sdims
    for i in range(8):
        p.initial_std[i] = float(std[i]) if i < sdims else 0.0
        p.clip_lo[i], p.clip_hi[i] = -np.inf, np.inf
    p.action_bound = int(clip is not None)
    if clip is not None:
        lo, hi = clip
        for i in range(sdims):
            p.clip_lo[i], p.clip_hi[i] = float(lo[i]), float(hi[i])
    p.discrete_mask = 0
    for ind in (discrete_ind or []):
        assert 0 <= int(ind) < sdims
        p.discrete_mask |= 1 << int(ind)
    p.sampler = {"gaussian": SAMPLER_GAUSSIAN, "correlated": SAMPLER_CORRELATED}[sampler] if isinstance(sampler, str) else int(sampler)
    p.beta0, p.beta1, p.kappa = float(beta0), float(beta1), float(kappa)
    mb = np.zeros(8) if mean_bias is None else np.asarray(mean_bias, np.float64).reshape(-1)
    for i in range(min(8, mb.shape[0])):
        p.mean_bias[i] = float(mb[i])
    p.use_mean0 = int(mean0 is not None)
    if mean0 is not None:
        m = np.asarray(mean0, dtype=np.float64).reshape(-1)
        assert m.shape[0] == nactions * sdims
        for i, v in enumerate(m):
            p.mean0[i] = float(v)
    p.reduce_std_scale = float(reduce_std_scale)
    p.cost_kind, p.finalweight = int(cost_kind), float(finalweight)
    ntask = spec.ncam * spec.ndesig
    tw = np.full(ntask, 1.0 / ntask) if task_weights is None else np.asarray(task_weights, dtype=np.float64).reshape(-1)
    for i in range(ntask):
        p.task_weights[i] = float(tw[i])
    p.n_ctx_actions = int(n_ctx_actions)
    p.seed, p.plan_index = int(seed), int(plan_index)
    p.k_futures, p.lambda_variance = int(k_futures), float(lambda_variance)
    return p


# ---------------------------------------------------------------------------------------------------
class EngineBackend:
    """Rollout evaluator on one GPU: owns the vf_engine handle for one (spec, max_samples)."""

    def __init__(self, spec: PredictorSpec, weights_per_view, max_samples: int, device: int = 0,
                 precision="f16x3", state_append=None):
        self.spec = spec
        self.engine = Engine(spec, max_samples, device=device, precision=precision)
        self.engine.load_weights(weights_per_view)
        self.state_append = None if state_append is None else np.asarray(state_append, np.float32).reshape(-1)
        self.n_context, self.sequence_length, self.n_cam = spec.context_frames, spec.seq_len, spec.ncam
        self._n_ctx_actions = 0

    # context dict -> engine feeds (what robonet's VPredEvaluation does before its sess.run; the legacy
    # equivalent is pred_util.get_context, reference pred_util.py:4-13)
    def set_context(self, context: Dict[str, Any], legacy_actions: bool = False) -> None:
        sp = self.spec
        C = sp.context_frames
        frames = np.asarray(context["context_frames"])[-C:]
        if frames.ndim == 4:
            frames = frames[:, None]
        frames = frames[:, :sp.ncam]
        states = None
        if sp.sdim > 0:
            st = np.asarray(context["context_states"], dtype=np.float32)[-C:]
            if self.state_append is not None:
                st = np.concatenate([st, np.tile(self.state_append[None], (C, 1))], axis=-1)
            if st.shape[-1] < sp.sdim:
                raise ValueError("context_states has %d dims, the model needs %d" % (st.shape[-1], sp.sdim))
            states = st[:, :sp.sdim]
        ca = None
        if not legacy_actions and C > 1:
            hist = np.asarray(context.get("context_actions", np.zeros((0, sp.adim))), dtype=np.float32)
            hist = hist.reshape(-1, sp.adim) if hist.size else np.zeros((0, sp.adim), np.float32)
            if hist.shape[0] < C - 1:       # fewer executed actions than context frames: pad with zero actions
                hist = np.concatenate([np.zeros((C - 1 - hist.shape[0], sp.adim), np.float32), hist], 0)
            ca = hist[-(C - 1):]
        self._n_ctx_actions = 0 if ca is None else ca.shape[0]
        pd = context.get("context_pixel_distributions")
        if pd is not None:
            pd = np.asarray(pd, np.float32)[-C:, :sp.ncam]
        self.engine.set_context(frames, states, ca, pd)

    def predict(self, context, actions, legacy_actions=False):
        self.set_context(context, legacy_actions)
        return self.engine.predict(np.asarray(actions, np.float32))

    def evaluate(self, context, actions, goal_pix, finalweight, task_weights):
        self.set_context(context)
        a = np.asarray(actions, np.float32)
        self.engine.predict(a, fetch=False)
        return self.engine.score(np.asarray(goal_pix, np.float32), COST_PIXEL_DISTANCE, task_weights, finalweight,
                                 M=a.shape[0])

    def evaluate_goal_image(self, context, actions, goal_image):
        self.set_context(context)
        a = np.asarray(actions, np.float32)
        self.engine.predict(a, fetch=False)
        return self.engine.score(np.asarray(goal_image, np.float32), COST_GOAL_IMAGE, None, 1.0, M=a.shape[0])

    def fetch_distrib(self, index: int):
        return self.engine.fetch([index], frames=False)[1][0]

    def fetch_top(self, indices):
        return self.engine.fetch(indices)

    def score_external(self, distrib, goal_pix, finalweight, task_weights):
        return self.engine.score_external(distrib, np.asarray(goal_pix, np.float32), task_weights, finalweight)

    def plan(self, context, *, num_samples, iterations, num_elites, nactions, repeat, std, clip, mean0,
             reduce_std_scale, goal_pix, finalweight, task_weights, seed, plan_index, noise=None,
             cost_kind=COST_PIXEL_DISTANCE, k_futures=1, lambda_variance=0.0, **sampler_kw):
        self.set_context(context)
        p = cem_params(self.spec, num_samples=num_samples, iterations=iterations, num_elites=num_elites,
                       nactions=nactions, repeat=repeat, std=std, clip=clip, mean0=mean0,
                       reduce_std_scale=reduce_std_scale, cost_kind=cost_kind, finalweight=finalweight,
                       task_weights=task_weights, n_ctx_actions=self._n_ctx_actions, seed=seed, plan_index=plan_index,
                       k_futures=k_futures, lambda_variance=lambda_variance, **sampler_kw)
        best, eidx, scores = self.engine.cem_plan(p, np.asarray(goal_pix, np.float32), noise)
        return {"best_actions": best, "elite_idx": eidx, "scores": scores}


def make_cost_backend(distrib_shape, device=0):
    """Smallest engine able to run the device cost kernel on foreign predictions (M,P,ncam,H,W,nd)."""
    M, P, ncam, H, W, nd = distrib_shape
    sp = specmod.spec_64(height=H, width=W, ncam=ncam, ndesig=nd, seq_len=P + 2, context_frames=2)

    class _CostOnly:
        def __init__(self):
            self.engine = Engine(sp, M, device=device)

        def score_external(self, distrib, goal_pix, finalweight, task_weights):
            return self.engine.score_external(distrib, np.asarray(goal_pix, np.float32), task_weights, finalweight)

    return _CostOnly()


# ---------------------------------------------------------------------------------------------------
class B200VPredEvaluation:
    """``predictor_class``-compatible predictor on the B200 engine (also usable inside the REFERENCE's
    own PixelCostController via its ``predictor_class`` hparam)."""

    def __init__(self, model_path, hparams, n_gpus=1, first_gpu=0):
        self._model_path = model_path
        self._hp = dict(hparams)
        self._n_gpus, self._first_gpu = n_gpus, first_gpu
        self.backend: Optional[EngineBackend] = None
        pol = self._hp.get("policy_hparams", {})
        ag = self._hp.get("ag_params", {})
        if model_path:
            self.spec, self._weights = load_weights(model_path)
        else:
            over = dict(pol.get("model_spec") or self._hp.get("model_spec") or {})
            if ag:
                over.setdefault("height", ag["image_height"])
                over.setdefault("width", ag["image_width"])
                over.setdefault("adim", ag["adim"])
                sdim = ag["sdim"] + (len(pol["state_append"]) if pol.get("state_append") else 0)
                over.setdefault("sdim", sdim)
            over.setdefault("ndesig", self._hp.get("designated_pixel_count", 1))
            family = over.pop("family", "64")
            self.spec = (specmod.spec_128 if str(family) == "128" else specmod.spec_64)(**over)
            seed = int(pol.get("model_seed", self._hp.get("model_seed", 0)) or 0)
            self._weights = [specmod.init_weights(self.spec, seed, v) for v in range(self.spec.ncam)]
        self.n_context = self.spec.context_frames
        self.sequence_length = self.spec.seq_len
        self.n_cam = self.spec.ncam
        self._precision = pol.get("precision", self._hp.get("precision", "f16x3"))
        # rollout capacity: every action sequence is rolled num_futures times under stochastic planning; with ngpu > 1 this
        # rank holds a contiguous 1/ngpu slice of the samples (setup_predictor.py:34-39; M % ngpu == 0 :70)
        M = int(pol.get("num_samples", self._hp.get("run_batch_size", 200)))
        if n_gpus > 1 and M % n_gpus:
            raise ValueError("num_samples (%d) must be divisible by ngpu (%d)" % (M, n_gpus))
        self._max_samples = (M // max(n_gpus, 1)) * max(int(pol.get("num_futures", 1) or 1), 1)
        self._state_append = pol.get("state_append")
        self._collective = pol.get("collective", self._hp.get("collective", "peer")) or "peer"
        self._shard_devices = pol.get("shard_devices", self._hp.get("shard_devices"))

    def restore(self):
        device, rank = self._first_gpu, 0
        if self._n_gpus > 1:
            # The reference builds ngpu towers inside ONE TF graph (setup_predictor.py:117-123).  Here: one process per GPU
            # (torchrun), every process constructs the same policy with the same (gpu_id, ngpu); rank r drives GPU
            # gpu_id + local_rank and holds samples [r*M/ngpu, (r+1)*M/ngpu).
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() == self._n_gpus):
                raise RuntimeError("ngpu=%d needs one process per GPU under torch.distributed (torchrun) with world size %d; "
                                   "in-process towers (setup_predictor.py:117-123) are not reproduced" % (self._n_gpus, self._n_gpus))
            rank = dist.get_rank()
            local_rank = int(os.environ.get("LOCAL_RANK", rank % self._n_gpus))
            device = self._first_gpu + local_rank if self._shard_devices is None else int(self._shard_devices[rank])
        backend = EngineBackend(self.spec, self._weights, self._max_samples, device=device,
                                precision=self._precision, state_append=self._state_append)
        if self._n_gpus > 1:
            from .distributed import ShardedBackend
            backend = ShardedBackend(backend, rank, self._n_gpus, collective=self._collective)
        self.backend = backend
        self._weights = None

    def __call__(self, context, inputs):
        actions = np.asarray(inputs["actions"])
        cap = self._max_samples
        if actions.shape[0] <= cap:
            frames, distrib, _ = self.backend.predict(context, actions)
        else:
            # more rollouts than the handle holds: run_batch_size-sized calls like the reference's rollout_predictions
            # (pred_util.py:21-48); the engine takes a ragged last chunk, so nothing is padded
            parts = [self.backend.predict(context, actions[i:i + cap]) for i in range(0, actions.shape[0], cap)]
            frames = np.concatenate([p[0] for p in parts], axis=0)
            distrib = np.concatenate([p[1] for p in parts], axis=0)
        return {"predicted_frames": frames, "predicted_pixel_distributions": distrib}


def setup_predictor(hyperparams, conf, gpu_id=0, ngpu=1, logger=None):
    """Legacy hook: returns ``predictor_func(input_images, input_one_hot_images, input_state, input_actions)
    -> (gen_images, gen_distrib, gen_states)`` with the reference's shapes (setup_predictor.py:98-114,164-200).
    ``conf`` keys honoured: orig_size, ncam, adim, sdim, ndesig, sequence_length, context_frames, batch_size,
    pretrained_model ('' -> synthetic weights), model_seed, precision."""
    H, W = conf["orig_size"]
    if conf.get("pretrained_model"):
        spec, weights = load_weights(conf["pretrained_model"])
    else:
        spec = specmod.spec_64(height=H, width=W, ncam=conf.get("ncam", 1), ndesig=conf.get("ndesig", 1),
                               adim=conf["adim"], sdim=conf["sdim"], seq_len=conf["sequence_length"],
                               context_frames=conf["context_frames"])
        weights = [specmod.init_weights(spec, int(conf.get("model_seed", 0)), v) for v in range(spec.ncam)]
    backend = EngineBackend(spec, weights, int(conf["batch_size"]), device=gpu_id, precision=conf.get("precision", "f16x3"))

    def predictor_func(input_images=None, input_one_hot_images=None, input_state=None, input_actions=None):
        frames = np.asarray(input_images)[0]                        # (C,ncam,H,W,3) float in [0,1] (get_context)
        frames_u8 = np.clip(np.rint(frames * 255.0), 0, 255).astype(np.uint8)
        ctx = {"context_frames": frames_u8,
               "context_states": None if input_state is None else np.asarray(input_state)[0],
               "context_pixel_distributions": None if input_one_hot_images is None else np.asarray(input_one_hot_images)[0]}
        if ctx["context_pixel_distributions"] is None:
            ctx["context_pixel_distributions"] = np.full((spec.context_frames, spec.ncam, H, W, spec.ndesig),
                                                         1.0 / (H * W), np.float32)
        gi, gd, gs = backend.predict(ctx, input_actions, legacy_actions=True)
        return gi, (None if input_one_hot_images is None else gd), gs

    predictor_func.backend = backend
    return predictor_func

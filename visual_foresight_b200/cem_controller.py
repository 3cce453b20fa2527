"""CEM visual-MPC controllers behind the reference's ``Policy.act()`` surface.

Drop-in for ``visual_mpc/policy/cem_controllers/cem_base_controller.py`` (hparams :42-64,
``perform_CEM`` :85-116, ``act`` :127-169) and ``pixel_cost_controller.py`` (ctor :20-50, hparams
:52-69, ``evaluate_rollouts`` :76-133, ``act`` :217-233).  Same constructor
``(ag_params, policyparams, gpu_id, ngpu)``, same ``reset()`` / ``act(...)`` keywords (resolved by
name through ``get_policy_args``), same return dict (``'actions'``, ``'plan_stat'`` with
``scores_itr{i}``).

Two execution paths, both on the GPU engine:
  * device path (default sampler, no host-only options): one ``vf_cem_*`` sequence per plan —
    sampling, 14 cell steps, cost, top-K and refit never leave the device;
  * plugin path (user ``sampler`` class, foreign ``predictor_class``, smooth_cov/blockdiag/
    rejection sampling): the host sampler draws actions, the engine predicts + scores them.
There is no CPU evaluation path in this package; tests inject an oracle backend explicitly.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from .hparams import HParams
from .policy import Policy
from .samplers import CorrelatedNoiseSampler, GaussianCEMSampler, action_bounds, per_dim_variance


class _Log(object):
    """Logger with the reference's call shape (utils/logger.py:3-25); silent unless asked."""

    def __init__(self, path: Optional[str] = None, echo: bool = False):
        self._path, self._echo = path, echo

    def log(self, *items):
        if self._echo:
            print(items)
        elif self._path:
            with open(self._path, "a") as f:
                f.write("".join(str(i) for i in items) + "\n")


class CEMBaseController(Policy):
    """Cross-entropy-method optimiser shell.  Subclasses implement ``evaluate_rollouts``."""

    def __init__(self, ag_params, policyparams):
        self._hp = self._default_hparams()
        self._override_defaults(policyparams)
        self.agentparams = ag_params
        if self._hp.logging_dir:
            import os
            self._logger = _Log(os.path.join(self._hp.logging_dir, "cem{}log.txt".format(ag_params.get("gpu_id", 0))))
        else:
            self._logger = _Log(echo=bool(self._hp.get("log_to_stdout", False)))
        self._adim, self._sdim = ag_params["adim"], ag_params["sdim"]
        self._n_iter = self._hp.iterations
        self._t = self._t_since_replan = None
        self._sampler = None
        self._best_indices = self._best_actions = None
        self._state = None
        self.plan_stat: Dict[str, Any] = {}
        assert self._hp.minimum_selection > 0, "must take at least 1 sample for refitting"

    # -- hyper-parameters -----------------------------------------------------------------------------
    def _default_hparams(self) -> HParams:
        hp = super()._default_hparams()
        for k, v in dict(append_action=None, verbose=True, verbose_every_iter=False, logging_dir="",
                         hard_coded_start_action=None, context_action_weight=[0.5, 0.5, 0.05, 1],
                         zeros_for_start_frames=True, replan_interval=0, sampler=GaussianCEMSampler, T=15,
                         iterations=3, num_samples=200, selection_frac=0., start_planning=0,
                         minimum_selection=10).items():
            hp.add_hparam(k, v)
        return hp

    def _override_defaults(self, policyparams):
        sampler_cls = policyparams.get("sampler", GaussianCEMSampler)
        for k, v in sampler_cls.get_default_hparams().items():     # sampler defaults join the controller's
            if k in self._hp:
                self._hp.set_hparam(k, v)
            else:
                self._hp.add_hparam(k, v)
        super()._override_defaults(policyparams)
        self._hp.sampler = sampler_cls

    # -- per-trajectory / per-step ----------------------------------------------------------------------
    def reset(self):
        self._best_indices = self._best_actions = None
        self._t_since_replan = None
        self._sampler = self._hp.sampler(self._hp, self._adim, self._sdim)
        self.plan_stat = {}

    def num_elites(self) -> int:
        k = self._hp.minimum_selection
        if self._hp.selection_frac:
            k = max(int(self._hp.selection_frac * self._hp.num_samples), k)
        return k

    def evaluate_rollouts(self, actions, cem_itr):
        raise NotImplementedError

    def _verbose_condition(self, cem_itr):
        return bool(self._hp.verbose and (self._hp.verbose_every_iter or cem_itr == self._n_iter - 1))

    def perform_CEM(self, state):
        """Host-driven CEM: sampler -> evaluate_rollouts -> stable top-K -> refit (reference :85-116)."""
        hp = self._hp
        K = self.num_elites()
        actions = self._sampler.sample_initial_actions(self._t, hp.num_samples, state[-1])
        for itr in range(self._n_iter):
            if hp.append_action:
                tail = np.tile(np.asarray(hp.append_action)[None, None], [actions.shape[0], actions.shape[1], 1])
                actions = np.concatenate((actions, tail), axis=-1)
            scores = np.asarray(self.evaluate_rollouts(actions, itr))
            assert scores.shape == (actions.shape[0],), "score shape should be (n_actions,)"
            self._best_indices = scores.argsort()[:K]
            self._best_actions = actions[self._best_indices]
            self.plan_stat["scores_itr{}".format(itr)] = scores
            if itr < self._n_iter - 1:
                elites = self._best_actions.copy()
                if hp.append_action:
                    elites = elites[:, :, :-len(hp.append_action)]
                actions = self._sampler.sample_next_actions(hp.num_samples, elites, scores[self._best_indices].copy())
        self._t_since_replan = 0

    def act(self, t=None, i_tr=None, state=None):
        hp = self._hp
        self._state, self.i_tr, self._t = state, i_tr, t
        if t < hp.start_planning:
            if hp.zeros_for_start_frames:
                assert hp.hard_coded_start_action is None
                action = np.zeros(self.agentparams["adim"])
            elif hp.hard_coded_start_action:
                action = np.array(hp.hard_coded_start_action)
            else:
                warm = hp.sampler(hp, self._adim, self._sdim)
                action = warm.sample_initial_actions(t, 1, state[-1])[0, 0] * hp.context_action_weight
                if hp.append_action:
                    action = np.concatenate((action, hp.append_action), axis=0)
        else:
            due = (not hp.replan_interval) or self._t_since_replan is None or \
                self._t_since_replan + 1 >= hp.replan_interval
            if due:
                self.perform_CEM(state)
            else:
                self._t_since_replan += 1
            action = self._best_actions[0, self._t_since_replan]
        assert action.shape == (self.agentparams["adim"],), "action shape does not match adim!"
        self._logger.log("time {}, action - {}".format(t, action))
        if self._best_actions is not None:
            tail = self._best_actions[:, min(self._t_since_replan + 1, hp.T - 1):]
            self._sampler.log_best_action(action, tail)
        else:
            self._sampler.log_best_action(action, None)
        return {"actions": action, "plan_stat": self.plan_stat}


# =======================================================================================================
class PixelCostController(CEMBaseController):
    """Designated-pixel expected-distance controller on the B200 engine."""

    def __init__(self, ag_params, policyparams, gpu_id=0, ngpu=1):
        CEMBaseController.__init__(self, ag_params, policyparams)
        hp = self._hp
        from .predictor import B200VPredEvaluation
        cls = hp.predictor_class if hp.predictor_class is not None else B200VPredEvaluation
        phparams = {"designated_pixel_count": hp.designated_pixel_count,
                    "run_batch_size": min(hp.vpred_batch_size, hp.num_samples)}
        if cls is B200VPredEvaluation or (isinstance(cls, type) and issubclass(cls, B200VPredEvaluation)):
            phparams.update(ag_params=ag_params, policy_hparams=hp.values())
        self.predictor = cls(hp.model_path, phparams, n_gpus=ngpu, first_gpu=gpu_id)
        self.predictor.restore()
        self._net_context = self.predictor.n_context
        if hp.start_planning < self._net_context - 1:
            hp.start_planning = self._net_context - 1
        self._n_desig = hp.designated_pixel_count
        self._img_height, self._img_width = ag_params["image_height"], ag_params["image_width"]
        # the reference hard-codes 1 (pixel_cost_controller.py:43); the engine supports ncam views
        self._n_cam = int(getattr(self.predictor, "n_cam", 1)) if hp.get("use_predictor_ncam", False) else 1
        self._desig_pix = self._goal_pix = self._images = None
        self._chosen_distrib = None
        self._backend = getattr(self.predictor, "backend", None)      # engine-backed rollout evaluator
        self._cost_backend = None                                     # lazily-built engine for foreign predictors
        self._verbose_worker = None
        self.last_verbose = None

    def _export_verbose(self, scores):
        """verbose=True in the reference renders the 10 best rollouts to HTML/GIF through a saver process
        (pixel_cost_controller.py:88-131, out of scope).  Here the same 10 rollouts are fetched from the
        device and kept as arrays in ``self.last_verbose`` for whoever wants to render them."""
        if self._backend is None:
            return
        top = np.argsort(scores, kind="stable")[:10].astype(np.int32)
        frames, distrib = self._backend.fetch_top(top)
        self.last_verbose = {"indices": top, "scores": np.asarray(scores)[top], "frames": frames, "distrib": distrib}

    def _default_hparams(self):
        hp = super()._default_hparams()
        for k, v in dict(predictor_class=None, model_path="", vpred_batch_size=200, designated_pixel_count=1,
                         verbose_img_height=128, predictor_propagation=False, only_take_first_view=False,
                         state_append=None, finalweight=10., use_predictor_ncam=False, device_cem=True,
                         cem_seed=0, task_weights=None, log_to_stdout=False, model_spec=None, model_seed=0,
                         precision="f16x3",
                         # ngpu > 1 (one process per GPU): how the per-iteration scores cross ranks ("peer": the engine's
                         # own NVLink peer-memory kernel; "nccl": torch.distributed all-gather; "host": gloo) and, optionally,
                         # the CUDA ordinal of every rank (default gpu_id + local rank)
                         collective="peer", shard_devices=None,
                         # stochastic planning (samplers/gaussian_sampler.py:139-141 repeats every action sequence K times;
                         # variants/ensemble_vidpred.py:56-58 scores mean + lambda * var): futures per action sequence
                         num_futures=1, lambda_variance=0.0).items():
            hp.add_hparam(k, v)
        return hp

    def reset(self):
        super().reset()
        self._chosen_distrib = None
        self._plan_counter = 0

    # -- inputs -----------------------------------------------------------------------------------------
    def _switch_on_pix(self, desig):
        onehot = np.zeros((self._net_context, self._n_cam, self._img_height, self._img_width, self._n_desig), np.float32)
        hi = np.array([self._img_height, self._img_width]).reshape(1, 2) - 1
        pix = np.clip(desig, np.zeros((1, 2)), hi).astype(int)
        for c in range(self._n_cam):
            for p in range(self._n_desig):
                onehot[:, c, pix[c, p, 0], pix[c, p, 1], p] = 1.
        return onehot

    def _make_input_distrib(self, itr):
        if self._hp.predictor_propagation and self._chosen_distrib is not None:
            return self._chosen_distrib[-self._net_context:]
        return self._switch_on_pix(self._desig_pix)

    def _task_weights(self):
        n = self._n_cam * self._n_desig
        if self._hp.task_weights is not None:
            w = np.asarray(self._hp.task_weights, dtype=np.float64).reshape(-1)
            assert w.shape[0] == n
            return w
        if self._hp.only_take_first_view:
            w = np.zeros(n)
            w[0] = 1.0                        # scores_per_task[:, 0] (reference :150-151)
            return w
        return np.full(n, 1.0 / n)

    # -- plugin path --------------------------------------------------------------------------------------
    def evaluate_rollouts(self, actions, cem_itr):
        hp = self._hp
        context = {"context_frames": self._images, "context_actions": self._sampler.chosen_actions,
                   "context_pixel_distributions": self._make_input_distrib(cem_itr), "context_states": self._state}
        if self._backend is not None:
            scores = self._backend.evaluate(context, actions, self._goal_pix, hp.finalweight, self._task_weights())
            if hp.predictor_propagation and cem_itr == hp.iterations - 1:
                best = int(np.argsort(scores, kind="stable")[0])
                self._chosen_distrib = self._backend.fetch_distrib(best)
            if self._verbose_condition(cem_itr):
                self._export_verbose(scores)
            return scores
        # foreign predictor_class: it returns host arrays; the cost still runs on the device
        pred = self.predictor(context, {"actions": actions})
        gen_distrib = pred["predicted_pixel_distributions"]
        scores = self._external_cost(gen_distrib)
        if hp.predictor_propagation and cem_itr == hp.iterations - 1:
            self._chosen_distrib = gen_distrib[int(np.argsort(scores, kind="stable")[0])]
        return scores

    def _external_cost(self, gen_distrib):
        if self._cost_backend is None:
            from .predictor import make_cost_backend
            self._cost_backend = make_cost_backend(gen_distrib.shape, self.agentparams.get("gpu_id", 0))
        return self._cost_backend.score_external(gen_distrib, self._goal_pix, self._hp.finalweight, self._task_weights())

    # -- device path --------------------------------------------------------------------------------------
    def _device_path_ok(self):
        """The whole CEM loop runs on the device for the Gaussian sampler (incl. discrete_ind, append_action, warm start) and
        for the correlated-noise sampler (correlated_noise.py:17-66 without refit_cov / smooth_across_last_action); every
        other sampler or option takes the host plugin path (sampler on the host, rollouts + cost on the device)."""
        hp = self._hp
        if self._backend is None or not hp.device_cem or not hasattr(self._backend, "plan"):
            return False
        n_app = len(hp.append_action) if hp.append_action else 0
        if self._backend.spec.adim != self._sampler_adim() + n_app:
            return False
        if hp.sampler is GaussianCEMSampler:
            return (not hp.rejection_sampling and not hp.cov_blockdiag and not hp.smooth_cov and not hp.reuse_cov
                    and not hp.add_zero_action and not (n_app and hp.reuse_mean))
        if hp.sampler is CorrelatedNoiseSampler:
            return not hp.refit_cov and not hp.smooth_across_last_action
        return False

    def _sampler_adim(self):
        return len(self._hp.initial_std) if self._hp.sampler is CorrelatedNoiseSampler else self._adim

    def perform_CEM(self, state):
        if not self._device_path_ok():
            return super().perform_CEM(state)
        hp = self._hp
        smp = self._sampler
        context = {"context_frames": self._images, "context_actions": smp.chosen_actions,
                   "context_pixel_distributions": self._make_input_distrib(0), "context_states": self._state}
        common = dict(iterations=self._n_iter, goal_pix=self._goal_pix, finalweight=hp.finalweight, task_weights=self._task_weights(),
                      seed=hp.cem_seed, plan_index=self._plan_counter, k_futures=hp.num_futures, lambda_variance=hp.lambda_variance,
                      append_action=hp.append_action if hp.append_action else None)
        if hp.sampler is CorrelatedNoiseSampler:
            M = hp.num_samples
            res = self._backend.plan(
                context, num_samples=M, num_elites=min(self.num_elites(), M), nactions=hp.nactions, repeat=1,
                std=np.asarray(hp.initial_std, np.float64), clip=None, mean0=None, reduce_std_scale=1.0, sampler="correlated",
                beta0=hp.beta_0, beta1=hp.beta_1, kappa=hp.kappa, mean_bias=hp.mean_bias, **common)
        else:
            warm = self._t >= hp.repeat - 1
            mean0, shrink = None, False
            if hp.reuse_mean and warm and smp._best_action_plans and smp._best_action_plans[-1] is not None:
                mean0 = smp._warm_start_mean(smp._best_action_plans[-1][0])
                shrink = True
            M = max(int(hp.num_samples * hp.reuse_factor), 1) if shrink else hp.num_samples
            lo, hi = action_bounds(hp, self._adim)
            res = self._backend.plan(
                context, num_samples=M, num_elites=min(self.num_elites(), M),
                nactions=hp.nactions, repeat=hp.repeat, std=np.sqrt(per_dim_variance(hp, self._adim)),
                clip=(lo, hi) if hp.action_bound else None, mean0=mean0,
                reduce_std_scale=(hp.reduce_std_dev if (self._t is not None and self._t >= 2) else 1.0),
                discrete_ind=hp.discrete_ind, **common)
        self._plan_counter += 1
        self._best_actions, self._best_indices = res["best_actions"], res["elite_idx"]
        for i in range(self._n_iter):
            self.plan_stat["scores_itr{}".format(i)] = res["scores"][i]
        if hp.predictor_propagation:
            self._chosen_distrib = self._backend.fetch_distrib(int(self._best_indices[0]) * max(int(hp.num_futures), 1))
        if self._verbose_condition(self._n_iter - 1):
            self._export_verbose(res["scores"][-1])
        self._t_since_replan = 0

    def act(self, t=None, i_tr=None, desig_pix=None, goal_pix=None, images=None, state=None, verbose_worker=None):
        self._desig_pix = np.array(desig_pix).reshape((self._n_cam, self._n_desig, 2))
        self._goal_pix = np.array(goal_pix).reshape((self._n_cam, self._n_desig, 2))
        self._images = images
        self._verbose_worker = verbose_worker
        return super().act(t, i_tr, state)


class GoalImController(PixelCostController):
    """Goal-image controller (reference ``goal_im_controller.py:12-93``): the cost of a rollout is the mean squared error
    between its FINAL predicted frame of view 0 and the goal image, planning starts at ``t = n_context``
    (:35).  Prediction and cost run on the engine; the sampler loop is the host plugin path.

    Deviations: the reference reads the goal image from a hard-coded file and compares frames in [0,1] with uint8 pixels;
    here the goal image arrives through ``act(goal_image=...)`` (the agent's ``goal_image`` entry, ``general_agent.py:142-151``)
    and is brought to [0,1] like the frames."""

    def __init__(self, ag_params, policyparams, gpu_id=0, ngpu=1):
        super().__init__(ag_params, policyparams, gpu_id, ngpu)
        if self._backend is None:
            raise NotImplementedError("GoalImController needs the engine-backed predictor (the cost runs on the device)")
        self._hp.start_planning = self._net_context
        self._goal_image = None

    def _device_path_ok(self):
        return False                          # the device CEM loop scores pixel distance; this cost goes through the plugin path

    def evaluate_rollouts(self, actions, cem_itr):
        context = {"context_frames": self._images, "context_actions": self._sampler.chosen_actions,
                   "context_pixel_distributions": self._make_input_distrib(cem_itr), "context_states": self._state}
        scores = self._backend.evaluate_goal_image(context, actions, self._goal_image)
        if self._verbose_condition(cem_itr):
            self._export_verbose(scores)
        return scores

    def act(self, t=None, i_tr=None, goal_image=None, images=None, state=None, verbose_worker=None):
        g = np.asarray(goal_image)
        while g.ndim > 3:                      # (T, ncam, H, W, 3) or (ncam, H, W, 3): last time step, view 0
            g = g[-1] if g.ndim == 5 else g[0]
        g = g.astype(np.float32)
        self._goal_image = g / 255.0 if g.max() > 1.5 else g
        zeros = np.zeros((self._n_cam, self._n_desig, 2))
        return super().act(t, i_tr, zeros, zeros, images, state, verbose_worker)

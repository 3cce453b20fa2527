"""Host-side action-sequence samplers: the ``sampler`` plugin point of the CEM controller
(reference ``cem_base_controller.py:52,66-76,82``; plugin contract ``samplers/cem_sampler.py:7-55``).

These run on the host when a user supplies a sampler class; the default planning path samples on the
device (csrc/cem.cu) with the same law.  They draw from the global ``np.random`` stream exactly like
the reference so a seeded run reproduces the reference's action tensors.

Reference functions mirrored (paths under visual_mpc/policy):
  utils/controller_utils.py:6-44   truncate_movement      -> clip_actions
  utils/controller_utils.py:47-84  construct_initial_sigma -> initial_covariance
  utils/controller_utils.py:87-96  reuse_cov              -> shifted_covariance (t=None bug fixed)
  utils/controller_utils.py:99-104 make_blockdiagonal     -> band_mask_covariance
  utils/controller_utils.py:107-117 discretize            -> discretize_actions
  cem_controllers/samplers/gaussian_sampler.py            -> GaussianCEMSampler
  cem_controllers/samplers/correlated_noise.py            -> CorrelatedNoiseSampler
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

_AXIS_STD = {"x": "initial_std", "y": "initial_std", "z": "initial_std_lift", "theta": "initial_std_rot",
             "grasp": "initial_std_grasp"}


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
def action_bounds(hp, adim: int):
    """Per-dimension (lo, hi) clip bounds; +-inf where the reference does not clip."""
    lo = np.full(adim, -np.inf)
    hi = np.full(adim, np.inf)
    order = hp.get("action_order") if "action_order" in hp else None
    if order is not None:
        for i, axis in enumerate(order):
            if axis in ("x", "y"):
                lo[i], hi[i] = -2.0 * hp.initial_std, 2.0 * hp.initial_std
            elif axis == "theta":
                lo[i], hi[i] = -math.pi / 4, math.pi / 4
        return lo, hi
    n = min(2, adim)
    lo[:n], hi[:n] = -2.0 * hp.initial_std, 2.0 * hp.initial_std
    if adim >= 4:
        lo[3], hi[3] = -math.pi / 4, math.pi / 4
    return lo, hi


def clip_actions(actions: np.ndarray, hp) -> np.ndarray:
    """In-place clip of xy displacement to +-2 sigma_xy and rotation to +-pi/4 on the last axis."""
    if actions.ndim not in (2, 3):
        raise NotImplementedError("actions must be (M, adim) or (M, n, adim)")
    lo, hi = action_bounds(hp, actions.shape[-1])
    for d in range(actions.shape[-1]):
        if np.isfinite(lo[d]) or np.isfinite(hi[d]):
            actions[..., d] = np.clip(actions[..., d], lo[d], hi[d])
    return actions


def per_dim_variance(hp, adim: int) -> List[float]:
    order = hp.get("action_order") if "action_order" in hp else None
    if order is not None:
        var = []
        for axis in order:
            if axis not in _AXIS_STD:
                raise NotImplementedError(axis)
            var.append(getattr(hp, _AXIS_STD[axis]) ** 2)
        return var
    var = [hp.initial_std ** 2, hp.initial_std ** 2]
    if adim >= 3:
        var.append(hp.initial_std_lift ** 2)
    if adim >= 4:
        var.append(hp.initial_std_rot ** 2)
    if adim == 5:
        var.append(hp.initial_std_grasp ** 2)
    return var


def initial_covariance(hp, adim: int, t: Optional[int] = None) -> np.ndarray:
    var = per_dim_variance(hp, adim)
    block = len(var)
    diag = np.array(np.tile(var, hp.nactions))
    if "reduce_std_dev" in hp:
        assert "reuse_mean" in hp
        if t is not None and t >= 2:
            # every block but the last one can be warm-started, so its spread is reduced
            diag[:(hp.nactions - 1) * block] *= hp.reduce_std_dev
    return np.diag(diag)


def shifted_covariance(sigma: np.ndarray, adim: int, hp) -> np.ndarray:
    """Shift the previous step's covariance one action forward and blend in a fraction of the
    initial one (reference reuse_cov; its `t=None >= 2` TypeError on Python 3 is not reproduced)."""
    assert hp.replan_interval == 3
    init = initial_covariance(hp, adim, None)
    out = np.zeros_like(sigma)
    out[:-adim, :-adim] = sigma[adim:, adim:] + init[:-adim, :-adim] * hp.reuse_cov
    out[-adim:, -adim:] = init[:adim, :adim]
    return out


def band_mask_covariance(cov: np.ndarray, nactions: int, adim: int) -> np.ndarray:
    mask = np.zeros_like(cov)
    for i in range(nactions - 1):
        mask[i * adim:(i + 2) * adim, i * adim:(i + 2) * adim] = 1.0
    return cov * mask


def discretize_actions(actions: np.ndarray, discrete_ind) -> np.ndarray:
    for ind in discrete_ind:
        actions[:, :, ind] = np.clip(np.floor(actions[:, :, ind]), 0, 4)
    return actions


# ---------------------------------------------------------------------------------------------------
class CEMSampler(object):
    """Plugin base (reference cem_sampler.py:7-55)."""

    def __init__(self, hp, adim, sdim, **kwargs):
        self._hp = hp
        self._adim, self._sdim = adim, sdim
        self._chosen_actions = []
        self._best_action_plans = []

    def sample_initial_actions(self, t, nsamples, current_state):
        raise NotImplementedError

    def sample_next_actions(self, n_samples, best_actions, scores):
        raise NotImplementedError

    def log_best_action(self, action, best_action_plans):
        self._chosen_actions.append(np.array(action, copy=True))
        self._best_action_plans.append(best_action_plans)

    @property
    def chosen_actions(self):
        return np.array(self._chosen_actions)

    @staticmethod
    def get_default_hparams():
        return {}


class GaussianCEMSampler(CEMSampler):
    def __init__(self, hp, adim, sdim, **kwargs):
        super().__init__(hp, adim, sdim, **kwargs)
        self._sigma = self._sigma_prev = self._mean = None
        self._last_reduce = None

    @staticmethod
    def get_default_hparams():
        return dict(action_order=None, initial_std=0.05, initial_std_lift=0.15, initial_std_rot=np.pi / 18,
                    initial_std_grasp=2, discrete_ind=None, reuse_mean=False, reduce_std_dev=1., reuse_cov=False,
                    rejection_sampling=True, cov_blockdiag=False, smooth_cov=False, nactions=5, repeat=3,
                    add_zero_action=False, action_bound=True, reuse_factor=0.5)

    # -- public plugin API --------------------------------------------------------------------------
    def sample_initial_actions(self, t, nsamples, current_state):
        hp = self._hp
        warm = t >= hp.repeat - 1
        shrink = False
        if hp.reuse_cov and warm and self._sigma is not None:
            self._sigma = shifted_covariance(self._sigma, self._adim, hp)
            shrink = True
        else:
            self._sigma = initial_covariance(hp, self._adim, t)
        self._sigma_prev = self._sigma

        if hp.reuse_mean and warm and self._mean is not None:
            assert self._best_action_plans[-1] is not None, "Cannot reuse mean if best actions are not logged!"
            self._mean = self._warm_start_mean(self._best_action_plans[-1][0])
            shrink = True
        else:
            self._mean = np.zeros(self._adim * hp.nactions)
        self._last_reduce = shrink
        return self._draw(nsamples, shrink)

    def sample_next_actions(self, n_samples, best_actions, scores):
        self._fit(best_actions)
        return self._draw(n_samples, self._last_reduce)

    # -- internals ------------------------------------------------------------------------------------
    def _warm_start_mean(self, plan):
        hp = self._hp
        rem = plan.shape[0] % hp.repeat
        if rem:
            plan = np.concatenate((plan, np.zeros((hp.repeat - rem, self._adim))), axis=0)
        first_of_group = plan.reshape(-1, hp.repeat, self._adim)[:, 0]
        mean = np.zeros((hp.nactions, self._adim))
        mean[:first_of_group.shape[0]] = first_of_group
        return mean.reshape(-1)

    def _draw(self, count, shrink):
        hp = self._hp
        if shrink:
            count = max(int(count * hp.reuse_factor), 1)
        if hp.rejection_sampling:
            return self._draw_rejection(count)
        seq = np.random.multivariate_normal(self._mean, self._sigma, count).reshape(count, hp.nactions, self._adim)
        if hp.discrete_ind is not None:
            seq = discretize_actions(seq, hp.discrete_ind)
        if hp.action_bound:
            seq = clip_actions(seq, hp)
        seq = np.repeat(seq, hp.repeat, axis=1)
        if hp.add_zero_action:
            seq[0] = 0
        return seq

    def _fit(self, elites):
        hp = self._hp
        per_group = elites.reshape(-1, hp.nactions, hp.repeat, self._adim)[:, :, -1]
        flat = per_group.reshape(per_group.shape[0], hp.nactions * self._adim)
        sigma = np.cov(flat, rowvar=False, bias=False)
        if hp.cov_blockdiag:
            sigma = band_mask_covariance(sigma, hp.nactions, self._adim)
        if hp.smooth_cov:
            sigma = 0.5 * sigma + 0.5 * self._sigma_prev
            self._sigma_prev = sigma
        self._sigma = sigma
        self._mean = flat.mean(axis=0)

    def _draw_rejection(self, count):
        """Per-sample redraw until xy and z stay within 1.5 sigma (reference gaussian_sampler.py:109-150).
        ``stochastic_planning`` is read with a default because the reference never declares it."""
        hp = self._hp
        lim_xy, lim_z = 1.5 * hp.initial_std, 1.5 * hp.initial_std_lift
        rows = []
        for _ in range(count):
            while True:
                cand = np.random.multivariate_normal(self._mean, self._sigma, 1).reshape(hp.nactions, self._adim)
                if np.all(np.abs(cand[:, :2]) <= lim_xy) and (self._adim < 3 or np.all(np.abs(cand[:, 2]) <= lim_z)):
                    break
            rows.append(cand)
        seq = np.stack(rows, axis=0)
        stoch = hp.get("stochastic_planning") if "stochastic_planning" in hp else None
        if stoch:
            seq = np.repeat(seq, stoch[0], 0)
        if hp.discrete_ind is not None:
            seq = discretize_actions(seq, hp.discrete_ind)
        return np.repeat(seq, hp.repeat, axis=1)


class CorrelatedNoiseSampler(CEMSampler):
    """AR(1)-smoothed Gaussian noise around a softmax-weighted elite mean."""

    def __init__(self, hp, adim, sdim, **kwargs):
        super().__init__(hp, len(hp.initial_std), sdim, **kwargs)

    @staticmethod
    def get_default_hparams():
        return dict(nactions=15, initial_std=[0.05, 0.05, 0.2, np.pi / 10], mean_bias=None, kappa=1, beta_0=0.5,
                    beta_1=0.5, smooth_across_last_action=False, refit_cov=False)

    def _noise(self, count, cov=None):
        hp = self._hp
        eps = np.random.normal(size=(count, hp.nactions, self._adim))
        bias = np.zeros(self._adim) if hp.mean_bias is None else np.asarray(hp.mean_bias)
        if cov is None:
            eps = eps * np.asarray(hp.initial_std).reshape(1, 1, -1) + bias[None, None]
        else:
            eps = (eps.reshape(count, -1) @ cov).reshape(count, hp.nactions, self._adim)
        out = eps.copy()
        for i in range(hp.nactions):
            if hp.smooth_across_last_action and i == 0 and len(self._chosen_actions):
                prev = np.asarray(self._chosen_actions[-1])[None]
            else:
                prev = out[:, i - 1]        # i == 0 wraps to the (still un-smoothed) last step, as in the reference
            out[:, i] = hp.beta_0 * eps[:, i] + hp.beta_1 * prev
        return out

    def sample_initial_actions(self, t, n_samples, current_state):
        return self._noise(n_samples)

    def sample_next_actions(self, n_samples, best_actions, scores):
        hp = self._hp
        reward = -np.asarray(scores)
        weight = np.exp(hp.kappa * (reward - reward.max()))
        mean = (best_actions * weight[:, None, None]).sum(0) / (weight.sum() + 1e-4)
        cov = np.cov(best_actions.reshape(best_actions.shape[0], -1).T) if hp.refit_cov else None
        return self._noise(n_samples, cov) + mean.reshape(1, best_actions.shape[1], self._adim)


class AutograspSampler(GaussianCEMSampler):
    """Gaussian CEM over the arm dimensions with a rule-based gripper command appended as the last action dimension
    (reference samplers/autograsp_sampler.py:5-58): the gripper closes from the first step at which the integrated
    z displacement (times ``action_norm_factor``) brings the arm below ``z_thresh`` and, unless ``reopen``, stays closed;
    ``deviation_prob`` flips single steps.  With ``no_refit`` (default) the rule is re-applied to every resampled batch,
    otherwise the gripper column is drawn per step from the elites' closing frequency.

    Deviation: the reference's ``sample_next_actions`` calls the parent without the ``scores`` argument and raises
    ``TypeError`` (autograsp_sampler.py:26); here the call is well-formed."""

    def __init__(self, hp, adim, sdim, **kwargs):
        super().__init__(hp, adim - 1, sdim, **kwargs)
        self._current_state = None

    @staticmethod
    def get_default_hparams():
        hp = GaussianCEMSampler.get_default_hparams()
        hp.update(deviation_prob=0, reopen=False, action_norm_factor=1.0, z_thresh=0.15, gripper_close_cmd=1,
                  gripper_open_cmd=-1, no_refit=True)
        return hp

    def sample_initial_actions(self, t, nsamples, current_state):
        self._current_state = current_state
        return self._with_gripper_rule(super().sample_initial_actions(t, nsamples, current_state))

    def sample_next_actions(self, n_samples, best_actions, scores):
        hp = self._hp
        arm = super().sample_next_actions(n_samples, best_actions[:, :, :-1], scores)
        if hp.no_refit:
            return self._with_gripper_rule(arm)
        p_close = (best_actions[:, :, -1] == hp.gripper_close_cmd).astype(np.float32).mean(axis=0)
        grip = np.zeros((arm.shape[0], arm.shape[1], 1), dtype=np.float32)
        for step in range(arm.shape[1]):                       # one uniform vector per step, like the reference
            close = np.random.uniform(size=arm.shape[0]) < p_close[step]
            grip[:, step, 0] = np.where(close, hp.gripper_close_cmd, hp.gripper_open_cmd)
        return np.concatenate((arm, grip), axis=-1)

    def _with_gripper_rule(self, arm):
        hp = self._hp
        z0 = self._current_state[2]
        closed = np.cumsum(arm[:, :, 2] * hp.action_norm_factor, axis=1) + z0 < hp.z_thresh
        grip = np.zeros((arm.shape[0], arm.shape[1], 1))
        for row in range(arm.shape[0]):                        # row order matters: deviation noise is drawn per sample
            mask = closed[row].copy()
            if not hp.reopen and mask.any():
                mask[int(np.argmax(mask)):] = True
            if hp.deviation_prob:
                flip = np.random.uniform(size=mask.shape[0]) < hp.deviation_prob
                mask = np.logical_xor(mask, flip)
            grip[row, :, 0] = np.where(mask, hp.gripper_close_cmd, hp.gripper_open_cmd)
        return np.concatenate((arm, grip), axis=-1)


class FoldingCEMSampler(CEMSampler):
    """Cloth-folding proposal distribution (reference samplers/folding_sampler.py:7-132).  A ``split_frac`` share of the
    samples (halved on the first iteration) follows two scripted motion templates built from uniformly drawn way-points —
    (reach, lower, lift, carry, lower) and (lift, carry, lower, hold) with tight covariance on the vertical moves — and the
    rest is drawn from the fitted Gaussian; xyz are clipped to ``max_shift``.  The ``np.random`` draw order is the
    reference's, so a seeded run reproduces its action tensors."""

    _REACH_CARRY = ((+1.0, "wide", 0), (-1.0, "tight", None), (+1.0, "tight", None), (+1.0, "wide", 1), (-1.0, "tight", None))

    def __init__(self, hp, adim, sdim, **kwargs):
        super().__init__(hp, adim, sdim, **kwargs)
        assert adim == 4, "Requires base action dimension of 4"
        assert hp.nactions >= 5, "Requires at least 5 steps"
        self._xy = None
        self._mean = self._cov = None

    @staticmethod
    def get_default_hparams():
        return dict(action_order=None, initial_std=0.05, initial_std_lift=0.15, initial_std_rot=np.pi / 18,
                    initial_std_grasp=2, nactions=5, repeat=3, max_shift=[1. / 5, 1. / 5, 1. / 3], split_frac=0.5)

    def sample_initial_actions(self, t, n_samples, current_state):
        self._xy = np.asarray(current_state)[:2]
        return self._propose(n_samples, np.zeros(self._hp.nactions * self._adim), initial_covariance(self._hp, self._adim, t))

    def sample_next_actions(self, n_samples, best_actions, scores):
        hp = self._hp
        last = best_actions.reshape(-1, hp.nactions, hp.repeat, self._adim)[:, :, -1]
        flat = last.reshape(last.shape[0], hp.nactions * self._adim)
        return self._propose(n_samples, flat.mean(axis=0), np.cov(flat, rowvar=False, bias=False))

    def _propose(self, count, mean, cov):
        hp, steps, rep = self._hp, self._hp.nactions, self._hp.repeat
        assert count % 3 == 0, "splits samples into setting with 3 means"
        self._mean, self._cov = np.array(mean, copy=True), np.array(cov, copy=True)
        wide = self._cov[:4, :4]
        tight = wide.copy()
        tight[:2, :2] /= 10
        tight[3, 3] /= 2
        covs = {"wide": wide, "tight": tight}
        n_tpl = max(int(int(count * hp.split_frac / 2) / 2), 1)     # both public entry points are "first iteration" calls

        def draw(mu, which):
            return np.random.multivariate_normal(np.asarray(mu, dtype=np.float64), covs[which], 1).reshape(-1)

        out = np.zeros((count, steps, self._adim))
        for row in range(n_tpl):                                 # template 1: reach, lower, lift, carry, lower
            p1, p2 = np.random.uniform(size=2), np.random.uniform(size=2)
            legs = ((p1 - self._xy) / rep, (p2 - p1) / rep)
            for step, (dz, which, leg) in enumerate(self._REACH_CARRY):
                dxy = legs[leg] if leg is not None else (0.0, 0.0)
                out[row, step] = draw([dxy[0], dxy[1], dz, 0.0], which)
            if steps > 5:                                        # the reference assigns these draws to an empty slice
                np.random.multivariate_normal(np.zeros(4), wide, steps - 5)
        for row in range(n_tpl, 2 * n_tpl):                      # template 2: lift, carry, lower, hold
            carry = (np.random.uniform(size=2) - self._xy) / rep
            out[row, 0] = draw([0.0, 0.0, 1.0, 0.0], "tight")
            out[row, 1] = draw([carry[0], carry[1], 1.0, 0.0], "wide")
            out[row, 2] = draw([0.0, 0.0, -1.0, 0.0], "tight")
            out[row, 3:] = draw([0.0, 0.0, 0.0, 0.0], "tight")
            if steps > 5:
                np.random.multivariate_normal(np.zeros(4), wide, steps - 5)
        rest = count - 2 * n_tpl
        out[2 * n_tpl:] = np.random.multivariate_normal(self._mean, self._cov, rest).reshape(rest, steps, self._adim)
        lim = np.asarray(hp.max_shift)
        out[:, :, :3] = np.clip(out[:, :, :3], -lim, lim)
        return np.repeat(out, rep, axis=1)

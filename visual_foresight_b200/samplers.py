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

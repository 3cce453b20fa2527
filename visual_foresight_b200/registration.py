"""Registration front end of the pixel-distance cost: where are the designated pixels NOW, and how much should each
registered copy of a task count?  (reference ``policy/cem_controllers/register_gtruth_controller.py``.)

The reference warps the first frame of the trajectory ("start") and the goal image onto the current frame with a
registration network (``registration_network.setup_registration.setup_gdn`` — not in the reference tree),
reads the designated pixel of every task through the warp fields, and weighs the resulting ``ntask x nreg`` designated
pixels by the inverse of their warp errors (``register_gtruth`` :54-112, ``get_warp_err`` :114-173).  The weights are the
``task_weights`` of the device cost.  Here the warper is a user-supplied callable with the reference's
signature ``warper(current[None], other[None]) -> (warped_image, flow, warp_pts)``; the arithmetic around it is restated
and pinned to the unmodified reference's ``get_warp_err`` / normalisation (tests/golden/make_registration_golden.py).

Deviations (documented, tested):
  * the reference module cannot be imported as shipped (``visualizer.render_utils``, ``visualizer.make_cem_visuals`` and
    ``registration_network`` are missing) and ``_prep_vidpred_inp`` has no parent implementation; the controller below
    performs the registration at the first CEM iteration of every plan, which is what that hook did;
  * with ``register_region=False`` the reference leaves the warp errors at zero (its point-wise branch is guarded by a
    constant) and the trade-off becomes inf/NaN; the point-wise colour distance — the evident intent — is used instead.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from .cem_controller import PixelCostController


def _window(center: int, half: int, hi: int) -> Tuple[int, int]:
    lo_, hi_ = np.clip(np.array((center - half, center + half + 1)), 0, hi)
    return int(lo_), int(hi_)


def warp_errors(icam: int, start_image, goal_image, start_warp_pts, goal_warp_pts, warped_start, warped_goal, desig_pix_t0,
                goal_pix, *, register: Sequence[str] = ("start", "goal"), region: bool = False, agent_height: int,
                agent_width: int, net_height: int):
    """Warp error and registered designated pixel of every (task, registered image) pair for camera ``icam``.

    desig_pix_t0, goal_pix: (ncam, ntask, 2) integer (row, col) in the agent's image resolution.
    *_warp_pts: (ncam, H, W, 2) warp fields in (x, y) order;  images (ncam, H, W, 3) float.
    Returns warperrs (ntask, nreg) and desig (ntask, nreg, 2) as (row, col), rescaled to the predictor's resolution."""
    nreg, ntask = len(register), np.asarray(desig_pix_t0).shape[1]
    errs = np.zeros((ntask, nreg))
    desig = np.zeros((ntask, nreg, 2))
    half = 5 if agent_height >= 96 else 2
    sources = {"start": (0, start_image, start_warp_pts, warped_start, desig_pix_t0),
               "goal": (1, goal_image, goal_warp_pts, warped_goal, goal_pix)}
    for p in range(ntask):
        for name in register:
            slot, image, pts, warped, pix = sources[name]
            r, c = int(pix[icam][p][0]), int(pix[icam][p][1])
            if region:
                # the start window is clipped to size-1 and the goal window to size, as in the reference (:142-143, :154-155)
                lim_r, lim_c = (agent_height - 1, agent_width - 1) if name == "start" else (agent_height, agent_width)
                r0, r1 = _window(r, half, lim_r)
                c0, c1 = _window(c, half, lim_c)
                errs[p, slot] = np.mean(np.square(image[icam][r0:r1, c0:c1] - warped[icam][r0:r1, c0:c1]))
                field = pts[icam][r0:r1, c0:c1]
                desig[p, slot] = (np.median(field[:, :, 1]), np.median(field[:, :, 0]))      # (x, y) -> (row, col)
            else:
                desig[p, slot] = pts[icam][r, c][::-1]
                errs[p, slot] = np.linalg.norm(image[icam][r, c] - warped[icam][r, c])
    return errs, desig * net_height / agent_height


def registration_tradeoff(warperrs) -> np.ndarray:
    """(ncam, ntask, nreg) warp errors -> weights ~ 1/err, normalised per task over cameras and registered images
    (reference :88-91); returned as (ncam, ntask * nreg) = (ncam, ndesig)."""
    w = 1.0 / np.asarray(warperrs, dtype=np.float64)
    w = w / w.sum(axis=0, keepdims=True).sum(axis=2, keepdims=True)
    return w.reshape(w.shape[0], -1)


class RegisterGtruthController(PixelCostController):
    """PixelCostController whose designated pixels and task weights come from registering the trajectory's first frame and
    the goal image onto the current frame.  policyparams: ``goal_image_warper`` (callable, required), ``register_gtruth``
    (subset of ['start', 'goal']), ``register_region``; ``designated_pixel_count`` must equal ntask * len(register_gtruth)."""

    def __init__(self, ag_params, policyparams, gpu_id=0, ngpu=1):
        super().__init__(ag_params, policyparams, gpu_id, ngpu)
        hp = self._hp
        if hp.goal_image_warper is None:
            raise ValueError("RegisterGtruthController needs policyparams['goal_image_warper']")
        self._warper: Callable = hp.goal_image_warper
        self._nreg = len(hp.register_gtruth)
        assert self._nreg and self._n_desig % self._nreg == 0
        self._ntask = self._n_desig // self._nreg
        self.reg_tradeoff = np.full((self._n_cam, self._n_desig), 1.0 / (self._n_cam * self._n_desig))
        self._start_image = self._goal_image = None
        self._desig_t0 = self._goal_sel = None

    def _default_hparams(self):
        hp = super()._default_hparams()
        hp.add_hparam("register_gtruth", ["start", "goal"])
        hp.add_hparam("register_region", False)
        hp.add_hparam("goal_image_warper", None)
        return hp

    def _task_weights(self):
        return np.asarray(self.reg_tradeoff, dtype=np.float64).reshape(-1)

    def register(self, current_frame):
        """current_frame: (ncam, H, W, 3) float in [0,1].  Updates the designated pixels and the trade-off."""
        hp = self._hp
        H, W = self._start_image.shape[1:3]
        w_start, _, p_start = self._warper(current_frame[None], self._start_image[None])
        w_start = np.asarray(w_start).reshape(self._n_cam, H, W, 3)
        p_start = np.asarray(p_start).reshape(self._n_cam, H, W, 2)
        w_goal = p_goal = None
        if "goal" in hp.register_gtruth:
            w_goal, _, p_goal = self._warper(current_frame[None], self._goal_image[None])
            w_goal = np.asarray(w_goal).reshape(self._n_cam, H, W, 3)
            p_goal = np.asarray(p_goal).reshape(self._n_cam, H, W, 2)
        errs, pix = [], []
        for cam in range(self._n_cam):
            e, d = warp_errors(cam, self._start_image, self._goal_image, p_start, p_goal, w_start, w_goal, self._desig_t0,
                               self._goal_sel, register=hp.register_gtruth, region=hp.register_region,
                               agent_height=self.agentparams["image_height"], agent_width=self.agentparams["image_width"],
                               net_height=self._img_height)
            errs.append(e)
            pix.append(d)
        errs = np.stack(errs, 0)
        self._desig_pix = np.stack(pix, 0).reshape(self._n_cam, self._n_desig, 2)
        self.reg_tradeoff = registration_tradeoff(errs)
        self.plan_stat["tradeoff"] = self.reg_tradeoff
        self.plan_stat["warperrs"] = errs.reshape(self._n_cam, self._n_desig)

    def perform_CEM(self, state):
        self.register(np.asarray(self._images[-1], dtype=np.float32) / 255.0)
        return super().perform_CEM(state)

    def act(self, goal_image=None, t=None, i_tr=None, desig_pix=None, goal_pix=None, images=None, state=None, verbose_worker=None):
        self._goal_sel = np.array(goal_pix).reshape((self._n_cam, self._ntask, 2))
        tiled = np.tile(self._goal_sel[:, :, None, :], [1, 1, self._nreg, 1]).reshape(self._n_cam, self._n_desig, 2)
        g = np.asarray(goal_image)
        self._goal_image = (g[-1] if g.ndim == 5 else g).astype(np.float32)
        if self._goal_image.max() > 1.5:
            self._goal_image = self._goal_image / 255.0
        if t == 0 or self._desig_t0 is None:
            self._desig_t0 = np.array(desig_pix).reshape((self._n_cam, self._ntask, 2))
            self._start_image = np.asarray(images[0], dtype=np.float32) / 255.0
        # the parent reshapes desig_pix to (ncam, ndesig, 2); until the first registration the t0 pixels stand in for every copy
        d0 = np.tile(self._desig_t0[:, :, None, :], [1, 1, self._nreg, 1]).reshape(self._n_cam, self._n_desig, 2)
        return super().act(t, i_tr, d0, tiled, images, state, verbose_worker)

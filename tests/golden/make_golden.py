"""Generates tests/golden/ref_cem_golden.npz by running the UNMODIFIED reference CEM / cost / sampler
code (under ref_shim) on seeded inputs.  Run in the authoring container only:

    python tests/golden/make_golden.py

The reference tree does not exist on the GPU box, so the outputs are committed as fixtures.
Reference functions exercised (paths relative to /root/reference/visual_mpc):
  policy/cem_controllers/pixel_cost_controller.py:135-215  (_eval_pixel_cost, _expected_distance,
                                                            _get_distancegrid, _switch_on_pix)
  policy/cem_controllers/cem_base_controller.py:85-169     (perform_CEM, act)
  policy/cem_controllers/samplers/gaussian_sampler.py      (_sample_actions, _fit_gaussians, sample_initial_actions)
  policy/cem_controllers/samplers/correlated_noise.py      (_sample_noise, sample_next_actions)
  policy/utils/controller_utils.py                          (construct_initial_sigma, truncate_movement,
                                                            make_blockdiagonal, discretize, reuse_cov)
  policy/policy.py:9-63                                     (get_policy_args, _override_defaults)
  video_prediction/pred_util.py:4-48                        (get_context, rollout_predictions)
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from fake_predictor import BlobPredictor  # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    ref_shim.install()
    from visual_mpc.policy.cem_controllers import PixelCostController
    from visual_mpc.policy.cem_controllers.samplers import GaussianCEMSampler, CorrelatedNoiseSampler
    from visual_mpc.policy.policy import get_policy_args
    from visual_mpc.policy.utils import controller_utils as cu
    from visual_mpc.video_prediction import pred_util
    HP = ref_shim._HParams

    G = {}
    rng = np.random.RandomState(1234)

    # ---------------- cost path ---------------------------------------------------------------
    H, W = 48, 64
    ag = {"adim": 4, "sdim": 4, "image_height": H, "image_width": W, "gpu_id": 0}
    pp = {"predictor_class": BlobPredictor, "rejection_sampling": False, "verbose": False,
          "num_samples": 24, "minimum_selection": 5}
    ctrl = quiet(PixelCostController, ag, dict(pp), 0, 1)
    quiet(ctrl.reset)
    for gi, goal in enumerate([np.array([36, 48]), np.array([0, 0]), np.array([47.0, 10.0])]):
        G["distgrid_goal%d" % gi] = goal.astype(np.float64)
        G["distgrid_out%d" % gi] = quiet(ctrl._get_distancegrid, goal)
    agc = dict(ag, image_height=32, image_width=32)          # small planes keep the fixture < 1 MB
    gen = (rng.rand(8, 13, 1, 32, 32, 2).astype(np.float32) ** 3) + 1e-4
    G["cost_gen_distrib"] = gen
    G["cost_goal_pix"] = np.array([[[18, 24], [5, 7]]])
    ctrl2 = quiet(PixelCostController, agc, dict(pp, designated_pixel_count=2), 0, 1)
    quiet(ctrl2.reset)
    ctrl2._goal_pix = G["cost_goal_pix"]
    G["cost_scores_2desig"] = quiet(ctrl2._eval_pixel_cost, 0, gen, None)
    dg = quiet(ctrl2._get_distancegrid, G["cost_goal_pix"][0, 0])
    G["cost_expected_distance_task0"] = quiet(ctrl2._expected_distance, 0, 0, gen[:, :, 0, :, :, 0], dg)
    ctrl3 = quiet(PixelCostController, agc, dict(pp, designated_pixel_count=2, finalweight=3.0), 0, 1)
    ctrl3._goal_pix = G["cost_goal_pix"]
    G["cost_scores_2desig_fw3"] = quiet(ctrl3._eval_pixel_cost, 0, gen, None)
    desig = np.array([[[12.7, 70.2], [-3, 5.5]]])
    G["switch_desig"] = desig
    ctrl4 = quiet(PixelCostController, ag, dict(pp, designated_pixel_count=2), 0, 1)
    G["switch_onehot"] = quiet(ctrl4._switch_on_pix, desig)

    # ---------------- controller utils -------------------------------------------------------
    def ghp(**over):
        d = GaussianCEMSampler.get_default_hparams()
        d.update(replan_interval=0)
        d.update(over)
        return HP(**d)

    for adim in (2, 3, 4, 5):
        for t in (0, 3):
            G["sigma0_adim%d_t%d" % (adim, t)] = quiet(cu.construct_initial_sigma, ghp(reduce_std_dev=0.5), adim, t)
    G["sigma0_order"] = quiet(cu.construct_initial_sigma, ghp(action_order=["x", "y", "z", "theta", "grasp"]), 5, 0)
    a3 = rng.randn(6, 5, 4) * 0.2
    a3[:, :, 3] *= 10
    G["trunc_in3"] = a3.copy()
    G["trunc_out3"] = cu.truncate_movement(a3.copy(), ghp())
    a2 = rng.randn(7, 4) * 0.3
    a2[:, 3] *= 5
    G["trunc_in2"] = a2.copy()
    G["trunc_out2"] = cu.truncate_movement(a2.copy(), ghp())
    G["trunc_out3_order"] = cu.truncate_movement(a3.copy(), ghp(action_order=["z", "x", "theta", "y"]))
    cov = np.cov(rng.randn(40, 20), rowvar=False)
    G["blockdiag_in"] = cov
    G["blockdiag_out"] = cu.make_blockdiagonal(cov, 5, 4)
    dz = rng.randn(3, 5, 4) * 3
    G["discretize_in"] = dz.copy()
    G["discretize_out"] = cu.discretize(dz.copy(), 3, 5, [2, 3])
    # cu.reuse_cov is not exercisable: it calls construct_initial_sigma(hp, adim) with t=None, and
    # `None >= 2` raises on Python 3 (controller_utils.py:78,94) — reference bug, documented in DESIGN.md.

    # ---------------- gaussian sampler ----------------------------------------------------------
    smp = GaussianCEMSampler(ghp(rejection_sampling=False), 4, 4)
    np.random.seed(7)
    G["gauss_init_actions_seed7"] = quiet(smp.sample_initial_actions, 1, 16, None)
    elites = np.repeat(rng.randn(10, 5, 4) * 0.1, 3, axis=1)
    G["gauss_fit_elites"] = elites
    smp._fit_gaussians(elites)
    G["gauss_fit_mean"], G["gauss_fit_sigma"] = smp._mean.copy(), smp._sigma.copy()
    np.random.seed(11)
    G["gauss_next_actions_seed11"] = quiet(smp.sample_next_actions, 16, elites, np.arange(10.0))
    smp2 = GaussianCEMSampler(ghp(rejection_sampling=False, cov_blockdiag=True, smooth_cov=True), 4, 4)
    np.random.seed(7)
    quiet(smp2.sample_initial_actions, 1, 16, None)
    smp2._fit_gaussians(elites)
    G["gauss_fit_sigma_blockdiag_smooth"] = smp2._sigma.copy()
    # reuse_mean path
    smp3 = GaussianCEMSampler(ghp(rejection_sampling=False, reuse_mean=True), 4, 4)
    plan = rng.randn(10, 13, 4) * 0.05
    smp3.log_best_action(np.zeros(4), plan)
    np.random.seed(3)
    G["gauss_reuse_plan"] = plan
    G["gauss_reuse_actions_seed3"] = quiet(smp3.sample_initial_actions, 4, 16, None)
    G["gauss_reuse_mean"] = smp3._mean.copy()

    # ---------------- correlated noise sampler --------------------------------------------------
    chp = HP(**CorrelatedNoiseSampler.get_default_hparams())
    cs = CorrelatedNoiseSampler(chp, 4, 4)
    np.random.seed(5)
    G["corr_init_seed5"] = quiet(cs.sample_initial_actions, 1, 12, None)
    best = rng.randn(6, 15, 4) * 0.1
    sc = rng.rand(6)
    G["corr_best"], G["corr_scores"] = best, sc
    np.random.seed(6)
    G["corr_next_seed6"] = quiet(cs.sample_next_actions, 12, best, sc)

    # ---------------- full act() loop with the blob predictor ---------------------------------
    pol = quiet(PixelCostController, ag, dict(pp), 0, 1)
    quiet(pol.reset)
    images = rng.randint(0, 256, size=(3, 1, H, W, 3)).astype(np.uint8)
    state = rng.uniform(-0.5, 0.5, size=(3, 4))
    G["act_images"], G["act_state"] = images, state
    G["act_desig"], G["act_goal"] = np.array([[12, 16]]), np.array([[30, 40]])
    np.random.seed(42)
    outs = []
    for t in range(3):
        obs = {"images": images[:t + 1], "state": state[:t + 1]}
        step = {"desig_pix": G["act_desig"], "goal_pix": G["act_goal"]}
        kw = get_policy_args(pol, obs, t, 0, step)
        assert sorted(kw.keys()) == sorted(["t", "i_tr", "desig_pix", "goal_pix", "images", "state", "verbose_worker"])
        o = quiet(pol.act, **kw)
        outs.append(o)
        G["act_t%d_action" % t] = np.array(o["actions"])
        for k, v in o["plan_stat"].items():
            G["act_t%d_%s" % (t, k)] = np.array(v)
        if t >= 1:
            G["act_t%d_best_indices" % t] = np.array(pol._best_indices)
            G["act_t%d_best_actions" % t] = np.array(pol._best_actions)
            call = pol.predictor.calls[-1]
            G["act_t%d_ctx_actions" % t] = call["context_actions"]
            G["act_t%d_ctx_distrib" % t] = call["context_pixel_distributions"]
            G["act_t%d_last_actions" % t] = call["actions"]
    G["act_start_planning"] = np.array(pol._hp.start_planning)

    # ---------------- pred_util -----------------------------------------------------------------
    ims = rng.randint(0, 256, size=(5, 2, 8, 8, 3)).astype(np.uint8)
    sts = rng.randn(5, 3)
    lf, ls = pred_util.get_context(2, 3, sts, ims, HP(state_append=[0.1, 0.2]))
    G["ctx_images"], G["ctx_states"] = ims, sts
    G["ctx_last_frames"], G["ctx_last_states"] = lf, ls

    seen = []

    def pf(input_images=None, input_state=None, input_actions=None, input_one_hot_images=None):
        seen.append(input_actions.copy())
        s = input_actions.sum(axis=(1, 2))
        return s[:, None] * np.ones((1, 2)), None, s[:, None] * 2.0

    acts = rng.randn(7, 3, 2)
    gi, gd, gs = pred_util.rollout_predictions(pf, 3, acts, lf, ls)
    G["rollout_actions"] = acts
    G["rollout_gen_images"] = np.concatenate(gi, 0)
    G["rollout_gen_states"] = np.concatenate(gs, 0)
    G["rollout_seen_last"] = seen[-1]
    G["rollout_ncalls"] = np.array(len(seen))

    # ---------------- override semantics --------------------------------------------------------
    msgs = []
    for bad in ({"iterations": 3}, {"not_a_param": 1}):
        try:
            quiet(PixelCostController, ag, dict(pp, **bad), 0, 1)
            msgs.append("ok")
        except Exception as e:  # noqa: BLE001
            msgs.append(type(e).__name__)
    G["override_errors"] = np.array(msgs)

    out = os.path.join(HERE, "ref_cem_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, "keys:", len(G), "bytes:", os.path.getsize(out))


if __name__ == "__main__":
    main()

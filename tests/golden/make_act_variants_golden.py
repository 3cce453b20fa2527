"""Generates tests/golden/ref_act_variants_golden.npz: the UNMODIFIED reference PixelCostController (under ref_shim, with the
shared BlobPredictor injected) driven for several MPC steps under hparam variants that the default-path fixture
(make_golden.py) does not reach:

  cem_base_controller.py:137-147  warm-up branches (zeros / hard-coded start action / sampler draw x context_action_weight)
  cem_base_controller.py:94-96,108-111  append_action
  cem_base_controller.py:89-91    selection_frac
  cem_base_controller.py:150-157  replan_interval
  gaussian_sampler.py:16-44 + pixel_cost_controller.py:161-165,199-204  reuse_mean / reduce_std_dev / predictor_propagation

    python tests/golden/make_act_variants_golden.py
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

H, W = 24, 32
AG = {"adim": 4, "sdim": 4, "image_height": H, "image_width": W, "gpu_id": 0}
BASE = {"rejection_sampling": False, "verbose": False, "num_samples": 18, "minimum_selection": 4}
VARIANTS = {
    "warm_sampler": dict(zeros_for_start_frames=False, start_planning=2),
    "warm_hard": dict(zeros_for_start_frames=False, hard_coded_start_action=[0.1, -0.2, 0.3, 0.0], start_planning=2),
    "append": dict(append_action=[0.7]),
    "selfrac": dict(selection_frac=0.5),
    "replan3": dict(replan_interval=3),
    "reuse": dict(reuse_mean=True, reduce_std_dev=0.5, replan_interval=3, predictor_propagation=True),
    # the `sampler` plugin point (cem_base_controller.py:52,66-76,82): class names are resolved per side by `samplers`
    "corr": dict(sampler="CorrelatedNoiseSampler"),
    "folding": dict(sampler="FoldingCEMSampler"),
    "autograsp": dict(sampler="AutograspSampler", iterations=1),      # the reference's refit path raises (autograsp_sampler.py:26)
    # two designated pixels: per-task scores averaged (pixel_cost_controller.py:148-153).  `only_take_first_view=True` with more
    # than one task cannot be recorded: the reference's logging loop indexes the sliced score matrix out of bounds (:158)
    "ndesig2": dict(designated_pixel_count=2),
}
STEPS = 6


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def inputs():
    rng = np.random.RandomState(77)
    return (rng.randint(0, 256, size=(STEPS, 1, H, W, 3)).astype(np.uint8), rng.uniform(-0.5, 0.5, size=(STEPS, 4)),
            np.array([[6, 8]]), np.array([[15, 20]]))


def ag_for(name):
    return dict(AG, adim=5) if name in ("append", "autograsp") else dict(AG)   # gripper value appended / drawn by rule


def drive(ctrl_cls, predictor_cls, get_policy_args, name, over, record, samplers=None):
    images, state, desig, goal = inputs()
    over = dict(over)
    if over.get("designated_pixel_count") == 2:
        desig, goal = np.array([[6, 8], [17, 5]]), np.array([[15, 20], [3, 27]])
    if isinstance(over.get("sampler"), str):
        over["sampler"] = getattr(samplers, over["sampler"])
        over.pop("rejection_sampling", None)
    base = dict(BASE)
    if "sampler" in over and "rejection_sampling" not in over["sampler"].get_default_hparams():
        base.pop("rejection_sampling")                              # not a hyper-parameter of this sampler: overriding it raises
    pol = quiet(ctrl_cls, ag_for(name), dict(base, predictor_class=predictor_cls, **over), 0, 1)
    if name == "append":
        pol._adim = 4                                               # the sampler draws the 4 arm dimensions (run.py passes env dims)
    quiet(pol.reset)
    np.random.seed(101)
    for t in range(STEPS):
        obs = {"images": images[:t + 1], "state": state[:t + 1]}
        kw = get_policy_args(pol, obs, t, 0, {"desig_pix": desig, "goal_pix": goal})
        out = quiet(pol.act, **kw)
        record("%s_t%d_action" % (name, t), np.array(out["actions"]))
        record("%s_t%d_ncalls" % (name, t), np.array(len(pol.predictor.calls)))
        if pol._best_indices is not None:
            record("%s_t%d_best_indices" % (name, t), np.array(pol._best_indices))
            last = max(int(k[len("scores_itr"):]) for k in out["plan_stat"] if k.startswith("scores_itr"))
            record("%s_t%d_scores_last" % (name, t), np.array(out["plan_stat"]["scores_itr%d" % last]))
    return pol


def main():
    ref_shim.install()
    from visual_mpc.policy.cem_controllers import PixelCostController
    from visual_mpc.policy.policy import get_policy_args
    import types
    from visual_mpc.policy.cem_controllers.samplers.autograsp_sampler import AutograspSampler
    from visual_mpc.policy.cem_controllers.samplers.correlated_noise import CorrelatedNoiseSampler
    from visual_mpc.policy.cem_controllers.samplers.folding_sampler import FoldingCEMSampler
    ref_samplers = types.SimpleNamespace(AutograspSampler=AutograspSampler, CorrelatedNoiseSampler=CorrelatedNoiseSampler,
                                         FoldingCEMSampler=FoldingCEMSampler)
    G = {}
    for name, over in VARIANTS.items():
        drive(PixelCostController, BlobPredictor, get_policy_args, name, over, G.__setitem__, ref_samplers)
    out = os.path.join(HERE, "ref_act_variants_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, len(G), "arrays", os.path.getsize(out) // 1024, "KB")


if __name__ == "__main__":
    main()

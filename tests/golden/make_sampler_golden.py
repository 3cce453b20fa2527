"""Generates tests/golden/ref_samplers_golden.npz by running the UNMODIFIED reference Autograsp and Folding samplers
(under ref_shim) on seeded inputs.  Run in the authoring container only:

    python tests/golden/make_sampler_golden.py

Reference functions exercised (paths relative to /root/reference/visual_mpc/policy/cem_controllers/samplers):
  autograsp_sampler.py:21-23,40-58   (sample_initial_actions, _sample_gripper; with and without reopen / deviation noise)
  autograsp_sampler.py:25-38         (sample_next_actions: raises TypeError in the reference — recorded as a flag; the
                                      no_refit=False gripper resampling law is pinned by calling its body's pieces)
  folding_sampler.py:18-118          (sample_initial_actions, sample_next_actions, _sample)
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


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    ref_shim.install()
    from visual_mpc.policy.cem_controllers.samplers.autograsp_sampler import AutograspSampler
    from visual_mpc.policy.cem_controllers.samplers.folding_sampler import FoldingCEMSampler
    HP = ref_shim._HParams
    G = {}
    rng = np.random.RandomState(99)

    # ---------------- autograsp ---------------------------------------------------------------------------------
    state = np.array([0.1, -0.2, 0.22, 0.0, -1.0])
    G["ag_state"] = state
    for tag, over in (("default", {}), ("reopen_dev", {"reopen": True, "deviation_prob": 0.3}),
                      ("scaled", {"action_norm_factor": 2.5, "z_thresh": 0.05, "deviation_prob": 0.1})):
        d = AutograspSampler.get_default_hparams()
        d.update(rejection_sampling=False, **over)
        smp = AutograspSampler(HP(**d), 5, 5)
        np.random.seed(21)
        G["ag_init_%s" % tag] = quiet(smp.sample_initial_actions, 1, 12, state)
    d = AutograspSampler.get_default_hparams()
    d.update(rejection_sampling=False)
    smp = AutograspSampler(HP(**d), 5, 5)
    np.random.seed(21)
    first = quiet(smp.sample_initial_actions, 1, 12, state)
    try:
        quiet(smp.sample_next_actions, 12, first[:6], np.arange(6.0))
        G["ag_next_raises"] = np.array(0)
    except TypeError:
        G["ag_next_raises"] = np.array(1)

    # ---------------- folding -----------------------------------------------------------------------------------
    fhp = HP(**FoldingCEMSampler.get_default_hparams())
    fs = FoldingCEMSampler(fhp, 4, 4)
    fstate = np.array([0.4, 0.6, 0.1, 0.0])
    G["fold_state"] = fstate
    np.random.seed(31)
    G["fold_init_seed31"] = quiet(fs.sample_initial_actions, 1, 24, fstate)
    elites = rng.randn(8, 15, 4) * 0.05
    elites = np.repeat(elites[:, ::3], 3, axis=1)
    G["fold_elites"] = elites
    np.random.seed(32)
    G["fold_next_seed32"] = quiet(fs.sample_next_actions, 24, elites, np.arange(8.0))
    fhp2 = HP(**dict(FoldingCEMSampler.get_default_hparams(), split_frac=0.9, nactions=6, repeat=2))
    fs2 = FoldingCEMSampler(fhp2, 4, 4)
    np.random.seed(33)
    G["fold_init6_seed33"] = quiet(fs2.sample_initial_actions, 0, 12, fstate)

    out = os.path.join(HERE, "ref_samplers_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, len(G), "arrays")


if __name__ == "__main__":
    main()

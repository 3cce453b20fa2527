"""Host-side mirror of the reference plugin surface (HParams, Policy glue, samplers, CEM controller)
against outputs of the UNMODIFIED reference code (tests/golden/ref_cem_golden.npz)."""
import contextlib
import io

import numpy as np
import pytest

from fake_predictor import BlobPredictor
from oracle.cem import OracleBackend
from visual_foresight_b200 import samplers as S
from visual_foresight_b200.cem_controller import PixelCostController
from visual_foresight_b200.hparams import HParams
from visual_foresight_b200.policy import NullPolicy, get_policy_args


def ghp(**over):
    d = S.GaussianCEMSampler.get_default_hparams()
    d.update(replan_interval=0)
    d.update(over)
    return HParams(**d)


def test_hparams_semantics():
    hp = HParams(a=1, b=None)
    with pytest.raises(ValueError):
        hp.add_hparam("a", 2)
    with pytest.raises(KeyError):
        hp.set_hparam("zzz", 2)
    hp.set_hparam("a", 5)
    hp.b = [1, 2]
    assert hp.a == 5 and hp.get("b") == [1, 2] and "a" in hp and "q" not in hp and hp.values() == {"a": 5, "b": [1, 2]}
    with pytest.raises(AttributeError):
        hp.nope


def test_initial_covariance(golden):
    for adim in (2, 3, 4, 5):
        for t in (0, 3):
            np.testing.assert_array_equal(S.initial_covariance(ghp(reduce_std_dev=0.5), adim, t), golden["sigma0_adim%d_t%d" % (adim, t)])
    np.testing.assert_array_equal(S.initial_covariance(ghp(action_order=["x", "y", "z", "theta", "grasp"]), 5, 0), golden["sigma0_order"])


def test_clip_band_discretize(golden):
    np.testing.assert_array_equal(S.clip_actions(golden["trunc_in3"].copy(), ghp()), golden["trunc_out3"])
    np.testing.assert_array_equal(S.clip_actions(golden["trunc_in2"].copy(), ghp()), golden["trunc_out2"])
    np.testing.assert_array_equal(S.clip_actions(golden["trunc_in3"].copy(), ghp(action_order=["z", "x", "theta", "y"])), golden["trunc_out3_order"])
    np.testing.assert_array_equal(S.band_mask_covariance(golden["blockdiag_in"], 5, 4), golden["blockdiag_out"])
    np.testing.assert_array_equal(S.discretize_actions(golden["discretize_in"].copy(), [2, 3]), golden["discretize_out"])


def test_shifted_covariance_runs():
    cov = np.cov(np.random.RandomState(0).randn(40, 20), rowvar=False)
    out = S.shifted_covariance(cov, 4, ghp(replan_interval=3, reuse_cov=0.25))
    init = S.initial_covariance(ghp(), 4)
    np.testing.assert_allclose(out[:-4, :-4], cov[4:, 4:] + 0.25 * init[:-4, :-4])
    np.testing.assert_allclose(out[-4:, -4:], init[:4, :4])
    assert np.all(out[-4:, :-4] == 0)


def test_gaussian_sampler_streams(golden):
    smp = S.GaussianCEMSampler(ghp(rejection_sampling=False), 4, 4)
    np.random.seed(7)
    np.testing.assert_array_equal(smp.sample_initial_actions(1, 16, None), golden["gauss_init_actions_seed7"])
    np.random.seed(11)
    nxt = smp.sample_next_actions(16, golden["gauss_fit_elites"], np.arange(10.0))
    np.testing.assert_allclose(smp._mean, golden["gauss_fit_mean"], atol=1e-15)
    np.testing.assert_allclose(smp._sigma, golden["gauss_fit_sigma"], rtol=1e-12, atol=1e-16)
    np.testing.assert_allclose(nxt, golden["gauss_next_actions_seed11"], rtol=1e-9, atol=1e-12)
    smp2 = S.GaussianCEMSampler(ghp(rejection_sampling=False, cov_blockdiag=True, smooth_cov=True), 4, 4)
    np.random.seed(7)
    smp2.sample_initial_actions(1, 16, None)
    smp2._fit(golden["gauss_fit_elites"])
    np.testing.assert_allclose(smp2._sigma, golden["gauss_fit_sigma_blockdiag_smooth"], rtol=1e-12, atol=1e-16)


def test_gaussian_reuse_mean(golden):
    smp = S.GaussianCEMSampler(ghp(rejection_sampling=False, reuse_mean=True), 4, 4)
    smp.log_best_action(np.zeros(4), golden["gauss_reuse_plan"])
    np.random.seed(3)
    a = smp.sample_initial_actions(4, 16, None)
    np.testing.assert_array_equal(smp._mean, golden["gauss_reuse_mean"])
    np.testing.assert_array_equal(a, golden["gauss_reuse_actions_seed3"])
    assert a.shape[0] == 16      # first call: no previous mean -> cold start (reference gaussian_sampler.py:23)
    b = smp.sample_initial_actions(5, 16, None)      # second call warm-starts from the logged plan
    assert b.shape[0] == 8       # reuse_factor 0.5
    plan = golden["gauss_reuse_plan"][0]
    want = np.zeros((5, 4))
    want[:5] = np.concatenate([plan, np.zeros((2, 4))])[::3][:5]
    np.testing.assert_array_equal(smp._mean, want.reshape(-1))


def test_rejection_sampler_bounds():
    smp = S.GaussianCEMSampler(ghp(), 4, 4)
    np.random.seed(0)
    a = smp.sample_initial_actions(0, 6, None)
    assert a.shape == (6, 15, 4)
    assert np.abs(a[:, :, :2]).max() <= 1.5 * 0.05 and np.abs(a[:, :, 2]).max() <= 1.5 * 0.15


def test_correlated_noise_sampler(golden):
    hp = HParams(**S.CorrelatedNoiseSampler.get_default_hparams())
    cs = S.CorrelatedNoiseSampler(hp, 4, 4)
    np.random.seed(5)
    np.testing.assert_array_equal(cs.sample_initial_actions(1, 12, None), golden["corr_init_seed5"])
    np.random.seed(6)
    np.testing.assert_allclose(cs.sample_next_actions(12, golden["corr_best"], golden["corr_scores"]),
                               golden["corr_next_seed6"], rtol=1e-12, atol=1e-15)


def test_autograsp_sampler_matches_reference(sampler_golden):
    """AutograspSampler (reference samplers/autograsp_sampler.py) draws the same np.random stream and applies the same
    gripper rule as the unmodified reference: initial batches bit-identical for the default, reopen + deviation-noise and
    scaled-threshold settings (fixtures: tests/golden/make_sampler_golden.py)."""
    g = sampler_golden
    assert int(g["ag_next_raises"]) == 1          # the reference's sample_next_actions is broken (missing argument)
    for tag, over in (("default", {}), ("reopen_dev", {"reopen": True, "deviation_prob": 0.3}),
                      ("scaled", {"action_norm_factor": 2.5, "z_thresh": 0.05, "deviation_prob": 0.1})):
        d = S.AutograspSampler.get_default_hparams()
        d.update(rejection_sampling=False, **over)
        smp = S.AutograspSampler(HParams(**d), 5, 5)
        np.random.seed(21)
        a = smp.sample_initial_actions(1, 12, g["ag_state"])
        np.testing.assert_array_equal(a, g["ag_init_" + tag])
        assert set(np.unique(a[:, :, -1])) <= {-1.0, 1.0}
    # the repaired sample_next_actions: arm dimensions refit like the Gaussian sampler, gripper by rule or by elite frequency
    d = S.AutograspSampler.get_default_hparams()
    d.update(rejection_sampling=False)
    smp = S.AutograspSampler(HParams(**d), 5, 5)
    np.random.seed(21)
    first = smp.sample_initial_actions(1, 12, g["ag_state"])
    nxt = smp.sample_next_actions(12, first[:6], np.arange(6.0))
    assert nxt.shape == (12, 15, 5)
    z = np.cumsum(nxt[:, :, 2], axis=1) + g["ag_state"][2] < 0.15
    for row in range(12):                                            # closes at the first crossing and stays closed
        want = np.zeros(15, bool)
        if z[row].any():
            want[int(np.argmax(z[row])):] = True
        np.testing.assert_array_equal(nxt[row, :, 4] == 1, want)
    d.update(no_refit=False)
    smp = S.AutograspSampler(HParams(**d), 5, 5)
    np.random.seed(21)
    first = smp.sample_initial_actions(1, 12, g["ag_state"])
    elites = first[:6].copy()
    elites[:, :5, 4], elites[:, 5:, 4] = -1, 1                       # elites all open for 5 steps, then all closed
    nxt = smp.sample_next_actions(12, elites, np.arange(6.0))
    assert np.all(nxt[:, :5, 4] == -1) and np.all(nxt[:, 5:, 4] == 1)


def test_folding_sampler_matches_reference(sampler_golden):
    """FoldingCEMSampler (reference samplers/folding_sampler.py) reproduces the unmodified reference's action tensors under
    the same np.random seed: initial proposal, refit proposal, and a 6-step / repeat-2 / split 0.9 configuration."""
    g = sampler_golden
    fs = S.FoldingCEMSampler(HParams(**S.FoldingCEMSampler.get_default_hparams()), 4, 4)
    np.random.seed(31)
    np.testing.assert_array_equal(fs.sample_initial_actions(1, 24, g["fold_state"]), g["fold_init_seed31"])
    np.random.seed(32)
    np.testing.assert_allclose(fs.sample_next_actions(24, g["fold_elites"], np.arange(8.0)), g["fold_next_seed32"],
                               rtol=1e-10, atol=1e-13)
    fs2 = S.FoldingCEMSampler(HParams(**dict(S.FoldingCEMSampler.get_default_hparams(), split_frac=0.9, nactions=6, repeat=2)), 4, 4)
    np.random.seed(33)
    a = fs2.sample_initial_actions(0, 12, g["fold_state"])
    np.testing.assert_array_equal(a, g["fold_init6_seed33"])
    lim = np.array([1. / 5, 1. / 5, 1. / 3])
    assert np.all(np.abs(a[:, :, :3]) <= lim + 1e-15)
    with pytest.raises(AssertionError):
        fs.sample_initial_actions(0, 10, g["fold_state"])            # sample count must split three ways
    with pytest.raises(AssertionError):
        S.FoldingCEMSampler(HParams(**S.FoldingCEMSampler.get_default_hparams()), 5, 4)


def test_pred_util_matches_reference(golden):
    """get_context / rollout_predictions (reference video_prediction/pred_util.py:4-48): context slice with state_append,
    chunking with a zero-padded last chunk, outputs cut back and returned per chunk."""
    from visual_foresight_b200 import pred_util as PU
    g = golden
    lf, ls = PU.get_context(2, 3, g["ctx_states"], g["ctx_images"], HParams(state_append=[0.1, 0.2]))
    np.testing.assert_array_equal(lf, g["ctx_last_frames"])
    np.testing.assert_array_equal(ls, g["ctx_last_states"])
    assert lf.dtype == np.float32 and lf.shape == (1, 2, 2, 8, 8, 3)
    lf2, ls2 = PU.get_context(2, 3, g["ctx_states"], g["ctx_images"])
    assert ls2.shape == (1, 2, 3)
    seen = []

    def pf(input_images=None, input_state=None, input_actions=None, input_one_hot_images=None):
        seen.append(input_actions.copy())
        s = input_actions.sum(axis=(1, 2))
        return s[:, None] * np.ones((1, 2)), None, s[:, None] * 2.0
    gi, gd, gs = PU.rollout_predictions(pf, 3, g["rollout_actions"], lf, ls)
    assert len(seen) == int(g["rollout_ncalls"]) == 3 and [a.shape[0] for a in gi] == [3, 3, 1] and gd == [None] * 3
    np.testing.assert_array_equal(np.concatenate(gi, 0), g["rollout_gen_images"])
    np.testing.assert_array_equal(np.concatenate(gs, 0), g["rollout_gen_states"])
    np.testing.assert_array_equal(seen[-1], g["rollout_seen_last"])          # zero-padded to the chunk size


def test_predictor_class_chunks_oversized_batches():
    """B200VPredEvaluation.__call__ with more action sequences than the handle's capacity: capacity-sized engine calls, ragged
    last chunk, outputs concatenated in order (the role of rollout_predictions, pred_util.py:21-48).  Stub backend: no GPU."""
    from visual_foresight_b200.predictor import B200VPredEvaluation
    pred = B200VPredEvaluation("", {"run_batch_size": 4, "model_spec": {"height": 8, "width": 8, "seq_len": 4}})
    calls = []

    class Stub:
        def predict(self, context, actions):
            calls.append(actions.shape[0])
            tag = actions[:, 0, 0].astype(np.float32)
            return tag[:, None] * np.ones((1, 2), np.float32), tag[:, None] * 2 * np.ones((1, 3), np.float32), None
    pred.backend = Stub()
    acts = np.arange(10, dtype=np.float64)[:, None, None] * np.ones((1, 3, 4))
    out = pred({}, {"actions": acts})
    assert calls == [4, 4, 2]
    np.testing.assert_array_equal(out["predicted_frames"][:, 0], np.arange(10, dtype=np.float32))
    np.testing.assert_array_equal(out["predicted_pixel_distributions"][:, 0], 2 * np.arange(10, dtype=np.float32))
    calls.clear()
    pred({}, {"actions": acts[:3]})
    assert calls == [3]


def test_policy_arg_resolution():
    pol = NullPolicy({"adim": 4}, {})
    assert get_policy_args(pol, {}, 0, 0) == {}
    assert pol.act()["actions"].shape == (4,)

    class Wants(NullPolicy):
        def act(self, t, i_tr, images, goal_pix=None, obs=None, extra=7):
            return {}
    kw = get_policy_args(Wants({"adim": 2}, {}), {"images": "IM"}, 3, 9, {"goal_pix": "GP"})
    assert kw == {"t": 3, "i_tr": 9, "images": "IM", "goal_pix": "GP", "obs": {"images": "IM"}, "extra": 7}


def test_required_param_raises():
    class Needs(NullPolicy):
        def act(self, t, unknown_thing):
            return {}
    with pytest.raises(ValueError):
        get_policy_args(Needs({"adim": 2}, {}), {}, 0, 0, {})


class _Injected(BlobPredictor):
    """BlobPredictor plus the rollout-evaluator backend attribute the controller looks for."""
    def __init__(self, *a, **k):
        BlobPredictor.__init__(self, *a, **k)
        self.backend = OracleBackend(self)


AG = {"adim": 4, "sdim": 4, "image_height": 48, "image_width": 64, "gpu_id": 0}
PP = {"predictor_class": _Injected, "rejection_sampling": False, "verbose": False, "num_samples": 24,
      "minimum_selection": 5, "device_cem": False}


def test_override_semantics(golden):
    for bad, want in zip(({"iterations": 3}, {"not_a_param": 1}), golden["override_errors"]):
        with pytest.raises({"ValueError": ValueError, "AttributeError": AttributeError}[str(want)]):
            PixelCostController(AG, dict(PP, **bad), 0, 1)


def test_full_act_matches_reference(golden):
    """Three MPC steps through get_policy_args -> act() with the seeded global RNG: actions sampled,
    context handed to the predictor, scores, elite indices and the returned action all equal the
    reference controller's."""
    pol = PixelCostController(AG, dict(PP), 0, 1)
    pol.reset()
    assert pol._hp.start_planning == int(golden["act_start_planning"])
    images, state = golden["act_images"], golden["act_state"]
    np.random.seed(42)
    for t in range(3):
        obs = {"images": images[:t + 1], "state": state[:t + 1]}
        kw = get_policy_args(pol, obs, t, 0, {"desig_pix": golden["act_desig"], "goal_pix": golden["act_goal"]})
        assert sorted(kw) == sorted(["t", "i_tr", "desig_pix", "goal_pix", "images", "state", "verbose_worker"])
        out = pol.act(**kw)
        np.testing.assert_allclose(out["actions"], golden["act_t%d_action" % t], rtol=1e-9, atol=1e-12)
        if t >= 1:
            call = pol.predictor.calls[-1]
            np.testing.assert_allclose(call["actions"], golden["act_t%d_last_actions" % t], rtol=1e-9, atol=1e-12)
            np.testing.assert_array_equal(call["context_pixel_distributions"], golden["act_t%d_ctx_distrib" % t])
            np.testing.assert_allclose(call["context_actions"], golden["act_t%d_ctx_actions" % t], rtol=1e-9, atol=1e-12)
            for i in range(3):
                np.testing.assert_allclose(out["plan_stat"]["scores_itr%d" % i], golden["act_t%d_scores_itr%d" % (t, i)], rtol=1e-9)
            np.testing.assert_array_equal(pol._best_indices, golden["act_t%d_best_indices" % t])
            np.testing.assert_allclose(pol._best_actions, golden["act_t%d_best_actions" % t], rtol=1e-9, atol=1e-12)


@pytest.mark.filterwarnings("ignore:covariance is not symmetric positive-semidefinite")
@pytest.mark.parametrize("name", ["warm_sampler", "warm_hard", "append", "selfrac", "replan3", "reuse", "corr", "folding", "autograsp",
                                  "ndesig2"])
def test_act_variants_match_reference(name):
    """Six MPC steps of act() under hparam variants the default-path fixture does not reach (warm-up branches, append_action,
    selection_frac, replan_interval, reuse_mean + reduce_std_dev + predictor_propagation, and the CorrelatedNoise / Folding /
    Autograsp samplers through the `sampler` plugin point): returned actions, number of
    predictor calls, elite indices and last-iteration scores equal the unmodified reference's, step by step
    (fixtures: tests/golden/make_act_variants_golden.py)."""
    import os
    import make_act_variants_golden as MG
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_act_variants_golden.npz"))
    got = {}
    over = dict(MG.VARIANTS[name], device_cem=False)
    MG.drive(PixelCostController, _Injected, get_policy_args, name, over, got.__setitem__, S)
    keys = [k for k in g.files if k.startswith(name + "_")]
    assert keys and sorted(keys) == sorted(got)
    for k in keys:
        if k.endswith("best_indices") or k.endswith("ncalls"):
            np.testing.assert_array_equal(got[k], g[k], err_msg=k)
        else:
            np.testing.assert_allclose(got[k], g[k], rtol=1e-9, atol=1e-12, err_msg=k)

"""-m gpu: the CUDA path through the C-ABI (ctypes -> libvfengine.so) against the oracle on seeded
inputs and against the committed golden fixtures.

Tolerances (stated here as the contract):
  * integer / index work (top-K elite sets, one-hot, shard invariance)      : bit-exact
  * predicted frames vs oracle (spec P, fp32), all precisions marked fp32-grade: max-abs <= 1e-4
  * scores                                                                     : rel 1e-5
Parity is vs the build's oracle (spec P); TF1 parity is unpinned."""
import os

import numpy as np
import pytest

import helpers as Hh
from oracle import cem as OC
from visual_foresight_b200 import spec as S

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_TOL = 1e-4


@pytest.fixture(scope="module")
def eng_small():
    from visual_foresight_b200.engine import Engine
    sp = S.spec_64(height=32, width=32, seq_len=5)
    e = Engine(sp, 8)
    yield e
    e.close()


# ---- kernels ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 16, 16, 8, 64, 5), (3, 8, 8, 64, 128, 5), (2, 32, 32, 6, 32, 5),
                                               (2, 12, 16, 40, 32, 3), (1, 6, 8, 20, 64, 3), (2, 16, 16, 32, 3, 3),
                                               (2, 16, 16, 53, 7, 3)])
def test_conv_simt_vs_torch(eng_small, B, H, W, Cin, Cout, k):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(B * 1000 + Cin)
    x = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    y = eng_small.debug_conv2d(x, w, b)
    ref = F.conv2d(torch.from_numpy(x).double().permute(0, 3, 1, 2), torch.from_numpy(w).double().permute(3, 2, 0, 1),
                   torch.from_numpy(b).double(), padding=k // 2).permute(0, 2, 3, 1).numpy()
    assert np.abs(y - ref).max() < 2e-5


def test_topk_is_stable_argsort(eng_small):
    rng = np.random.default_rng(0)
    for n, k in [(5, 5), (200, 10), (600, 30), (4096, 205), (1000, 667)]:
        s = rng.standard_normal(n)
        s[rng.integers(0, n, n // 4)] = s[rng.integers(0, n, n // 4)]      # ties
        if n > 100:
            s[7] = np.nan
        np.testing.assert_array_equal(eng_small.topk(s, k), np.argsort(s, kind="stable")[:k])
    np.testing.assert_array_equal(eng_small.topk(np.zeros(37), 37), np.arange(37))


def test_topk_on_reference_scores(eng_small, golden):
    for t in (1, 2):
        np.testing.assert_array_equal(eng_small.topk(golden["act_t%d_scores_itr2" % t], 5), golden["act_t%d_best_indices" % t])


def test_refit_vs_reference(eng_small, golden):
    mean, cov, fac = eng_small.refit(golden["gauss_fit_elites"], 5, 3)
    np.testing.assert_allclose(mean, golden["gauss_fit_mean"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(cov, golden["gauss_fit_sigma"], rtol=1e-11, atol=1e-16)
    np.testing.assert_allclose(fac @ fac.T, golden["gauss_fit_sigma"], rtol=1e-10, atol=1e-15)


def test_pixel_cost_vs_reference_golden(golden):
    """device cost kernel on the reference's own input -> the reference's own scores."""
    from visual_foresight_b200.engine import Engine
    gen = golden["cost_gen_distrib"]
    e = Engine(S.spec_64(height=32, width=32, ndesig=2, seq_len=15), gen.shape[0])
    for fw, key in ((10.0, "cost_scores_2desig"), (3.0, "cost_scores_2desig_fw3")):
        sc = e.score_external(gen, golden["cost_goal_pix"], finalweight=fw)
        np.testing.assert_allclose(sc, golden[key], rtol=1e-5)
    # ragged / edge: a single plane, goal outside the image, one-hot distribution
    one = np.zeros((1, 1, 1, 32, 32, 2), np.float32)
    one[0, 0, 0, 3, 4, 0] = 1
    one[0, 0, 0, 23, 31, 1] = 2
    sc = e.score_external(one, np.array([[[40.0, -3.0], [0, 0]]]), finalweight=10.0)
    want = 0.5 * (np.hypot(40 - 3, -3 - 4) + np.hypot(23, 31))
    np.testing.assert_allclose(sc, [want], rtol=1e-6)
    e.close()


# ---- predictor ----------------------------------------------------------------------------------------
def _engine_rollout(sp, weights, inp, acts, precision="fp32_simt"):  # fp32 FFMA checker path unless a test asks for the tensor-core path
    from visual_foresight_b200.engine import Engine
    e = Engine(sp, acts.shape[0], precision=precision)
    e.load_weights(weights)
    e.set_context(inp["frames"], inp["states"] if sp.sdim else None, inp["ctx_actions"])
    e.set_desig(inp["desig"])
    out = e.predict(acts)
    return e, out


@pytest.mark.parametrize("case", ["a", "b"])
def test_rollout_vs_golden_fixture(case):
    from make_predictor_golden import CASES
    c = dict(CASES[case])
    M = c.pop("M")
    sp = S.spec_64(**c)
    w = Hh.make_weights(sp, seed=3)
    inp = Hh.synth_inputs(sp, seed=5)
    acts = Hh.gaussian_actions(sp, M, sp.seq_len - sp.context_frames + 2, seed=7)
    e, (gi, gd, gs) = _engine_rollout(sp, w, inp, acts)
    g = np.load(os.path.join(GOLD, "oracle_predictor_golden.npz"))
    assert np.abs(gi - g[case + "_frames"]).max() <= FRAME_TOL
    assert np.abs(gd - g[case + "_distrib"]).max() <= 1e-5
    np.testing.assert_allclose(gs, g[case + "_states"], atol=1e-5)
    np.testing.assert_allclose(gd.sum(axis=(3, 4)), 1.0, atol=1e-5)        # renormalised distributions
    e.close()


def test_rollout_intermediates_vs_oracle():
    """layer-by-layer parity at the last cell step (localises a mismatch to a kernel)."""
    import torch
    from oracle.predictor import OraclePredictor
    sp = S.spec_64(height=32, width=32, seq_len=4)
    w = Hh.make_weights(sp, seed=1)
    inp = Hh.synth_inputs(sp, seed=2)
    acts = Hh.gaussian_actions(sp, 2, 3, seed=3)
    e, (gi, gd, gs) = _engine_rollout(sp, w, inp, acts)
    dbg = {sp.seq_len - 2: {}}
    o = OraclePredictor(sp, w[0], torch.float32)
    onehot = OC.switch_on_pix(inp["desig"], 2, 1, 32, 32, 1)
    oi, od, os_ = o.rollout(inp["frames"][:, 0].astype(np.float32) / 255, inp["states"], onehot[:, 0],
                            Hh.step_actions(sp, inp["ctx_actions"], acts), debug_steps=dbg)
    d = dbg[sp.seq_len - 2]
    for name in ["enc0.conv", "enc0.lstm.h", "enc0.lstm.c", "enc1.conv", "enc1.lstm.h", "enc2.conv", "enc2.lstm.h",
                 "dec0.conv", "dec0.lstm.h", "dec1.conv", "dec1.lstm.h", "dec2.conv", "scratch", "mask_logits"]:
        ref = d[name].permute(0, 2, 3, 1).contiguous().numpy()
        got = e.debug_fetch(name).reshape(ref.shape)
        assert np.abs(got - ref).max() <= 2e-4, name
    kref = d["cdna.kernels"].numpy().reshape(2, -1)
    assert np.abs(e.debug_fetch("cdna.kernels").reshape(2, -1) - kref).max() <= 1e-5
    assert np.abs(gi - oi[:, :, None]).max() <= FRAME_TOL
    e.close()


def test_rollout_c1_config_vs_oracle():
    """BASELINE config c1 (M=8, S=5, 48x64) — frames within 1e-4 of the oracle, fp64 oracle as arbiter."""
    import torch
    sp = S.spec_64(height=48, width=64, seq_len=5)
    w = Hh.make_weights(sp, seed=0)
    inp = Hh.synth_inputs(sp, seed=0)
    acts = Hh.gaussian_actions(sp, 8, 15, seed=0)
    e, (gi, gd, gs) = _engine_rollout(sp, w, inp, acts)
    oi, od, os_ = Hh.oracle_rollout(sp, w, inp, acts)
    o64 = Hh.oracle_rollout(sp, w, inp, acts, dtype=torch.float64)[0]
    assert np.abs(gi - oi).max() <= FRAME_TOL
    assert np.abs(gi - o64).max() <= FRAME_TOL
    assert np.abs(gd - od).max() <= 1e-5
    sc = e.score(inp["goal"], M=8)
    np.testing.assert_allclose(sc, OC.eval_pixel_cost(od, inp["goal"]), rtol=1e-5)
    e.close()


def test_rollout_128px_spec_tensor_core_vs_oracle():
    """BASELINE c5 geometry family: 128x128 frames, 4-level encoder whose first layer is not recurrent, 256-channel
    conv-LSTMs (gate convs with 1024 output channels) on the tensor-core path — frames within 1e-4 of the fp32 oracle."""
    sp = S.spec_128(seq_len=4)
    w = Hh.make_weights(sp, seed=2)
    inp = Hh.synth_inputs(sp, seed=4)
    acts = Hh.gaussian_actions(sp, 2, 15, seed=6)
    e, (gi, gd, gs) = _engine_rollout(sp, w, inp, acts, precision="f16x3")
    oi, od, os_ = Hh.oracle_rollout(sp, w, inp, acts)
    assert np.abs(gi - oi).max() <= FRAME_TOL
    assert np.abs(gd - od).max() <= 1e-5
    np.testing.assert_allclose(gs, os_, atol=1e-5)
    e.close()


@pytest.mark.parametrize("precision,rnn_z", [("fp32_simt", False), ("f16x3", False), ("f16x3", True)])
def test_rollout_with_latents_vs_oracle(precision, rnn_z):
    """stochastic predictor input (nz > 0): the per-step latent z is tiled into every conv like the action/state vector
    (spec P1); identical z on both sides -> frames within 1e-4."""
    import torch
    from oracle.predictor import OracleMultiViewPredictor
    from visual_foresight_b200.engine import Engine
    sp = S.spec_64(height=32, width=32, seq_len=5, nz=8, rnn_z=rnn_z)   # rnn_z: the latent passes through a dense LSTM first
    w = Hh.make_weights(sp, seed=11)
    inp = Hh.synth_inputs(sp, seed=12)
    M = 3
    acts = Hh.gaussian_actions(sp, M, 15, seed=13)
    zs = np.random.default_rng(14).standard_normal((M, sp.seq_len - 1, sp.nz)).astype(np.float32)
    e = Engine(sp, M, precision=precision)
    e.load_weights(w)
    e.set_context(inp["frames"], inp["states"], inp["ctx_actions"])
    e.set_desig(inp["desig"])
    gi, gd, gs = e.predict(acts, zs=zs)
    onehot = OC.switch_on_pix(inp["desig"], sp.context_frames, sp.ncam, sp.height, sp.width, sp.ndesig)
    oi, od, os_ = OracleMultiViewPredictor(sp, w, torch.float32).rollout(
        inp["frames"].astype(np.float32) / 255.0, inp["states"], onehot, Hh.step_actions(sp, inp["ctx_actions"], acts), zs)
    assert np.abs(gi - oi).max() <= FRAME_TOL
    assert np.abs(gd - od).max() <= 1e-5
    z2 = zs.copy()
    z2[1] += 1.0                                             # the latent really is an input: changing it changes the frames
    gi2 = e.predict(acts, zs=z2)[0]
    assert np.abs(gi2[1] - gi[1]).max() > 1e-4 and np.abs(gi2[0] - gi[0]).max() == 0.0
    e.close()


@pytest.mark.parametrize("name,kw,M", [
    ("c3", dict(height=48, width=64, seq_len=13, ncam=2, ndesig=2, adim=4, sdim=5), 600),      # Sawyer two-view, full size
    ("c4-shard", dict(height=64, width=64, seq_len=15), 512),                                  # M=4096 / 8 GPUs
])
def test_full_size_configs_batch_invariance(name, kw, M):
    """BASELINE c3 / one c4 shard at FULL sample count on the tensor-core path: the oracle only rolls three of the samples
    (first, middle, last); every sample's distributions stay normalised, scores are finite, equal action rows give equal
    frames (batch-position invariance), and the elite set equals the stable argsort of the device scores."""
    from visual_foresight_b200.engine import Engine
    sp = S.spec_64(**kw)
    w = Hh.make_weights(sp, seed=21)
    inp = Hh.synth_inputs(sp, seed=22)
    acts = Hh.gaussian_actions(sp, M, 15, seed=23)
    acts[M - 2] = acts[1]                                    # duplicate action row at another batch position
    e = Engine(sp, M, precision="f16x3")
    e.load_weights(w)
    e.set_context(inp["frames"], inp["states"] if sp.sdim else None, inp["ctx_actions"])
    e.set_desig(inp["desig"])
    e.predict(acts, fetch=False)
    pick = [0, M // 2, M - 1, 1, M - 2]
    gi, gd = e.fetch(pick)
    oi, od, _ = Hh.oracle_rollout(sp, w, inp, acts[pick[:3]])
    assert np.abs(gi[:3] - oi).max() <= FRAME_TOL, name
    assert np.abs(gd[:3] - od).max() <= 1e-5
    np.testing.assert_array_equal(gi[3], gi[4])
    np.testing.assert_array_equal(gd[3], gd[4])
    sc = e.score(inp["goal"], M=M)
    assert sc.shape == (M,) and np.all(np.isfinite(sc))
    np.testing.assert_allclose(sc[pick[:3]], OC.eval_pixel_cost(od, inp["goal"]), rtol=1e-5)
    K = max(10, M // 20)
    np.testing.assert_array_equal(e.topk(sc, K), np.argsort(sc, kind="stable")[:K])
    e.close()


@pytest.mark.parametrize("precision", ["fp32_simt", "f16x3"])
def test_ragged_batches_and_call_order_errors(precision):
    """Ragged sample counts (M = 1, an odd M that leaves a partly filled tile group, M = capacity) give bit-identical
    per-sample results — a sample never depends on its batch — and the C-ABI reports misuse as errors (negative status +
    vf_last_error) instead of computing on stale state: the Python layer raises, like the reference's asserts."""
    from visual_foresight_b200.engine import Engine, EngineError
    sp = S.spec_64(height=32, width=32, seq_len=5)
    w = Hh.make_weights(sp, seed=41)
    inp = Hh.synth_inputs(sp, seed=42)
    acts = Hh.gaussian_actions(sp, 9, 15, seed=43)
    e = Engine(sp, 9, precision=precision)
    with pytest.raises(EngineError, match="weight"):
        e.set_context(inp["frames"], inp["states"], inp["ctx_actions"])
        e.set_desig(inp["desig"])
        e.predict(acts[:1])                                  # no weights loaded
    e.close()
    e = Engine(sp, 9, precision=precision)
    e.load_weights(w)
    with pytest.raises(EngineError, match="vf_set_context"):
        e.predict(acts[:1])                                  # rollout before the context
    e.set_context(inp["frames"], inp["states"], inp["ctx_actions"])
    with pytest.raises(EngineError, match="designated-pixel"):
        e.predict(acts[:1])                                  # context without a designated-pixel distribution
    e.set_desig(inp["desig"])
    with pytest.raises(EngineError, match="max_samples"):
        e.predict(np.concatenate([acts, acts]))              # beyond the handle's capacity
    with pytest.raises(EngineError, match="actions per sample"):
        e.predict(acts[:, :1])                               # too few actions for the horizon
    full = e.predict(acts)
    for m in (1, 7):
        part = e.predict(acts[:m])
        for a_, b_ in zip(part, full):
            np.testing.assert_array_equal(a_, b_[:m])
    sc = e.score(inp["goal"], M=7)
    assert sc.shape == (7,) and np.all(np.isfinite(sc))
    np.testing.assert_array_equal(e.topk(sc, 7), np.argsort(sc, kind="stable"))      # K == M
    e.close()


# ---- CEM ------------------------------------------------------------------------------------------------
def _plan_kwargs(sp, M, K, iters, seed=0):
    from visual_foresight_b200.samplers import GaussianCEMSampler, action_bounds, per_dim_variance
    from visual_foresight_b200.hparams import HParams
    hp = HParams(**GaussianCEMSampler.get_default_hparams())
    lo, hi = action_bounds(hp, sp.adim)
    return dict(num_samples=M, iterations=iters, num_elites=K, nactions=5, repeat=3,
                std=np.sqrt(per_dim_variance(hp, sp.adim)), clip=(lo, hi), mean0=None, reduce_std_scale=1.0,
                finalweight=10.0, task_weights=None, seed=seed, plan_index=0)


def test_cem_plan_vs_oracle_explicit_noise():
    """Whole perform_CEM on device with the SAME standard-normal noise as the oracle planner: sampled
    actions equal, scores within tolerance, elite index sets bit-exact, best actions equal."""
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_64(height=32, width=32, seq_len=6)
    w = Hh.make_weights(sp, seed=4)
    inp = Hh.synth_inputs(sp, seed=4)
    M, K, iters = 12, 4, 3
    kw = _plan_kwargs(sp, M, K, iters)
    noise = np.random.default_rng(9).standard_normal((iters, M, 20)).astype(np.float32)
    be = EngineBackend(sp, w, M, precision="fp32_simt")
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": OC.switch_on_pix(inp["desig"], 2, 1, 32, 32, 1)}
    res = be.plan(ctx, goal_pix=inp["goal"], noise=noise, **kw)

    def evaluate(actions):
        _, od, _ = Hh.oracle_rollout(sp, w, inp, actions.astype(np.float32))
        return OC.eval_pixel_cost(od, inp["goal"])
    best, idx, scores, all_actions = OC.cem_plan(evaluate, num_samples=M, iterations=iters, num_elites_k=K, nactions=5,
                                                 repeat=3, adim=4, std=kw["std"], noise=noise, clip=kw["clip"])
    np.testing.assert_allclose(res["scores"], scores, rtol=1e-5)
    np.testing.assert_array_equal(res["elite_idx"], idx)
    np.testing.assert_allclose(res["best_actions"], best, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(be.engine.cem_actions(), all_actions[-1], rtol=1e-12, atol=1e-15)
    be.engine.close()


def test_stochastic_plan_futures_vs_oracle():
    """Stochastic planning (BASELINE c5 family: nz = 8, K futures per action sequence): the device plan equals the oracle
    plan that rolls np.repeat(actions, K, 0) with the restated Philox latents and scores mean_k + lambda * var_k —
    per-sequence scores within tolerance, elite sets bit-exact, best actions equal."""
    import torch
    from oracle.predictor import OracleMultiViewPredictor
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_64(height=32, width=32, seq_len=6, nz=8)
    w = Hh.make_weights(sp, seed=31)
    inp = Hh.synth_inputs(sp, seed=32)
    M, K, iters, KF, LAM, SEED, PLAN = 6, 2, 2, 3, 0.5, 77, 4
    kw = _plan_kwargs(sp, M, K, iters, seed=SEED)
    kw["plan_index"] = PLAN
    noise = np.random.default_rng(33).standard_normal((iters, M, 20)).astype(np.float32)
    be = EngineBackend(sp, w, M * KF, precision="fp32_simt")
    onehot = OC.switch_on_pix(inp["desig"], 2, 1, 32, 32, 1)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": onehot}
    res = be.plan(ctx, goal_pix=inp["goal"], noise=noise, k_futures=KF, lambda_variance=LAM, **kw)
    oracle = OracleMultiViewPredictor(sp, w, torch.float32)
    it = [0]

    def evaluate(actions):
        rep = np.repeat(actions.astype(np.float32), KF, axis=0)
        zs = OC.philox_latents(SEED, PLAN, it[0], M * KF, sp.seq_len - 1, sp.nz)
        it[0] += 1
        _, od, _ = oracle.rollout(inp["frames"].astype(np.float32) / 255.0, inp["states"], onehot,
                                  Hh.step_actions(sp, inp["ctx_actions"], rep), zs)
        return OC.reduce_futures(OC.eval_pixel_cost(od, inp["goal"]), KF, LAM)
    best, idx, scores, all_actions = OC.cem_plan(evaluate, num_samples=M, iterations=iters, num_elites_k=K, nactions=5,
                                                 repeat=3, adim=4, std=kw["std"], noise=noise, clip=kw["clip"])
    np.testing.assert_allclose(res["scores"], scores, rtol=1e-5)
    np.testing.assert_array_equal(res["elite_idx"], idx)
    np.testing.assert_allclose(res["best_actions"], best, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(be.engine.cem_actions(), all_actions[-1], rtol=1e-12, atol=1e-15)
    # the futures of one action sequence really differ (the latent reaches the frames)
    gi, _ = be.engine.fetch([0, 1])
    assert np.abs(gi[0] - gi[1]).max() > 1e-5
    be.engine.close()


@pytest.mark.timeout(900)
def test_c5_shard_stochastic_128px_full_size():
    """BASELINE c5 at the shape ONE of its 8 GPUs sees: 128x128 frames, S = 15, nz = 8 with the recurrent latent, 25 action
    sequences x 10 futures = 250 rollouts on the tensor-core path.  The oracle rolls the first two futures of action
    sequence 0 with the restated Philox latents (frames within 1e-4); the per-sequence plan scores equal
    mean + lambda * var of the device's own per-rollout scores; elites = stable argsort."""
    import torch
    from oracle.predictor import OracleMultiViewPredictor
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_128(seq_len=15, nz=8, rnn_z=True)
    w = Hh.make_weights(sp, seed=51)
    inp = Hh.synth_inputs(sp, seed=52)
    M, KF, K, LAM, SEED, PLAN = 25, 10, 5, 0.25, 5, 2
    kw = _plan_kwargs(sp, M, K, 1, seed=SEED)
    kw["plan_index"] = PLAN
    noise = np.random.default_rng(53).standard_normal((1, M, 20)).astype(np.float32)
    be = EngineBackend(sp, w, M * KF, precision="f16x3")
    onehot = OC.switch_on_pix(inp["desig"], 2, 1, sp.height, sp.width, 1)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": onehot}
    res = be.plan(ctx, goal_pix=inp["goal"], noise=noise, k_futures=KF, lambda_variance=LAM, **kw)
    per_rollout = be.engine.score(inp["goal"], M=M * KF)
    assert np.all(np.isfinite(per_rollout))
    np.testing.assert_allclose(res["scores"][0], OC.reduce_futures(per_rollout, KF, LAM), rtol=1e-12)
    np.testing.assert_array_equal(res["elite_idx"], np.argsort(res["scores"][0], kind="stable")[:K])
    acts = be.engine.cem_actions()                                         # (M, T, adim) float64
    zs = OC.philox_latents(SEED, PLAN, 0, 2, sp.seq_len - 1, sp.nz)         # rollouts 0, 1 = futures 0, 1 of sequence 0
    rep = np.repeat(acts[:1].astype(np.float32), 2, axis=0)
    oi, od, _ = OracleMultiViewPredictor(sp, w, torch.float32).rollout(
        inp["frames"].astype(np.float32) / 255.0, inp["states"], onehot, Hh.step_actions(sp, inp["ctx_actions"], rep), zs)
    gi, gd = be.engine.fetch([0, 1])
    assert np.abs(gi - oi).max() <= FRAME_TOL
    assert np.abs(gd - od).max() <= 1e-5
    np.testing.assert_allclose(per_rollout[:2], OC.eval_pixel_cost(od, inp["goal"]), rtol=1e-5)
    be.engine.close()


@pytest.mark.parametrize("PREC", ["fp32_simt", "f16x3"])
def test_philox_sampling_matches_restatement_and_shards_are_invariant(PREC):
    """Device Philox normals == numpy restatement; a plan split into two shards (run back to back on one
    GPU, scores exchanged through the host) gives bit-identical scores and elites to the unsharded plan."""
    from visual_foresight_b200.predictor import EngineBackend, cem_params
    sp = S.spec_64(height=32, width=32, seq_len=4)
    w = Hh.make_weights(sp, seed=6)
    inp = Hh.synth_inputs(sp, seed=6)
    M, K, iters = 8, 3, 2
    kw = _plan_kwargs(sp, M, K, iters, seed=1234567890123)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": OC.switch_on_pix(inp["desig"], 2, 1, 32, 32, 1)}
    full = EngineBackend(sp, w, M, precision=PREC)
    res = full.plan(ctx, goal_pix=inp["goal"], **kw)
    # iteration-0 actions of the full plan vs the numpy Philox restatement
    full.set_context(ctx)
    p = cem_params(sp, n_ctx_actions=1, **{k: v for k, v in kw.items()})
    full.engine.cem_begin(p, inp["goal"].astype(np.float32))
    full.engine.cem_iter_rollout(0)
    a0 = full.engine.cem_actions()
    std = kw["std"]
    for m in (0, 5):
        for d in (0, 7, 19):
            z = OC.philox_normal(kw["seed"], 0, 0, m, d)
            want = np.clip(std[d % 4] * z, kw["clip"][0][d % 4], kw["clip"][1][d % 4])
            np.testing.assert_allclose(a0[m, (d // 4) * 3, d % 4], want, rtol=1e-10, atol=1e-14)
    # two shards
    halves = [EngineBackend(sp, w, M // 2, precision=PREC), EngineBackend(sp, w, M // 2, precision=PREC)]
    for r, b in enumerate(halves):
        b.set_context(ctx)
        pr = cem_params(sp, n_ctx_actions=1, global_samples=M, sample_offset=r * (M // 2),
                        **dict(kw, num_samples=M // 2))
        b.engine.cem_begin(pr, inp["goal"].astype(np.float32))
    for it in range(iters):
        seg = []
        for r, b in enumerate(halves):
            b.engine.cem_iter_rollout(it)
            seg.append(b.engine.cem_scores_read(it, r * (M // 2), M // 2))
        for r, b in enumerate(halves):
            b.engine.cem_scores_write(it, (1 - r) * (M // 2), seg[1 - r])
            b.engine.cem_iter_select(it)
    outs = [b.engine.cem_finish() for b in halves]
    for best, eidx, scores in outs:
        np.testing.assert_array_equal(scores, res["scores"])
        np.testing.assert_array_equal(eidx, res["elite_idx"])
        np.testing.assert_array_equal(best, res["best_actions"])
    for b in halves + [full]:
        b.engine.close()


# ---- policy surface -----------------------------------------------------------------------------------------
def test_controller_with_foreign_predictor_matches_reference(golden):
    """The reference's full act() golden, with the (foreign) BlobPredictor returning host arrays and the
    DEVICE cost kernel scoring them: same sampled actions, scores, elites and chosen action."""
    from fake_predictor import BlobPredictor
    from visual_foresight_b200.cem_controller import PixelCostController
    from visual_foresight_b200.policy import get_policy_args
    ag = {"adim": 4, "sdim": 4, "image_height": 48, "image_width": 64, "gpu_id": 0}
    pol = PixelCostController(ag, {"predictor_class": BlobPredictor, "rejection_sampling": False, "verbose": False,
                                   "num_samples": 24, "minimum_selection": 5}, 0, 1)
    pol.reset()
    np.random.seed(42)
    for t in range(3):
        obs = {"images": golden["act_images"][:t + 1], "state": golden["act_state"][:t + 1]}
        out = pol.act(**get_policy_args(pol, obs, t, 0, {"desig_pix": golden["act_desig"], "goal_pix": golden["act_goal"]}))
        np.testing.assert_allclose(out["actions"], golden["act_t%d_action" % t], rtol=1e-9, atol=1e-12)
        if t >= 1:
            for i in range(3):
                np.testing.assert_allclose(out["plan_stat"]["scores_itr%d" % i], golden["act_t%d_scores_itr%d" % (t, i)], rtol=1e-5)
            np.testing.assert_array_equal(pol._best_indices, golden["act_t%d_best_indices" % t])


def test_controller_device_path_end_to_end():
    """Policy.act() on the engine (device CEM): contract of the returned dict, determinism, replan gate."""
    from visual_foresight_b200.cem_controller import PixelCostController
    from visual_foresight_b200.policy import get_policy_args
    ag = {"adim": 4, "sdim": 4, "image_height": 32, "image_width": 32, "gpu_id": 0}
    pp = {"rejection_sampling": False, "verbose": False, "num_samples": 16, "minimum_selection": 4,
          "model_spec": {"seq_len": 6}, "replan_interval": 2, "cem_seed": 5}
    rng = np.random.default_rng(0)
    images = rng.integers(0, 256, (4, 1, 32, 32, 3), dtype=np.uint8)
    state = rng.uniform(-.5, .5, (4, 4))
    outs = []
    for rep in range(2):
        pol = PixelCostController(ag, dict(pp), 0, 1)
        pol.reset()
        acts = []
        for t in range(4):
            obs = {"images": images[:t + 1], "state": state[:t + 1]}
            o = pol.act(**get_policy_args(pol, obs, t, 0, {"desig_pix": np.array([[8, 8]]), "goal_pix": np.array([[24, 20]])}))
            assert o["actions"].shape == (4,)
            acts.append(o["actions"].copy())
            if t == 1:
                assert sorted(o["plan_stat"]) == ["scores_itr0", "scores_itr1", "scores_itr2"]
                assert o["plan_stat"]["scores_itr0"].shape == (16,) and np.all(np.isfinite(o["plan_stat"]["scores_itr2"]))
                first_plan = pol._best_actions.copy()
            if t == 2:      # replan_interval=2: step 2 replays the plan made at step 1
                np.testing.assert_array_equal(o["actions"], first_plan[0, 1])
        assert np.all(acts[0] == 0)
        outs.append(np.stack(acts))
        pol.predictor.backend.engine.close()
    np.testing.assert_array_equal(outs[0], outs[1])


# ---- tcgen05 implicit-GEMM convolution ------------------------------------------------------------------------
MMA_SHAPES = [(3, 32, 32, 64, 128, 5), (2, 16, 16, 128, 256, 5), (5, 8, 8, 256, 512, 5), (2, 16, 16, 64, 128, 3),
              (4, 6, 8, 64, 128, 5), (7, 8, 8, 32, 128, 5), (2, 12, 16, 32, 128, 5), (1, 24, 32, 32, 128, 3),
              # zero-padded channel chunks / output tiles (thin layers: Cout 3, 7, 32, 64; Cin 40, 56)
              (2, 16, 16, 56, 7, 3), (3, 16, 16, 32, 3, 3), (2, 32, 32, 40, 32, 3), (1, 64, 64, 64, 32, 3),
              (2, 16, 16, 128, 64, 3), (1, 32, 32, 128, 160, 3),
              # row-stacked thin path: 5x5 with an 8-channel input (first encoder conv), 48x64 frames, many samples, Cout 64
              (2, 64, 64, 8, 32, 5), (3, 48, 64, 8, 32, 5), (9, 32, 32, 32, 64, 3), (5, 24, 32, 64, 32, 3), (2, 8, 8, 16, 16, 5),
              # thin layer whose k*np exceeds 256 columns: one MMA per tap
              (2, 16, 16, 32, 64, 5),
              # CTA-pair kernel off the 16x16 map: 12x16 / 8x16 (48x64 and 32x64 inputs), two 256-channel tiles, odd sample count
              (3, 12, 16, 128, 256, 5), (2, 8, 16, 64, 512, 5), (1, 4, 16, 32, 256, 3)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", MMA_SHAPES)
@pytest.mark.parametrize("impl,tol", [(1, 1e-4), (2, 6e-3)])
def test_conv_mma_vs_fp64(eng_small, B, H, W, Cin, Cout, k, impl, tol):
    """tcgen05 conv (fp16 hi/lo split x3 = fp32-grade; single pass = fp16-grade) vs a float64 convolution and
    vs the fp32 FFMA kernel on the same inputs."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(B * 100 + Cin + k)
    x = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    y = eng_small.debug_conv2d(x, w, b, impl=impl)
    ref = F.conv2d(torch.from_numpy(x).double().permute(0, 3, 1, 2), torch.from_numpy(w).double().permute(3, 2, 0, 1),
                   torch.from_numpy(b).double(), padding=k // 2).permute(0, 2, 3, 1).numpy()
    err = np.abs(y - ref).max()
    if impl == 1:
        # the tensor core's fp32 accumulation rounds toward zero: the bias grows with the accumulation-chain length
        tol = max(tol, 4e-8 * k * k * Cin)
    assert err < tol, "max abs err %g" % err
    if impl == 1:
        ys = eng_small.debug_conv2d(x, w, b, impl=0)
        assert np.abs(y - ys).max() < tol      # dominated by the tensor core's fp32 accumulate rounding (~2^-15 of |y| at K=1600)


def test_full_horizon_c2_frames_within_tolerance():
    """BASELINE c2 geometry (S=15, 64x64): 14 recurrent cell steps with the conv-LSTM convolutions on tcgen05
    (fp16 hi/lo x3) stay within 1e-4 max-abs of the fp32 oracle on every one of the 13 predicted frames, and the
    elite set computed from device scores equals the one computed from oracle scores."""
    sp = S.spec_64(height=64, width=64, seq_len=15)
    w = Hh.make_weights(sp, seed=0)
    inp = Hh.synth_inputs(sp, seed=0)
    acts = Hh.gaussian_actions(sp, 6, 15, seed=0)
    e, (gi, gd, gs) = _engine_rollout(sp, w, inp, acts, precision="f16x3")
    oi, od, os_ = Hh.oracle_rollout(sp, w, inp, acts)
    per_frame = np.abs(gi - oi).max(axis=(0, 2, 3, 4, 5))
    print("c2 horizon, f16x3: per-frame max-abs err", np.array2string(per_frame, precision=2))
    assert per_frame.max() <= FRAME_TOL
    sc = e.score(inp["goal"], M=6)
    osc = OC.eval_pixel_cost(od, inp["goal"])
    np.testing.assert_allclose(sc, osc, rtol=1e-5)
    np.testing.assert_array_equal(e.topk(sc, 3), OC.elite_select(osc, 3))
    e.close()


@pytest.mark.parametrize("precision,tol", [("f16x3", FRAME_TOL), ("f16x1", 5e-2)])
def test_rollout_tensor_core_path_vs_oracle(precision, tol):
    """Full predictor rollout with the conv-LSTM convolutions on tcgen05: the fp32-grade mode holds the 1e-4
    frame tolerance of BASELINE.json; the single-pass mode is reported with its own (looser) bound."""
    sp = S.spec_64(height=64, width=64, seq_len=6)
    w = Hh.make_weights(sp, seed=11)
    inp = Hh.synth_inputs(sp, seed=11)
    acts = Hh.gaussian_actions(sp, 5, 15, seed=11)
    e, (gi, gd, gs) = _engine_rollout(sp, w, inp, acts, precision=precision)
    oi, od, os_ = Hh.oracle_rollout(sp, w, inp, acts)
    err = float(np.abs(gi - oi).max())
    print("precision %s: frames max-abs err %.3g, distrib %.3g" % (precision, err, float(np.abs(gd - od).max())))
    assert err <= tol
    e.close()


def test_goal_image_cost_and_controller():
    """SURVEY a7 (goal_im_controller.py:87-93): score = mean squared error of the final predicted frame of view 0 against
    the goal image.  The device kernel against numpy on the fetched frames (rel 1e-5), then GoalImController.act() end to
    end: zeros before t = n_context, a plan afterwards whose recorded scores are that cost of the sampled actions."""
    from visual_foresight_b200.cem_controller import GoalImController
    from visual_foresight_b200.engine import COST_GOAL_IMAGE
    sp = S.spec_64(height=32, width=32, seq_len=6)
    w = Hh.make_weights(sp, seed=31)
    inp = Hh.synth_inputs(sp, seed=32)
    acts = Hh.gaussian_actions(sp, 6, 6, seed=33)
    e, (gi, gd, gs) = _engine_rollout(sp, w, inp, acts, precision="f16x3")
    goal = np.random.default_rng(34).uniform(0, 1, (32, 32, 3)).astype(np.float32)
    sc = e.score(goal, cost_kind=COST_GOAL_IMAGE, M=6)
    want = ((gi[:, -1, 0].astype(np.float64) - goal) ** 2).mean(axis=(1, 2, 3))
    np.testing.assert_allclose(sc, want, rtol=1e-5)
    e.close()

    ag = {"adim": 4, "sdim": 4, "image_height": 32, "image_width": 32, "gpu_id": 0}
    pp = {"num_samples": 12, "minimum_selection": 4, "iterations": 2, "rejection_sampling": False, "verbose": False,
          "model_spec": {"height": 32, "width": 32, "seq_len": 15}, "model_seed": 31}
    pol = GoalImController(ag, pp, 0, 1)
    pol.reset()
    rng = np.random.default_rng(35)
    images = rng.integers(0, 256, size=(4, 1, 32, 32, 3), dtype=np.uint8)
    state = rng.uniform(-0.5, 0.5, size=(4, 4))
    goal_u8 = rng.integers(0, 256, size=(1, 1, 32, 32, 3), dtype=np.uint8)
    np.random.seed(3)
    for t in range(3):
        out = pol.act(t=t, i_tr=0, goal_image=goal_u8, images=images[:t + 1], state=state[:t + 1])
        assert out["actions"].shape == (4,)
        if t < 2:
            assert np.all(out["actions"] == 0)                  # start_planning = n_context (goal_im_controller.py:35)
    s0, s1 = out["plan_stat"]["scores_itr0"], out["plan_stat"]["scores_itr1"]
    assert s0.shape == (12,) and np.all(np.isfinite(s0)) and np.all(s0 > 0) and np.all(s0 < 1)
    best = pol._backend.engine.fetch([int(np.argmin(s1))])[0][0]
    np.testing.assert_allclose(s1.min(), ((best[-1, 0].astype(np.float64) - goal_u8[0, 0] / 255.0) ** 2).mean(), rtol=1e-5)


@pytest.mark.parametrize("precision", ["fp32_simt", "f16x3"])
def test_shared_prefix_steps_are_bit_identical(precision, monkeypatch):
    """The cell steps fed only by context (frame, state AND action: tau < min(n_ctx_actions, C-1)) are the same for every
    sample, so the engine runs them once on one sample and replicates the recurrent state (engine.cu: rollout_body).
    With three context frames (two shared steps) the predicted frames, distributions and states equal, bit for bit, the
    rollout that recomputes the prefix for every sample (VF_SHARED_PREFIX=0), and both stay within tolerance of the oracle."""
    from visual_foresight_b200.engine import Engine
    sp = S.spec_64(height=32, width=32, seq_len=7, context_frames=3)
    w = Hh.make_weights(sp, seed=21)
    inp = Hh.synth_inputs(sp, seed=22)
    assert np.asarray(inp["ctx_actions"]).shape[0] == 2
    acts = Hh.gaussian_actions(sp, 5, 6, seed=23)
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("VF_SHARED_PREFIX", flag)
        e = Engine(sp, 5, precision=precision)
        e.load_weights(w)
        e.set_context(inp["frames"], inp["states"], inp["ctx_actions"])
        e.set_desig(inp["desig"])
        outs.append(e.predict(acts))
        outs.append(e.predict(acts))          # second call replays the captured graph
        e.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            np.testing.assert_array_equal(a, b)
    oi, od, os_ = Hh.oracle_rollout(sp, w, inp, acts)
    assert np.abs(outs[0][0] - oi).max() <= FRAME_TOL
    assert np.abs(outs[0][1] - od).max() <= 1e-5


def test_plan_prefix_cache_is_bit_identical(monkeypatch):
    """Within one plan the context does not change, so CEM iterations 1.. start from the recurrent state iteration 0 saved
    after the shared-prefix steps instead of recomputing them (engine.cu: PREFIX_SAVE / PREFIX_RESTORE).  Two consecutive
    plans (the second replays captured graphs, with a NEW context in between) give bit-identical scores, elites and actions
    with the cache on and off."""
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_64(height=32, width=32, seq_len=6)
    w = Hh.make_weights(sp, seed=41)
    M, K = 12, 4
    results = []
    for flag in ("1", "0"):
        monkeypatch.setenv("VF_PREFIX_CACHE", flag)
        be = EngineBackend(sp, w, M, precision="f16x3")
        out = []
        for plan in range(3):
            inp = Hh.synth_inputs(sp, seed=42 + plan % 2)                 # the context changes between plans
            onehot = OC.switch_on_pix(inp["desig"], 2, 1, sp.height, sp.width, 1)
            ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
                   "context_pixel_distributions": onehot}
            kw = _plan_kwargs(sp, M, K, 3, seed=7)
            kw["plan_index"] = plan
            res = be.plan(ctx, goal_pix=inp["goal"], **kw)
            out.append((res["scores"].copy(), res["elite_idx"].copy(), res["best_actions"].copy()))
        be.engine.close()
        results.append(out)
    for a, b in zip(*results):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    assert not np.array_equal(results[0][0][0], results[0][1][0])       # different contexts really give different plans


# ---- multi-GPU (only when the box has >= 2 GPUs; the 1-GPU round-end run skips it) ------------------------------
def test_two_gpu_sharded_plan_is_bit_identical():
    """torchrun x2: contiguous shards + NCCL in-place all-gather of the float64 scores per CEM iteration give the
    same scores / elite indices / best actions, bit for bit, as the single-GPU plan (tests/multigpu_check.py)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29531", os.path.join(here, "multigpu_check.py")],
                       capture_output=True, text=True, timeout=300)
    assert "MULTIGPU_CHECK_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

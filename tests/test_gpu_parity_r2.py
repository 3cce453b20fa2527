"""GPU parity, round 2: the headline arithmetic on a WHOLE plan, and the device branches round 1 left unchecked
(VERDICT r01 "What's weak" #1): non-uniform task weights, the device warm start (reuse_mean / reuse_factor / reduce_std_dev),
predictor_propagation from the device, the legacy setup_predictor hook, and the engine-owned peer-memory score exchange.
Everything goes ctypes -> libvfengine.so; the oracle (oracle/, CPU) is the checker only.

Tolerances: scores rel 1e-5, frames 1e-4 max-abs, distributions 1e-5, sampled / best actions 1e-12, elite index SETS exact."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as Hh
from oracle import cem as OC
from visual_foresight_b200 import spec as S

pytestmark = pytest.mark.gpu
FRAME_TOL = 1e-4


def _plan_kwargs(sp, M, K, iters, seed=0, nactions=5, repeat=3):
    from visual_foresight_b200.samplers import GaussianCEMSampler, action_bounds, per_dim_variance
    from visual_foresight_b200.hparams import HParams
    hp = HParams(**GaussianCEMSampler.get_default_hparams())
    lo, hi = action_bounds(hp, sp.adim)
    return dict(num_samples=M, iterations=iters, num_elites=K, nactions=nactions, repeat=repeat,
                std=np.sqrt(per_dim_variance(hp, sp.adim)), clip=(lo, hi), mean0=None, reduce_std_scale=1.0,
                finalweight=10.0, task_weights=None, seed=seed, plan_index=0)


def _oracle_eval(sp, w, frames_u8, states, distrib, ctx_actions, goal, task_weights=None):
    """evaluate(actions) for OC.cem_plan: oracle rollout from an explicit context (distribution given, not rebuilt)."""
    import torch
    from oracle.predictor import OracleMultiViewPredictor
    pred = OracleMultiViewPredictor(sp, w, torch.float32)
    fr = np.asarray(frames_u8, np.float32) / 255.0

    def evaluate(actions):
        sa = Hh.step_actions(sp, ctx_actions, np.asarray(actions, np.float32))
        _, od, _ = pred.rollout(fr, states if sp.sdim else None, distrib, sa)
        return OC.eval_pixel_cost(od, goal, task_weights=task_weights)
    return evaluate


# ---- 1. the headline config, whole plan, product arithmetic ----------------------------------------------------------
@pytest.mark.timeout(1500)
def test_c2_full_plan_f16x3_elite_sets_vs_oracle():
    """BASELINE c2 exactly (M=200, S=15, 64x64, 3 CEM iterations, K=10) on the tensor-core path (f16x3) with the SAME
    standard-normal noise as the oracle planner: every iteration's scores within rel 1e-5, the elite index SET of every
    iteration identical (CEM feeds elites forward: one flipped elite in iteration 0 would change iterations 1 and 2
    entirely), final elite order and best actions equal."""
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_64(height=64, width=64, seq_len=15, adim=4, sdim=4)
    w = Hh.make_weights(sp, seed=0)
    inp = Hh.synth_inputs(sp, seed=0)
    M, K, iters = 200, 10, 3
    kw = _plan_kwargs(sp, M, K, iters)
    noise = np.random.default_rng(2026).standard_normal((iters, M, 20)).astype(np.float32)
    onehot = OC.switch_on_pix(inp["desig"], 2, 1, 64, 64, 1)
    be = EngineBackend(sp, w, M, precision="f16x3")
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": onehot}
    res = be.plan(ctx, goal_pix=inp["goal"], noise=noise, **kw)
    last_actions = be.engine.cem_actions()
    be.engine.close()
    evaluate = _oracle_eval(sp, w, inp["frames"], inp["states"], onehot, inp["ctx_actions"], inp["goal"])
    best, idx, scores, all_actions = OC.cem_plan(evaluate, num_samples=M, iterations=iters, num_elites_k=K, nactions=5,
                                                 repeat=3, adim=4, std=kw["std"], noise=noise, clip=kw["clip"])
    for it in range(iters):
        np.testing.assert_allclose(res["scores"][it], scores[it], rtol=1e-5, err_msg="iteration %d" % it)
        got = set(np.argsort(res["scores"][it], kind="stable")[:K].tolist())
        want = set(OC.elite_select(scores[it], K).tolist())
        assert got == want, "elite set of iteration %d differs: %s vs %s" % (it, sorted(got), sorted(want))
    np.testing.assert_array_equal(res["elite_idx"], idx)
    np.testing.assert_allclose(res["best_actions"], best, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(last_actions, all_actions[-1], rtol=1e-12, atol=1e-15)


# ---- 2. registration-weighted cost -----------------------------------------------------------------------
def test_nonuniform_task_weights_two_views_two_pixels():
    """ncam=2 x ndesig=2 with the 1/warp-error trade-off weights of register_gtruth_controller.py:88-91 (non-uniform):
    vf_score and a whole device plan against OC.eval_pixel_cost(task_weights=...)."""
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_64(height=32, width=32, seq_len=6, ncam=2, ndesig=2, adim=4, sdim=5)
    w = Hh.make_weights(sp, seed=11)
    inp = Hh.synth_inputs(sp, seed=12)
    err = np.array([0.8, 2.5, 1.3, 0.4])
    tw = (1.0 / err) / (1.0 / err).sum()                               # tradeoff = (1/warperr) / sum(1/warperr)
    onehot = OC.switch_on_pix(inp["desig"], 2, 2, 32, 32, 2)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": onehot}
    M, K, iters = 12, 4, 2
    be = EngineBackend(sp, w, M, precision="f16x3")
    acts = Hh.gaussian_actions(sp, M, 15, seed=13)
    gi, gd, _ = be.predict(ctx, acts)
    sc = be.engine.score(inp["goal"].astype(np.float32), task_weights=tw, M=M)
    np.testing.assert_allclose(sc, OC.eval_pixel_cost(gd, inp["goal"], task_weights=tw), rtol=1e-5)     # device cost kernel on device frames
    _, od, _ = Hh.oracle_rollout(sp, w, inp, acts)
    np.testing.assert_allclose(sc, OC.eval_pixel_cost(od, inp["goal"], task_weights=tw), rtol=1e-5)     # against the oracle rollout
    uniform = be.engine.score(inp["goal"].astype(np.float32), M=M)
    assert np.abs(uniform - sc).max() > 1e-3                            # the weights really matter
    kw = _plan_kwargs(sp, M, K, iters)
    kw["task_weights"] = tw
    noise = np.random.default_rng(5).standard_normal((iters, M, 20)).astype(np.float32)
    res = be.plan(ctx, goal_pix=inp["goal"], noise=noise, **kw)
    evaluate = _oracle_eval(sp, w, inp["frames"], inp["states"], onehot, inp["ctx_actions"], inp["goal"], task_weights=tw)
    best, idx, scores, _ = OC.cem_plan(evaluate, num_samples=M, iterations=iters, num_elites_k=K, nactions=5, repeat=3,
                                       adim=4, std=kw["std"], noise=noise, clip=kw["clip"])
    np.testing.assert_allclose(res["scores"], scores, rtol=1e-5)
    np.testing.assert_array_equal(res["elite_idx"], idx)
    np.testing.assert_allclose(res["best_actions"], best, rtol=1e-12, atol=1e-15)
    be.engine.close()


# ---- 3. device warm start + predictor_propagation through the policy --------------------------------------------------
def _philox_noise(seed, plan, iters, M, width):
    """the device sampler's draws (float64 Box-Muller on Philox4x32-10, cem.cu) restated by the oracle"""
    zz = np.zeros((iters, M, width), np.float64)
    for it in range(iters):
        for m in range(M):
            for j in range(width):
                zz[it, m, j] = OC.philox_normal(seed, plan, it, m, j)
    return zz


@pytest.mark.timeout(900)
def test_device_warm_start_and_propagation_vs_oracle():
    """PixelCostController on the device path with reuse_mean, reuse_factor, reduce_std_dev and predictor_propagation
    (gaussian_sampler.py:16-44, controller_utils.py:76-96, pixel_cost_controller.py:161-165,199-204): every plan the policy
    issues (recorded at the backend boundary) is replayed by the oracle planner with the oracle predictor on the bit-exact
    Philox restatement of the device noise — the shifted warm-start mean, the shrunken sample count, the variance scaling of
    all but the last action block and the propagated distribution all have to agree for the scores / elites / actions to."""
    from visual_foresight_b200.cem_controller import PixelCostController
    from visual_foresight_b200.policy import get_policy_args
    H = W = 32
    ag = {"adim": 4, "sdim": 4, "image_height": H, "image_width": W, "gpu_id": 0}
    pp = {"rejection_sampling": False, "verbose": False, "num_samples": 16, "minimum_selection": 4, "iterations": 2,
          "model_spec": {"seq_len": 6}, "cem_seed": 17, "model_seed": 31, "reuse_mean": True,      # reuse_factor stays 0.5 (default)
          "reduce_std_dev": 0.25, "predictor_propagation": True}
    pol = PixelCostController(ag, dict(pp), 0, 1)
    hp = pol._hp
    sp = pol.predictor.spec
    w = [S.init_weights(sp, 31, 0)]
    pol.reset()
    calls = []
    real_plan = pol._backend.plan

    def spy(context, **kw):
        res = real_plan(context, **kw)
        calls.append((dict(context), dict(kw), {k: np.array(v) for k, v in res.items()}))
        return res
    pol._backend.plan = spy
    rng = np.random.default_rng(3)
    T_ep = 5
    images = rng.integers(0, 256, (T_ep, 1, H, W, 3), dtype=np.uint8)
    state = rng.uniform(-.5, .5, (T_ep, 4))
    desig, goal = np.array([[9, 7]]), np.array([[22, 25]])
    for t in range(T_ep):
        obs = {"images": images[:t + 1], "state": state[:t + 1]}
        out = pol.act(**get_policy_args(pol, obs, t, 0, {"desig_pix": desig, "goal_pix": goal}))
        assert out["actions"].shape == (4,)
    assert len(calls) == T_ep - 1                               # planning starts at t = n_context - 1 = 1
    warm = [c for c in calls if c[1]["mean0"] is not None]
    assert warm, "reuse_mean never produced a warm start"
    assert any(c[1]["num_samples"] < hp.num_samples for c in calls), "reuse_factor never shrank the sample count"
    assert any(c[1]["reduce_std_scale"] != 1.0 for c in calls), "reduce_std_dev never applied"
    for ci, (context, kw, res) in enumerate(calls):
        M, K, iters = kw["num_samples"], kw["num_elites"], kw["iterations"]
        D = kw["nactions"] * sp.adim
        noise = _philox_noise(kw["seed"], kw["plan_index"], iters, M, max(D, K))
        frames = np.asarray(context["context_frames"])[-sp.context_frames:]
        states = np.asarray(context["context_states"], np.float32)[-sp.context_frames:]
        distrib = np.asarray(context["context_pixel_distributions"], np.float32)[-sp.context_frames:]
        ca = np.asarray(context["context_actions"], np.float32).reshape(-1, sp.adim)[-(sp.context_frames - 1):]
        if ci > 0:                                                   # propagated distribution = best rollout of the previous plan
            assert not np.array_equal(distrib, pol._switch_on_pix(pol._desig_pix)), "predictor_propagation not in effect"
            np.testing.assert_allclose(distrib.sum(axis=(2, 3)), 1.0, atol=1e-5)
        evaluate = _oracle_eval(sp, w, frames, states, distrib, ca, np.asarray(kw["goal_pix"], np.float64).reshape(1, 1, 2))
        best, idx, scores, _ = OC.cem_plan(evaluate, num_samples=M, iterations=iters, num_elites_k=K, nactions=kw["nactions"],
                                           repeat=kw["repeat"], adim=sp.adim, std=kw["std"], noise=noise, clip=kw["clip"],
                                           mean0=kw["mean0"], reduce_std_scale=kw["reduce_std_scale"])
        np.testing.assert_allclose(res["scores"], scores, rtol=2e-5, err_msg="plan %d" % ci)
        np.testing.assert_array_equal(res["elite_idx"], idx)
        np.testing.assert_allclose(res["best_actions"], best, rtol=1e-9, atol=1e-12)
    # the distribution the policy carries forward is the device rollout of the best sample of the LAST plan
    context, kw, res = calls[-1]
    import torch
    from oracle.predictor import OracleMultiViewPredictor
    sa = Hh.step_actions(sp, np.asarray(context["context_actions"], np.float32).reshape(-1, sp.adim)[-(sp.context_frames - 1):],
                         res["best_actions"][:1].astype(np.float32))
    _, od, _ = OracleMultiViewPredictor(sp, w, torch.float32).rollout(
        np.asarray(context["context_frames"])[-sp.context_frames:].astype(np.float32) / 255.0,
        np.asarray(context["context_states"], np.float32)[-sp.context_frames:],
        np.asarray(context["context_pixel_distributions"], np.float32)[-sp.context_frames:], sa)
    assert np.abs(pol._chosen_distrib - od[0]).max() <= 1e-5
    pol.predictor.backend.engine.close()


# ---- 4. legacy setup_predictor hook ------------------------------------------------------------------------
def test_setup_predictor_hook_vs_oracle():
    """netconf['setup_predictor'](...) -> predictor_func(input_images, input_one_hot_images, input_state, input_actions)
    -> (gen_images, gen_distrib, gen_states) in the reference's shapes (setup_predictor.py:98-114,164-200): the legacy
    hook feeds ALL S-1 actions from input_actions (no context actions)."""
    import torch
    from oracle.predictor import OracleMultiViewPredictor
    from visual_foresight_b200.predictor import setup_predictor
    H, W, M = 32, 32, 5
    conf = {"orig_size": [H, W], "ncam": 1, "ndesig": 1, "adim": 4, "sdim": 4, "sequence_length": 6, "context_frames": 2,
            "batch_size": M, "model_seed": 8}
    pf = setup_predictor({}, conf, gpu_id=0, ngpu=1, logger=None)
    sp = pf.backend.spec
    w = [S.init_weights(sp, 8, 0)]
    inp = Hh.synth_inputs(sp, seed=9)
    images = (inp["frames"].astype(np.float32) / 255.0)[None]                  # (1, C, ncam, H, W, 3) like pred_util.get_context
    onehot = OC.switch_on_pix(inp["desig"], 2, 1, H, W, 1)
    acts = Hh.gaussian_actions(sp, M, sp.seq_len - 1, seed=10)
    gi, gd, gs = pf(input_images=images, input_one_hot_images=onehot[None], input_state=inp["states"][None], input_actions=acts)
    assert gi.shape == (M, sp.n_pred, 1, H, W, 3) and gd.shape == (M, sp.n_pred, 1, H, W, 1) and gs.shape == (M, sp.n_pred, 4)
    oi, od, os_ = OracleMultiViewPredictor(sp, w, torch.float32).rollout(images[0], inp["states"], onehot, acts)
    assert np.abs(gi - oi).max() <= FRAME_TOL
    assert np.abs(gd - od).max() <= 1e-5
    np.testing.assert_allclose(gs, os_, atol=1e-5)
    gi2, gd2, _ = pf(input_images=images, input_one_hot_images=None, input_state=inp["states"][None], input_actions=acts)
    assert gd2 is None and np.array_equal(gi2, gi)                          # no designated pixels: frames only (setup_predictor.py:186-198)
    pf.backend.engine.close()


# ---- 5. engine-owned peer-memory score exchange ---------------------------------------------------------------------------
def test_peer_exchange_two_handles_one_process(monkeypatch):
    """vf_comm_export / vf_comm_connect / vf_cem_exchange with two handles of ONE process on one device (plain pointers):
    the sharded plan over the engine's exchange kernel is bit-identical to the single-handle plan, over consecutive plans
    (window reuse, arrival-counter epochs)."""
    from visual_foresight_b200.predictor import EngineBackend, cem_params
    monkeypatch.setenv("VF_NO_GRAPH", "1")        # graph instantiation may synchronise the device while a peer kernel is waiting
    monkeypatch.setenv("VF_COMM_TIMEOUT_MS", "20000")
    sp = S.spec_64(height=32, width=32, seq_len=6)
    w = Hh.make_weights(sp, seed=3)
    inp = Hh.synth_inputs(sp, seed=3)
    Mg, K, iters = 16, 4, 3
    kw = _plan_kwargs(sp, Mg, K, iters, seed=99)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": OC.switch_on_pix(inp["desig"], 2, 1, 32, 32, 1)}
    full = EngineBackend(sp, w, Mg, precision="f16x3")
    refs = []
    for pi in range(3):
        k2 = dict(kw)
        k2["plan_index"] = pi
        refs.append(full.plan(ctx, goal_pix=inp["goal"], **k2))
    full.engine.close()
    halves = [EngineBackend(sp, w, Mg // 2, precision="f16x3") for _ in range(2)]
    descs = [b.engine.comm_export(8, 64) for b in halves]
    for r, b in enumerate(halves):
        b.engine.comm_connect(r, 2, descs)
        b.set_context(ctx)
        b.engine.predict(Hh.gaussian_actions(sp, Mg // 2, 15, seed=1), fetch=False)   # weights finalised, buffers allocated up front
    goal = np.asarray(inp["goal"], np.float32)
    for pi in range(3):
        for r, b in enumerate(halves):
            k2 = {k: v for k, v in kw.items() if k != "num_samples"}
            k2["plan_index"] = pi
            p = cem_params(sp, num_samples=Mg // 2, global_samples=Mg, sample_offset=r * (Mg // 2),
                           n_ctx_actions=b._n_ctx_actions, **k2)
            b.engine.cem_begin(p, goal)
        for it in range(iters):
            for b in halves:                       # nothing here blocks the host: rank 0's exchange kernel waits ON THE GPU
                b.engine.cem_iter_rollout(it)      # for rank 1's scores while rank 1's rollout runs beside it
                b.engine.cem_exchange(it)
            for b in halves:
                b.engine.cem_iter_select(it)
        outs = [b.engine.cem_finish() for b in halves]
        for best, eidx, scores in outs:
            np.testing.assert_array_equal(scores, refs[pi]["scores"])
            np.testing.assert_array_equal(eidx, refs[pi]["elite_idx"])
            np.testing.assert_array_equal(best, refs[pi]["best_actions"])
    for b in halves:
        b.engine.close()


@pytest.mark.timeout(900)
def test_two_rank_policy_and_planner_one_gpu_two_processes():
    """torchrun x2 with both ranks on cuda:0 (gloo process group): the exchange windows are mapped ACROSS PROCESSES with
    cudaIpc, the planner over the peer exchange and PixelCostController(ngpu=2).act() on both ranks equal the single-rank
    results bit for bit (tests/multigpu_check.py).  This is the multi-rank path a 1-GPU box can exercise."""
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ)
    env["VF_CHECK_SAME_DEVICE"] = "1"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(here, "multigpu_check.py")],
                       capture_output=True, text=True, timeout=800, env=env)
    assert "MULTIGPU_CHECK_OK world=2" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ---- 6. sampler families / options on the device path --------------------------------------------------
def test_device_correlated_noise_plan_vs_oracle():
    """CorrelatedNoiseSampler on the device (vf_cem_params.sampler = VF_SAMPLER_CORRELATED): AR(1)-smoothed noise with the
    reference's index -1 wrap, softmax-weighted elite mean (correlated_noise.py:17-66), a constant appended action dim
    (cem_base_controller.py:94-96), against OC.cem_plan_correlated on the same standard normals."""
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_64(height=32, width=32, seq_len=6, adim=5, sdim=4)           # 4 sampled dims + 1 appended constant
    w = Hh.make_weights(sp, seed=21)
    inp = Hh.synth_inputs(sp, seed=22)
    M, K, iters, nact = 14, 5, 3, 7
    std = np.array([0.05, 0.05, 0.2, np.pi / 10])
    bias = np.array([0.01, -0.02, 0.0, 0.05])
    noise = np.random.default_rng(8).standard_normal((iters, M, nact * 4)).astype(np.float32)
    onehot = OC.switch_on_pix(inp["desig"], 2, 1, 32, 32, 1)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": onehot}
    be = EngineBackend(sp, w, M, precision="f16x3")
    res = be.plan(ctx, num_samples=M, iterations=iters, num_elites=K, nactions=nact, repeat=1, std=std, clip=None, mean0=None,
                  reduce_std_scale=1.0, goal_pix=inp["goal"], finalweight=10.0, task_weights=None, seed=0, plan_index=0, noise=noise,
                  sampler="correlated", beta0=0.6, beta1=0.4, kappa=2.0, mean_bias=bias, append_action=[0.25])
    last_actions = be.engine.cem_actions()
    be.engine.close()
    evaluate = _oracle_eval(sp, w, inp["frames"], inp["states"], onehot, inp["ctx_actions"], inp["goal"])
    best, idx, scores, all_actions = OC.cem_plan_correlated(evaluate, num_samples=M, iterations=iters, num_elites_k=K, nactions=nact,
                                                            adim=4, initial_std=std, noise=noise, beta0=0.6, beta1=0.4, kappa=2.0,
                                                            mean_bias=bias, append_action=[0.25])
    assert best.shape == (K, nact, 5) and np.all(best[..., 4] == 0.25)
    np.testing.assert_allclose(res["scores"], scores, rtol=1e-5)
    np.testing.assert_array_equal(res["elite_idx"], idx)
    # This sampler's next mean is weighted by exp(kappa * score VALUES) (not only their ranks), so the actions of iterations >= 1
    # inherit the predictor tolerance of the scores: 1e-5 relative on scores of ~30 px, times kappa = 2 -> weights differ by
    # ~6e-4 relative -> actions by ~1e-6.  Against the all-oracle plan that is the achievable bound ...
    np.testing.assert_allclose(res["best_actions"], best, rtol=0, atol=2e-5)
    np.testing.assert_allclose(last_actions, all_actions[-1], rtol=0, atol=2e-5)
    # ... and the sampler arithmetic itself is checked tightly by replaying the recursion with the DEVICE's scores
    mean = np.zeros((nact, 4))
    for it in range(iters):
        z = np.asarray(noise[it], np.float64).reshape(M, nact, 4)
        acts = OC.correlated_noise(z, std, 0.6, 0.4, bias) + mean[None]
        eidx = OC.elite_select(res["scores"][it], K)
        if it < iters - 1:
            mean = OC.correlated_elite_mean(acts[eidx], res["scores"][it][eidx], 2.0)
    np.testing.assert_allclose(last_actions[..., :4], acts, rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(res["best_actions"][..., :4], acts[eidx], rtol=1e-11, atol=1e-14)
    assert np.all(last_actions[..., 4] == 0.25)


def test_device_gaussian_discrete_and_appended_dims_vs_oracle():
    """Gaussian CEM on the device with discrete_ind (floor + clip to [0,4] before truncate_movement, gaussian_sampler.py:87-88,
    controller_utils.py:107-117) and append_action, against the oracle planner on the same noise."""
    from visual_foresight_b200.predictor import EngineBackend
    sp = S.spec_64(height=32, width=32, seq_len=6, adim=5, sdim=4)
    w = Hh.make_weights(sp, seed=31)
    inp = Hh.synth_inputs(sp, seed=32)
    M, K, iters = 12, 4, 3
    kw = _plan_kwargs(S.spec_64(height=32, width=32, seq_len=6, adim=4, sdim=4), M, K, iters)     # std / clip of the 4 sampled dims
    kw["std"] = np.array([0.05, 0.05, 1.5, np.pi / 18])            # dim 2 wide enough for several discrete levels
    lo, hi = kw["clip"]
    noise = np.random.default_rng(12).standard_normal((iters, M, 20)).astype(np.float32)
    onehot = OC.switch_on_pix(inp["desig"], 2, 1, 32, 32, 1)
    ctx = {"context_frames": inp["frames"], "context_states": inp["states"], "context_actions": inp["ctx_actions"],
           "context_pixel_distributions": onehot}
    be = EngineBackend(sp, w, M, precision="f16x3")
    res = be.plan(ctx, goal_pix=inp["goal"], noise=noise, discrete_ind=[2], append_action=[-0.5], **kw)
    be.engine.close()
    evaluate = _oracle_eval(sp, w, inp["frames"], inp["states"], onehot, inp["ctx_actions"], inp["goal"])
    best, idx, scores, all_actions = OC.cem_plan(evaluate, num_samples=M, iterations=iters, num_elites_k=K, nactions=5, repeat=3,
                                                 adim=4, std=kw["std"], noise=noise, clip=(lo, hi), discrete_ind=[2], append_action=[-0.5])
    assert set(np.unique(all_actions[0][..., 2])) <= {0.0, 1.0, 2.0, 3.0, 4.0} and len(np.unique(all_actions[0][..., 2])) > 1
    np.testing.assert_allclose(res["scores"], scores, rtol=1e-5)
    np.testing.assert_array_equal(res["elite_idx"], idx)
    np.testing.assert_allclose(res["best_actions"], best, rtol=1e-12, atol=1e-15)


def test_controller_takes_the_device_path_for_correlated_noise():
    """A RoboNet/Franka-style policy (sampler = CorrelatedNoiseSampler, experiments/robonet/franka/franka.py:37-55) plans on the
    device: _device_path_ok(), no sampler round trip, deterministic in cem_seed."""
    from visual_foresight_b200.cem_controller import PixelCostController
    from visual_foresight_b200.policy import get_policy_args
    from visual_foresight_b200.samplers import CorrelatedNoiseSampler
    ag = {"adim": 4, "sdim": 4, "image_height": 32, "image_width": 32, "gpu_id": 0}
    pp = {"sampler": CorrelatedNoiseSampler, "verbose": False, "num_samples": 16, "minimum_selection": 4, "iterations": 2,
          "model_spec": {"seq_len": 6}, "cem_seed": 3, "nactions": 6}
    rng = np.random.default_rng(1)
    images = rng.integers(0, 256, (3, 1, 32, 32, 3), dtype=np.uint8)
    state = rng.uniform(-.5, .5, (3, 4))
    runs = []
    for rep_ in range(2):
        pol = PixelCostController(ag, dict(pp), 0, 1)
        pol.reset()
        assert pol._device_path_ok()
        called = []
        pol._sampler.sample_initial_actions = lambda *a, **k: called.append(1)      # the host sampler must not be consulted
        outs = []
        for t in range(3):
            obs = {"images": images[:t + 1], "state": state[:t + 1]}
            o = pol.act(**get_policy_args(pol, obs, t, 0, {"desig_pix": np.array([[8, 8]]), "goal_pix": np.array([[20, 22]])}))
            outs.append(o["actions"].copy())
        assert not called and pol._best_actions.shape == (4, 6, 4)
        assert sorted(o["plan_stat"]) == ["scores_itr0", "scores_itr1"]
        runs.append(np.stack(outs))
        pol.predictor.backend.engine.close()
    np.testing.assert_array_equal(runs[0], runs[1])
    assert np.abs(runs[0][1:]).max() > 0

"""Registration front end (visual_foresight_b200/registration.py) against the unmodified reference's get_warp_err and
trade-off normalisation (fixtures: tests/golden/make_registration_golden.py -> ref_registration_golden.npz)."""
import os

import numpy as np
import pytest

from visual_foresight_b200 import registration as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_registration_golden.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def _run(g, tag, region, H, W):
    f64 = lambda k: g["%s_%s" % (tag, k)].astype(np.float64)
    errs, pix = [], []
    for cam in range(2):
        e, d = R.warp_errors(cam, f64("start"), f64("goal"), f64("spts"), f64("gpts"), f64("wstart"), f64("wgoal"),
                             g[tag + "_d0"], g[tag + "_gp"], register=["start", "goal"], region=region,
                             agent_height=H, agent_width=W, net_height=48)
        errs.append(e)
        pix.append(d)
    return np.stack(errs, 0), np.stack(pix, 0)


@pytest.mark.parametrize("tag,H,W", [("reg48", 48, 64), ("reg96", 96, 128)])
def test_region_warp_errors_and_tradeoff_match_reference(g, tag, H, W):
    errs, pix = _run(g, tag, True, H, W)
    np.testing.assert_array_equal(errs, g[tag + "_errs"])
    np.testing.assert_array_equal(pix, g[tag + "_pix"])
    tr = R.registration_tradeoff(errs)
    np.testing.assert_allclose(tr, g[tag + "_tradeoff"], rtol=1e-14)
    np.testing.assert_allclose(tr.reshape(2, 2, 2).sum(axis=(0, 2)), 1.0, rtol=1e-12)     # per task, over cameras x registrations


def test_pointwise_pixels_match_reference_and_errors_are_repaired(g):
    """register_region=False: the registered pixels are the reference's; its warp errors stay zero (dead branch), ours are the
    point-wise colour distances the dead branch computes."""
    errs, pix = _run(g, "pt48", False, 48, 64)
    np.testing.assert_array_equal(pix, g["pt48_pix"])
    assert np.all(g["pt48_errs"] == 0)
    d0, gp = g["pt48_d0"], g["pt48_gp"]
    s, ws = g["pt48_start"].astype(np.float64), g["pt48_wstart"].astype(np.float64)
    want = np.linalg.norm(s[1][d0[1, 0, 0], d0[1, 0, 1]] - ws[1][d0[1, 0, 0], d0[1, 0, 1]])
    assert errs[1, 0, 0] == want and np.all(errs > 0)


def test_controller_registers_and_weights_tasks():
    """RegisterGtruthController with an identity-flow warper: designated pixels stay at their t=0 / goal positions, the
    trade-off follows 1/err, and the foreign-predictor plugin path consumes the weights (needs no GPU: the blob predictor
    returns host arrays and the cost goes through a stub backend)."""
    from fake_predictor import BlobPredictor
    H, W = 48, 64

    def warper(cur, other):
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        pts = np.stack([xx, yy], -1)[None].astype(np.float64)
        return other * 0.9 + 0.05, None, np.tile(pts, (cur.shape[1], 1, 1, 1))

    ag = {"adim": 4, "sdim": 4, "image_height": H, "image_width": W, "gpu_id": 0}
    pp = {"predictor_class": BlobPredictor, "rejection_sampling": False, "verbose": False, "num_samples": 12,
          "minimum_selection": 4, "designated_pixel_count": 2, "goal_image_warper": warper, "iterations": 1}
    pol = R.RegisterGtruthController(ag, pp, 0, 1)

    class StubCost:
        def score_external(self, distrib, goal_pix, finalweight, weights):
            self.weights, self.goal = np.array(weights), np.array(goal_pix)
            return np.arange(distrib.shape[0], dtype=np.float64)
    pol._cost_backend = StubCost()
    pol.reset()
    rng = np.random.RandomState(0)
    images = rng.randint(0, 256, size=(2, 1, H, W, 3)).astype(np.uint8)
    goal_image = rng.rand(1, 1, H, W, 3).astype(np.float32)
    state = rng.uniform(-0.5, 0.5, size=(2, 4))
    np.random.seed(1)
    for t in range(2):
        out = pol.act(goal_image=goal_image, t=t, i_tr=0, desig_pix=np.array([[10, 20]]), goal_pix=np.array([[30, 40]]),
                      images=images[:t + 1], state=state[:t + 1])
    assert out["actions"].shape == (4,)
    np.testing.assert_array_equal(pol._desig_pix, np.array([[[10, 20], [30, 40]]], dtype=np.float64))
    np.testing.assert_array_equal(pol._cost_backend.goal, np.array([[[30, 40], [30, 40]]]))
    tr = pol.plan_stat["tradeoff"]
    np.testing.assert_allclose(tr.sum(), 1.0)
    np.testing.assert_allclose(pol._cost_backend.weights, tr.reshape(-1))
    e = pol.plan_stat["warperrs"]
    np.testing.assert_allclose(tr[0, 0] / tr[0, 1], e[0, 1] / e[0, 0])
    with pytest.raises(ValueError):
        R.RegisterGtruthController(ag, dict(pp, goal_image_warper=None), 0, 1)

"""Synthetic planner inputs of the shapes BASELINE.json names: uniform uint8 context frames
(low-pass filtered so CDNA / instance norm see image-like statistics), uniform states, designated pixel at
(H/4, W/4), goal pixel at (3H/4, 3W/4), Gaussian action sequences with the reference sampler's default
standard deviations (gaussian_sampler.py:54-57).  There is no dataset or checkpoint to load (no network)."""
import numpy as np


def synth_inputs(spec, seed=0):
    rng = np.random.default_rng(seed)
    C, H, W = spec.context_frames, spec.height, spec.width
    f = rng.integers(0, 256, size=(C, spec.ncam, H, W, 3), dtype=np.uint8).astype(np.float32)
    for _ in range(2):
        f = (f + np.roll(f, 1, 2) + np.roll(f, 1, 3) + np.roll(f, (1, 1), (2, 3))) / 4
    frames = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    states = rng.uniform(-0.5, 0.5, size=(C, max(spec.sdim, 1))).astype(np.float32)[:, :spec.sdim]
    desig = np.zeros((spec.ncam, spec.ndesig, 2))
    goal = np.zeros((spec.ncam, spec.ndesig, 2))
    for c in range(spec.ncam):
        for p in range(spec.ndesig):
            desig[c, p] = [H // 4 + 3 * p, W // 4 + 2 * c]
            goal[c, p] = [3 * H // 4 - 2 * p, 3 * W // 4 - c]
    ctx_actions = (rng.standard_normal((C - 1, spec.adim)) * 0.05).astype(np.float32)
    return dict(frames=frames, states=states, desig=desig, goal=goal, ctx_actions=ctx_actions)


def gaussian_actions(spec, M, T, seed=0):
    rng = np.random.default_rng(seed + 1000)
    std = np.array([0.05, 0.05, 0.15, np.pi / 18, 2.0])[:spec.adim]
    return (rng.standard_normal((M, T, spec.adim)) * std).astype(np.float32)

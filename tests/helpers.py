"""Shared synthetic inputs and oracle wrappers for the parity tests."""
import numpy as np

from oracle import cem as OC
from oracle.predictor import OracleMultiViewPredictor
from visual_foresight_b200 import spec as S
from visual_foresight_b200.synthetic import gaussian_actions, synth_inputs  # noqa: F401


def step_actions(spec, ctx_actions, actions):
    """per-cell-step action tensor the oracle consumes: context actions prepended (C-1), S-1 total."""
    M = actions.shape[0]
    ca = np.tile(np.asarray(ctx_actions, np.float32)[None], (M, 1, 1))
    return np.concatenate([ca, actions], axis=1)[:, :spec.seq_len - 1]


def oracle_rollout(spec, weights, inp, actions, dtype=None):
    import torch
    pred = OracleMultiViewPredictor(spec, weights, dtype or torch.float32)
    onehot = OC.switch_on_pix(inp["desig"], spec.context_frames, spec.ncam, spec.height, spec.width, spec.ndesig)
    frames = inp["frames"].astype(np.float32) / 255.0
    sa = step_actions(spec, inp["ctx_actions"], actions)
    return pred.rollout(frames, inp["states"] if spec.sdim else None, onehot, sa)


def make_weights(spec, seed=0):
    return [S.init_weights(spec, seed, v) for v in range(spec.ncam)]

"""Context slicing and chunked rollout helpers with the reference's call signatures
(``visual_mpc/video_prediction/pred_util.py``: ``get_context`` :4-13, ``rollout_predictions`` :21-48), for code that drives a
``setup_predictor``-style ``predictor_func`` (e.g. the reference's ``GoalImController``, goal_im_controller.py:76-81).

The engine itself takes any sample count up to its capacity in one call, so inside this package the chunking only matters
when a caller asks for more rollouts than the handle was created for (``B200VPredEvaluation.__call__``)."""
from __future__ import annotations

import numpy as np


def get_context(n_context, t, state, images, hp=None):
    """Last ``n_context`` frames as float32 in [0, 1] and states, each with a leading batch axis of 1; ``hp.state_append``
    (a constant tail, e.g. the Sawyer configs' [0.41, 0.4, 0.184]) is tiled onto every context state."""
    lo, hi = t - n_context + 1, t + 1
    frames = np.asarray(images)[lo:hi].astype(np.float32, copy=False) / 255.
    states = np.asarray(state)[lo:hi][None]
    if hp is not None and getattr(hp, "state_append", None):
        tail = np.tile(np.asarray(hp.state_append).reshape(1, 1, -1), (1, n_context, 1))
        states = np.concatenate((states, tail), axis=-1)
    return frames[None], states


def rollout_predictions(predictor, b_size, actions, context_frames, context_states=None, input_distribs=None, logger=None):
    """Runs ``predictor`` over ``actions`` in chunks of ``b_size`` (the last chunk zero-padded to ``b_size``, its outputs cut
    back) and returns three lists (frames, distributions, states) with one entry per chunk — concatenate along axis 0."""
    actions = np.asarray(actions)
    total = actions.shape[0]
    out = ([], [], [])
    for start in range(0, max(total, 1), b_size):
        chunk = actions[start:start + b_size]
        n = chunk.shape[0]
        if start + b_size >= total and n < b_size:            # only the final chunk is ever short
            padded = np.zeros((b_size,) + chunk.shape[1:])
            padded[:n] = chunk
            chunk = padded
        if logger:
            logger.log("Vpred run: {} with {} actions".format(start // b_size, n))
        res = predictor(input_images=context_frames, input_state=context_states, input_actions=chunk,
                        input_one_hot_images=input_distribs)
        for lst, arr in zip(out, res):
            lst.append(None if arr is None else arr[:n])
    return out

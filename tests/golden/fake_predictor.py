"""Deterministic stand-in predictor with the ``predictor_class`` contract of the reference
(``pixel_cost_controller.py:29-36,83-84,175``).  It is shared by ``make_golden.py`` (where it is
injected into the REFERENCE's PixelCostController) and by the tests (where it is injected into this
repo's controller), so both sides see bit-identical "predictions".  Test infrastructure only."""
import numpy as np


class BlobPredictor(object):
    """Moves a Gaussian blob by the cumulative xy action; frames are a smooth function of the blob."""
    n_context = 2
    sequence_length = 15
    n_cam = 1
    calls = None

    def __init__(self, model_path, hparams, n_gpus=1, first_gpu=0):
        self.hparams = dict(hparams)
        self.n_gpus, self.first_gpu = n_gpus, first_gpu
        self.restored = False
        self.calls = []

    def restore(self):
        self.restored = True

    def __call__(self, context, inputs):
        actions = np.asarray(inputs["actions"], dtype=np.float64)
        frames = context["context_frames"]
        distrib = np.asarray(context["context_pixel_distributions"], dtype=np.float64)
        self.calls.append({k: (None if v is None else np.array(v)) for k, v in context.items()})
        self.calls[-1]["actions"] = actions.copy()
        H, W = frames.shape[2], frames.shape[3]
        nd = distrib.shape[-1]
        ncam = distrib.shape[1]
        P = self.sequence_length - self.n_context
        M = actions.shape[0]
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        out_d = np.zeros((M, P, ncam, H, W, nd), dtype=np.float32)
        out_f = np.zeros((M, P, ncam, H, W, 3), dtype=np.float32)
        for c in range(ncam):
            for p in range(nd):
                d0 = distrib[-1, c, :, :, p]
                tot = d0.sum()
                cy = (d0 * yy).sum() / tot
                cx = (d0 * xx).sum() / tot
                pos = np.cumsum(actions[:, :P, :2], axis=1) * 40.0          # (M,P,2)
                py = cy + pos[:, :, 0] + 1.5 * p
                px = cx + pos[:, :, 1] - 0.5 * c
                g = np.exp(-0.5 * ((yy[None, None] - py[:, :, None, None]) ** 2 +
                                   (xx[None, None] - px[:, :, None, None]) ** 2) / 4.0) + 1e-6
                out_d[:, :, c, :, :, p] = (0.7 * g).astype(np.float32)      # deliberately un-normalised
                out_f[:, :, c] = np.clip(g[..., None] * np.array([0.9, 0.5, 0.2]), 0, 1).astype(np.float32)
        return {"predicted_frames": out_f, "predicted_pixel_distributions": out_d}

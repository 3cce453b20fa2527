"""Import shim that makes the reference's CEM/cost/sampler code importable in THIS container
(SURVEY.md Appendix B).  Only used by ``make_golden.py`` and by tests that are skipped when
/root/reference is absent (it does not exist on the GPU box).  Nothing here is product code."""
import inspect
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("VF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "visual_mpc"))


class _HParams(object):
    """Minimal stand-in for tensorflow.contrib.training.HParams (reference policy/policy.py:4,51-63)."""

    def __init__(self, **kw):
        object.__setattr__(self, "_d", {})
        for k, v in kw.items():
            self.add_hparam(k, v)

    def add_hparam(self, name, value):
        if name in self._d:
            raise ValueError("Hyperparameter name is reserved: %s" % name)
        self._d[name] = value

    def set_hparam(self, name, value):
        if name not in self._d:
            raise KeyError(name)
        self._d[name] = value

    def get(self, key, default=None):
        return self._d.get(key, default)

    def values(self):
        return dict(self._d)

    def __contains__(self, key):
        return key in self._d

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_d")[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self._d[name] = value


def install():
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if getattr(install, "_done", False):
        return
    if not hasattr(np, "int"):
        np.int = int            # removed alias used at pixel_cost_controller.py:208

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    tf = mod("tensorflow")
    contrib = mod("tensorflow.contrib")
    training = mod("tensorflow.contrib.training", HParams=_HParams)
    tf.contrib = contrib
    contrib.training = training
    mod("funcsigs", signature=inspect.signature, Parameter=inspect.Parameter)
    mod("imp")
    mpl = mod("matplotlib")
    plt = mod("matplotlib.pyplot")
    mpl.pyplot = plt
    rn = mod("robonet")
    rv = mod("robonet.video_prediction")
    rt = mod("robonet.video_prediction.testing", VPredEvaluation=None)
    rn.video_prediction = rv
    rv.testing = rt
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            mod("cv2")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install._done = True

"""Generates tests/golden/oracle_predictor_golden.npz: seeded outputs of oracle/predictor.py (spec P) on a
small configuration.  These pin the ORACLE against regressions and give the GPU tests a fixture that does
not need torch-CPU time.  NOTE: produced by the build's own oracle — TF1 parity is unpinned.

    python tests/golden/make_predictor_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from visual_foresight_b200 import spec as S  # noqa: E402
import helpers as Hh  # noqa: E402

CASES = {
    "a": dict(height=32, width=32, seq_len=5, ncam=1, ndesig=1, adim=4, sdim=4, M=3),
    "b": dict(height=48, width=64, seq_len=4, ncam=2, ndesig=2, adim=4, sdim=5, M=2),
}


def main():
    out = {}
    for name, c in CASES.items():
        c = dict(c)
        M = c.pop("M")
        sp = S.spec_64(**c)
        w = Hh.make_weights(sp, seed=3)
        inp = Hh.synth_inputs(sp, seed=5)
        acts = Hh.gaussian_actions(sp, M, sp.seq_len - sp.context_frames + 2, seed=7)
        gi, gd, gs = Hh.oracle_rollout(sp, w, inp, acts)
        out[name + "_frames"], out[name + "_distrib"], out[name + "_states"] = gi, gd, gs
    p = os.path.join(HERE, "oracle_predictor_golden.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p))


if __name__ == "__main__":
    main()

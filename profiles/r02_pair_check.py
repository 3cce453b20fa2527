"""CTA-pair (cta_group::2) gate convolution: correctness vs float64 and timing vs the single-CTA kernel.
usage: VF_CTA_PAIR=1 python profiles/r02_pair_check.py   (and once with VF_CTA_PAIR=0 for the reference timing)"""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visual_foresight_b200 import spec as S
from visual_foresight_b200.engine import Engine

e = Engine(S.spec_64(height=32, width=32, seq_len=4), 4, precision="f16x3")
rng = np.random.default_rng(0)
for (B, H, W, Cin, Cout, k) in [(2, 16, 16, 128, 256, 5), (5, 16, 16, 64, 512, 5), (3, 16, 16, 128, 256, 3)]:
    x = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    y = e.debug_conv2d(x, w, b, impl=1)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).double().permute(0, 3, 1, 2), torch.from_numpy(w).double().permute(3, 2, 0, 1),
                                     torch.from_numpy(b).double(), padding=k // 2).permute(0, 2, 3, 1).numpy()
    err = np.abs(y - ref).max()
    print("shape", (B, H, W, Cin, Cout, k), "max abs err vs float64: %.3g" % err, "OK" if err < 1e-4 else "FAIL", flush=True)
for (B, H, W, Cin, Cout, k) in [(200, 16, 16, 128, 256, 5), (512, 16, 16, 128, 256, 5)]:
    ms = e.debug_conv_time(B, H, W, Cin, Cout, k, impl=1, reps=20)
    print("time", (B, H, W, Cin, Cout, k), "%.1f us per launch  (VF_CTA_PAIR=%s)" % (ms * 1e3, os.environ.get("VF_CTA_PAIR", "0")), flush=True)

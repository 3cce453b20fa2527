"""Per-layer timing of every convolution of one cell step, each launched alone and back to back (vf_debug_conv_time: CUDA
events on the engine's stream, data resident in HBM) — a fast A/B loop for conv_mma.cu changes that needs no ncu:

    python profiles/conv_microbench.py [--samples 200] [--size 64] [--reps 20] [--precision f16x3]
    VF_ROWGROUPS=0 python profiles/conv_microbench.py          # any engine switch applies

Shapes are the EXECUTED ones (action/state channels folded into the bias, first conv as the 3x3 block conv); TFLOP/s counts
2*MAC of the executed shape times the number of MMA passes.  Back-to-back launches of one layer keep its weights and (for the
small maps) activations in L2, so these times are a few percent below the in-pipeline ones of profiles/launches_*.csv."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def layers(size):
    s = size
    return [  # name, H, W, Cin, Cout, k
        ("enc0 (block conv)", s // 2, s // 2, 32, 32, 3),
        ("lstm0", s // 2, s // 2, 64, 128, 5),
        ("enc1", s // 2, s // 2, 32, 64, 3),
        ("lstm1", s // 4, s // 4, 128, 256, 5),
        ("enc2", s // 4, s // 4, 64, 128, 3),
        ("lstm2", s // 8, s // 8, 256, 512, 5),
        ("dec0", s // 4, s // 4, 128, 64, 3),
        ("lstm3", s // 4, s // 4, 128, 256, 5),
        ("dec1", s // 2, s // 2, 128, 32, 3),
        ("lstm4", s // 2, s // 2, 64, 128, 5),
        ("dec2", s, s, 64, 32, 3),
        ("scratch0 / masks0", s, s, 32, 32, 3),
        ("scratch1", s, s, 32, 3, 3),
        ("masks1", s, s, 56, 7, 3),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--precision", default="f16x3", choices=["fp32_simt", "f16x3", "f16x1"])
    a = ap.parse_args()
    from visual_foresight_b200 import spec as S
    from visual_foresight_b200.engine import PRECISIONS, Engine
    impl = PRECISIONS[a.precision]
    passes = {"fp32_simt": 1, "f16x3": 3, "f16x1": 1}[a.precision]
    e = Engine(S.spec_64(height=32, width=32, seq_len=4), 2, precision=a.precision)     # a small handle: the tool allocates per call
    total = 0.0
    print("%-20s %9s %6s %6s %3s %10s %12s" % ("layer", "HxW", "Cin", "Cout", "k", "us/launch", "TFLOP/s raw"))
    for name, H, W, cin, cout, k in layers(a.size):
        ms = e.debug_conv_time(a.samples, H, W, cin, cout, k, impl=impl, reps=a.reps)
        fl = 2.0 * a.samples * H * W * k * k * cin * cout * passes
        total += ms
        print("%-20s %4dx%-4d %6d %6d %3d %10.1f %12.1f" % (name, H, W, cin, cout, k, ms * 1e3, fl / (ms * 1e-3) / 1e12))
    print("sum of one cell step's convolutions (scratch0 and masks0 counted once each): %.1f us" %
          ((total + e.debug_conv_time(a.samples, a.size, a.size, 32, 32, 3, impl=impl, reps=a.reps)) * 1e3))
    e.close()


if __name__ == "__main__":
    main()

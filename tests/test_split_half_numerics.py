"""Numerics model of the tensor-core path (DESIGN 4.2 "fp32-grade numerics on fp16 tensor cores"), emulated in NumPy: both
operands are split x = hi + lo into two fp16 values (weights pre-scaled by a power of two so that lo stays normal), and
hi*hi + hi*lo + lo*hi accumulate in fp32.  fp16 x fp16 products are exact in fp32, so a float32 matmul of the fp16 parts is
a faithful model up to summation order.  Checked against float64: the three-pass form is fp32-grade at the largest
reduction length of the predictor (K = 25 taps x 256 channels = 6400), the single-pass form is not."""
import numpy as np


def split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def pow2_scale(w):
    s = int(np.floor(np.log2(16384.0 / np.abs(w).max())))           # max |w| * 2^s in [8192, 16384] (conv_mma.cu prepare_weights)
    return np.float32(2.0 ** s), s


def test_three_pass_split_is_fp32_grade_and_one_pass_is_not():
    rng = np.random.default_rng(0)
    K, N, P = 6400, 64, 96
    x = rng.standard_normal((P, K)).astype(np.float32) * np.float32(0.7)       # normalised activations, O(1)
    w = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)           # fan-in scaled weights
    ref = x.astype(np.float64) @ w.astype(np.float64)
    sc, s = pow2_scale(w)
    xh, xl = split(x)
    wh, wl = split(w * sc)
    assert (np.abs(wl) < 2.0 ** -14).mean() < 1e-3                             # lo parts of the scaled weights are normal fp16 numbers
    assert (np.abs(split(w)[1]) < 2.0 ** -14).mean() > 0.2                     # ... which they would often not be without the scale
    three = (xh @ wh + xh @ wl + xl @ wh) * np.float32(2.0 ** -s)
    one = (xh @ wh) * np.float32(2.0 ** -s)
    plain32 = x @ w
    e3, e1, e32 = (np.abs(v - ref).max() for v in (three, one, plain32))
    assert e3 <= 1e-4 / 20                                                       # well inside the 1e-4 frame contract per conv
    assert e3 <= 4 * e32 + 1e-7                                                  # same class as a float32 matmul
    assert e1 >= 20 * e3 and e1 > 3e-4                                           # fp16 inputs alone: ~2^-11 relative per operand
    # the dropped lo*lo term is below fp32 resolution of the result
    assert np.abs(xl @ wl).max() * 2.0 ** -s < 1e-6


def test_split_reconstructs_to_22_bits():
    """hi + lo carries 22 significant bits of x; below |x| ~ 2^-3 the lo part enters the fp16 subnormal range and the error
    floor is its spacing 2^-24 (6e-8 absolute — irrelevant for O(1) activations)."""
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(100000) * 3).astype(np.float32)
    hi, lo = split(x)
    err = np.abs((hi + lo).astype(np.float64) - x)
    assert np.all(err <= 2.0 ** -22 * np.abs(x) + 2.0 ** -24)

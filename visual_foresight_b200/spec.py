"""Predictor specification ("spec P") as data: layer table, weight names/shapes, seeded
initialisation and the algorithmic FLOP / byte counters the roofline is computed from.

The reference does not contain the predictor arithmetic (it lives in the
un-vendored ``video_prediction`` package, call sites
``visual_mpc/video_prediction/vpred_model_interface.py:52-88``).  This module is the canonical
definition this build owns; both the CUDA engine and ``oracle/predictor.py`` consume it.

Tensor contract at the boundary follows ``visual_mpc/video_prediction/setup_predictor.py:98-114``
(images ``(1, C, ncam, H, W, 3)``, actions ``(M, S, adim)``, states ``(1, C, sdim)``, pixel
distributions ``(1, C, ncam, H, W, ndesig)``) and ``vpred_model_interface.py:75-88`` (outputs
``(M, P, ncam, H, W, ·)``).
"""
from __future__ import annotations

import dataclasses
import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np

MAX_LAYERS = 8


@dataclasses.dataclass(frozen=True)
class PredictorSpec:
    height: int = 64
    width: int = 64
    ncam: int = 1
    ndesig: int = 1
    adim: int = 4
    sdim: int = 4               # state dims fed to the net (0 = use_state False)
    nz: int = 0
    rnn_z: bool = False         # use_rnn_z: the latent passes through a dense LSTM(nz) before being tiled into the convs
    seq_len: int = 15           # S  (BASELINE "H")
    context_frames: int = 2     # C
    ngf: int = 32
    # (out_channels, has_rnn) per layer; encoder layer 0 is 5x5, the rest 3x3 (conv_pool2d)
    encoder: Tuple[Tuple[int, bool], ...] = ((32, True), (64, True), (128, True))
    # decoder layers are upsample_conv2d 3x3
    decoder: Tuple[Tuple[int, bool], ...] = ((64, True), (32, True), (32, False))
    num_transformed: int = 4    # CDNA kernels
    cdna_ksize: int = 5
    lstm_ksize: int = 5
    norm_eps: float = 1e-6
    forget_bias: float = 1.0

    # ---- derived -------------------------------------------------------------------------
    @property
    def n_steps(self) -> int:           # conv-RNN cell steps
        return self.seq_len - 1

    @property
    def n_pred(self) -> int:            # predicted frames P = S - C
        return self.seq_len - self.context_frames

    @property
    def sa_dim(self) -> int:            # A: spatially-constant vector tiled into every conv
        return self.adim + self.sdim + self.nz

    @property
    def n_masks(self) -> int:           # transformed images + prev image + first image + scratch
        return self.num_transformed + 3

    def enc_hw(self, i: int) -> Tuple[int, int]:
        """resolution of encoder layer i OUTPUT (after the 2x2 pool)."""
        return self.height >> (i + 1), self.width >> (i + 1)

    def dec_hw(self, i: int) -> Tuple[int, int]:
        """resolution of decoder layer i OUTPUT (after the x2 upsample)."""
        n = len(self.encoder)
        return self.height >> (n - 1 - i), self.width >> (n - 1 - i)

    def validate(self) -> None:
        n = len(self.encoder)
        assert 1 <= n <= MAX_LAYERS and len(self.decoder) == n, "encoder/decoder depth mismatch"
        assert self.height % (1 << n) == 0 and self.width % (1 << n) == 0
        assert self.decoder[-1][0] == self.ngf
        assert self.context_frames >= 1 and self.seq_len > self.context_frames
        assert self.cdna_ksize % 2 == 1 and self.lstm_ksize % 2 == 1
        assert not self.rnn_z or self.nz > 0, "rnn_z needs nz > 0"


def spec_64(**kw) -> PredictorSpec:
    """64-px family (c1, c2, c3, c4): enc [(32,rnn),(64,rnn),(128,rnn)], dec [(64,rnn),(32,rnn),(32,-)]."""
    return PredictorSpec(**kw)


def spec_128(**kw) -> PredictorSpec:
    """128-px family (c5): enc [(32,-),(64,rnn),(128,rnn),(256,rnn)], dec [(256,rnn),(128,rnn),(64,-),(32,-)]."""
    kw.setdefault("height", 128)
    kw.setdefault("width", 128)
    return PredictorSpec(
        encoder=((32, False), (64, True), (128, True), (256, True)),
        decoder=((256, True), (128, True), (64, False), (32, False)),
        **kw,
    )


# --------------------------------------------------------------------------------------------
# layer table: every convolution of one cell step, in execution order
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class ConvDesc:
    name: str          # weight prefix
    kind: str          # 'enc' | 'dec' | 'lstm' | 'head'
    ksize: int
    hw: Tuple[int, int]        # resolution the conv RUNS at (before pool / after upsample)
    cin_spatial: int   # channels that vary over space
    cin_const: int     # tiled action/state/z channels (sa)
    cout: int
    bias: bool
    norm: bool


def conv_table(spec: PredictorSpec) -> List[ConvDesc]:
    spec.validate()
    A = spec.sa_dim
    t: List[ConvDesc] = []
    n = len(spec.encoder)
    c_prev = 2 * 3                       # concat(image, first)
    hw = (spec.height, spec.width)
    enc_out: List[int] = []
    for i, (oc, rnn) in enumerate(spec.encoder):
        k = 5 if i == 0 else 3
        t.append(ConvDesc(f"enc{i}.conv", "enc", k, hw, c_prev, A, oc, True, True))
        hw = (hw[0] // 2, hw[1] // 2)
        if rnn:
            t.append(ConvDesc(f"enc{i}.lstm", "lstm", spec.lstm_ksize, hw, oc + oc, A, 4 * oc, False, True))
        enc_out.append(oc)
        c_prev = oc
    for i, (oc, rnn) in enumerate(spec.decoder):
        cin = c_prev + (enc_out[n - 1 - i] if i > 0 else 0)
        hw = (hw[0] * 2, hw[1] * 2)
        t.append(ConvDesc(f"dec{i}.conv", "dec", 3, hw, cin, A, oc, True, True))
        if rnn:
            t.append(ConvDesc(f"dec{i}.lstm", "lstm", spec.lstm_ksize, hw, oc + oc, A, 4 * oc, False, True))
        c_prev = oc
    g = spec.ngf
    full = (spec.height, spec.width)
    t.append(ConvDesc("scratch.conv0", "head", 3, full, g, 0, g, True, True))
    t.append(ConvDesc("scratch.conv1", "head", 3, full, g, 0, 3, True, False))
    t.append(ConvDesc("masks.conv0", "head", 3, full, g, 0, g, True, True))
    t.append(ConvDesc("masks.conv1", "head", 3, full, g + 3 * spec.n_masks, 0, spec.n_masks, True, False))
    return t


def cdna_feature_dim(spec: PredictorSpec) -> int:
    """flatten(smallest encoder rnn output) in NHWC order."""
    n = len(spec.encoder)
    h, w = spec.enc_hw(n - 1)
    return h * w * spec.encoder[-1][0]


def weight_shapes(spec: PredictorSpec) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape for ONE view.  Conv weights are HWIO (kh, kw, cin_total, cout) with the input
    channel order [spatial channels..., sa channels (action, state, z)] and, for LSTM cells,
    [x, sa, h_prev] (P2/P3 of SURVEY.md 8a)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for d in conv_table(spec):
        cin = d.cin_spatial + d.cin_const
        s[d.name + ".w"] = (d.ksize, d.ksize, cin, d.cout)
        if d.bias:
            s[d.name + ".b"] = (d.cout,)
        if d.kind == "lstm":
            f = d.cout // 4
            s[d.name + ".gates_gamma"] = (d.cout,)
            s[d.name + ".gates_beta"] = (d.cout,)
            s[d.name + ".cell_gamma"] = (f,)
            s[d.name + ".cell_beta"] = (f,)
        elif d.norm:
            s[d.name + ".gamma"] = (d.cout,)
            s[d.name + ".beta"] = (d.cout,)
    kk = spec.cdna_ksize * spec.cdna_ksize * spec.num_transformed
    s["cdna.dense.w"] = (cdna_feature_dim(spec), kk)
    s["cdna.dense.b"] = (kk,)
    if spec.sdim > 0:
        s["state.dense.w"] = (spec.adim + spec.sdim, spec.sdim)
        s["state.dense.b"] = (spec.sdim,)
    if spec.rnn_z:                       # BasicLSTMCell(nz): kernel [(z, h), 4*nz] with gate order i, j, f, o
        s["zrnn.w"] = (2 * spec.nz, 4 * spec.nz)
        s["zrnn.b"] = (4 * spec.nz,)
    return s


def init_weights(spec: PredictorSpec, seed: int = 0, view: int = 0, affine_jitter: float = 0.1) -> Dict[str, np.ndarray]:
    """Seeded fan-in-scaled normal weights (activations stay O(1)); gamma = 1 + jitter*N, beta = jitter*N
    so the affine paths are exercised.  Deterministic across platforms (numpy PCG64)."""
    rng = np.random.Generator(np.random.PCG64([seed, view, 0x5EED]))
    out: Dict[str, np.ndarray] = {}
    for name, shp in weight_shapes(spec).items():
        if name.endswith(".w"):
            fan_in = int(np.prod(shp[:-1]))
            w = rng.standard_normal(shp) / math.sqrt(fan_in)
        elif name.endswith("gamma"):
            w = 1.0 + affine_jitter * rng.standard_normal(shp)
        elif name.endswith("beta") or name.endswith(".b"):
            w = affine_jitter * rng.standard_normal(shp)
        else:
            raise AssertionError(name)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


# --------------------------------------------------------------------------------------------
# roofline counters
# --------------------------------------------------------------------------------------------
def flops_per_sample_step(spec: PredictorSpec, executed: bool = False) -> Dict[str, float]:
    """Algorithmic FLOPs (2*MAC) per sample per cell step, per view.

    ``executed=False`` counts the tiled sa channels as convolution input channels (the way the
    reference graph computes them).  ``executed=True`` counts what the engine
    actually issues: sa channels folded into a per-sample border-class bias."""
    out: Dict[str, float] = {}
    for d in conv_table(spec):
        cin = d.cin_spatial + (0 if executed else d.cin_const)
        out[d.name] = 2.0 * d.hw[0] * d.hw[1] * d.ksize * d.ksize * cin * d.cout
    kk = spec.cdna_ksize ** 2
    out["cdna.dense"] = 2.0 * cdna_feature_dim(spec) * kk * spec.num_transformed
    px = spec.height * spec.width
    out["cdna.apply"] = 2.0 * px * kk * spec.num_transformed * (3 + spec.ndesig)
    out["composite"] = 2.0 * px * spec.n_masks * (3 + spec.ndesig)
    out["total"] = float(sum(out.values()))
    out["conv_lstm"] = float(sum(v for k, v in out.items() if k.endswith(".lstm")))
    return out


def flops_per_plan(spec: PredictorSpec, num_samples: int, iterations: int, executed: bool = False) -> float:
    return flops_per_sample_step(spec, executed)["total"] * spec.n_steps * num_samples * spec.ncam * iterations


def composite_bytes_per_sample_step(spec: PredictorSpec) -> float:
    """Algorithmic HBM bytes of the fused CDNA-apply + mask composite + cost kernel:
    read prev image+distrib, first image+distrib, scratch, mask logits, kernels; write gen image+distrib."""
    px = spec.height * spec.width
    ch = 3 + spec.ndesig
    rd = px * 4 * (ch + ch + 3 + spec.n_masks) + 4 * spec.cdna_ksize ** 2 * spec.num_transformed
    wr = px * 4 * ch
    return float(rd + wr)

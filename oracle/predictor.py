"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the action-conditioned
CDNA / conv-LSTM video-prediction forward pass ("spec P").

PARITY UNPINNED vs TF1: the arithmetic the reference drives lives in the un-vendored package
``video_prediction`` (febert/video_prediction-1 @ branch ``dev``, no commit pin;
reference ``README.md:17``, call sites ``visual_mpc/video_prediction/vpred_model_interface.py:52-88``).
No source, hparams JSON, checkpoint or golden vector for it exists under /root/reference, and
TensorFlow is not installable here, so this file restates the published SAVP/CDNA generator
algorithm from the layer table in ``visual_foresight_b200/spec.py``.  Every parity claim made
against this oracle reads "vs build oracle (spec P); TF1 parity unpinned".

I/O contract follows the reference:
  * inputs  — ``setup_predictor.py:98-114`` / ``pixel_cost_controller.py:77-83``
  * outputs — ``vpred_model_interface.py:75-88`` (gen_images (M,P,ncam,H,W,3), gen_distrib
    (M,P,ncam,H,W,ndesig), gen_states (M,P,sdim))

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from visual_foresight_b200.spec import PredictorSpec


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def _conv_same(x, w_hwio, bias=None):
    """x NCHW, w HWIO -> SAME zero-padded stride-1 cross-correlation (TF conv2d semantics)."""
    w = w_hwio.permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, w, bias, padding=w.shape[-1] // 2)


def _instance_norm(x, gamma, beta, eps):
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)      # biased
    return (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def _tile_concat(x, sa):
    if sa.shape[1] == 0:
        return x
    b, _, h, w = x.shape
    return torch.cat([x, sa.view(b, -1, 1, 1).expand(b, sa.shape[1], h, w)], dim=1)


def _pad_symmetric(x, p):
    """TF 'SYMMETRIC' padding (edge pixel repeated: [b a | a b c d | d c]); NOT torch 'reflect'."""
    top = x[:, :, :p].flip(2)
    bot = x[:, :, -p:].flip(2)
    x = torch.cat([top, x, bot], dim=2)
    left = x[:, :, :, :p].flip(3)
    right = x[:, :, :, -p:].flip(3)
    return torch.cat([left, x, right], dim=3)


def _cdna_apply(img, kern, ksize):
    """img (B,C,H,W); kern (B,N,k,k) per-sample kernels; same kernel for every channel.
    returns list of N tensors (B,C,H,W):  T_n[y,x,c] = sum_{u,v} pad(img)[y+u, x+v, c] * k[u,v,n]  (P6)."""
    b, c, h, w = img.shape
    n = kern.shape[1]
    p = ksize // 2
    xp = _pad_symmetric(img, p)                                     # (B,C,H+2p,W+2p)
    xp = xp.reshape(1, b * c, h + 2 * p, w + 2 * p)
    wt = kern.view(b, 1, n, ksize, ksize).expand(b, c, n, ksize, ksize).reshape(b * c * n, 1, ksize, ksize)
    out = F.conv2d(xp, wt, groups=b * c)                            # (1, B*C*N, H, W)
    out = out.view(b, c, n, h, w)
    return [out[:, :, i] for i in range(n)]


class OraclePredictor:
    """One view of the predictor.  ``weights``: name -> ndarray as produced by spec.init_weights."""

    def __init__(self, spec: PredictorSpec, weights: Dict[str, np.ndarray], dtype=torch.float32):
        spec.validate()
        self.spec = spec
        self.dtype = dtype
        self.w = {k: _t(v, dtype) for k, v in weights.items()}

    # -- building blocks ---------------------------------------------------------------------
    def _lstm(self, name, x, sa, state):
        spec = self.spec
        w = self.w
        f = w[name + ".cell_gamma"].shape[0]
        if state is None:
            b, _, h, wd = x.shape
            c = torch.zeros(b, f, h, wd, dtype=self.dtype)
            hprev = torch.zeros(b, f, h, wd, dtype=self.dtype)
        else:
            c, hprev = state
        inp = torch.cat([_tile_concat(x, sa), hprev], dim=1)
        g = _conv_same(inp, w[name + ".w"])
        g = _instance_norm(g, w[name + ".gates_gamma"], w[name + ".gates_beta"], spec.norm_eps)
        i, j, fg, o = torch.split(g, f, dim=1)
        c_new = c * torch.sigmoid(fg + spec.forget_bias) + torch.sigmoid(i) * torch.tanh(j)
        c_new = _instance_norm(c_new, w[name + ".cell_gamma"], w[name + ".cell_beta"], spec.norm_eps)
        h_new = torch.tanh(c_new) * torch.sigmoid(o)
        return h_new, (c_new, h_new)

    def _conv_norm_relu(self, name, x):
        w = self.w
        y = _conv_same(x, w[name + ".w"], w[name + ".b"])
        y = _instance_norm(y, w[name + ".gamma"], w[name + ".beta"], self.spec.norm_eps)
        return F.relu(y)

    def cell(self, image, distrib, first, first_distrib, sa, states, debug: Optional[dict] = None):
        """One cell step (P1-P9).  image/first (B,3,H,W); distrib/first_distrib (B,nd,H,W) or None."""
        spec, w = self.spec, self.w
        n = len(spec.encoder)
        new_states: List = [None] * (2 * n)
        x = torch.cat([image, first], dim=1)
        enc_out = []
        for i, (oc, rnn) in enumerate(spec.encoder):
            name = f"enc{i}.conv"
            y = _conv_same(_tile_concat(x, sa), w[name + ".w"], w[name + ".b"])
            y = F.avg_pool2d(y, 2)
            y = F.relu(_instance_norm(y, w[name + ".gamma"], w[name + ".beta"], spec.norm_eps))
            if debug is not None:
                debug[name] = y
            if rnn:
                y, new_states[i] = self._lstm(f"enc{i}.lstm", y, sa, states[i])
                if debug is not None:
                    debug[f"enc{i}.lstm.h"] = y
                    debug[f"enc{i}.lstm.c"] = new_states[i][0]
            enc_out.append(y)
            x = y
        smallest = x
        for i, (oc, rnn) in enumerate(spec.decoder):
            name = f"dec{i}.conv"
            if i > 0:
                x = torch.cat([x, enc_out[n - 1 - i]], dim=1)
            y = _tile_concat(x, sa)
            y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=False)
            y = _conv_same(y, w[name + ".w"], w[name + ".b"])
            y = F.relu(_instance_norm(y, w[name + ".gamma"], w[name + ".beta"], spec.norm_eps))
            if debug is not None:
                debug[name] = y
            if rnn:
                y, new_states[n + i] = self._lstm(f"dec{i}.lstm", y, sa, states[n + i])
                if debug is not None:
                    debug[f"dec{i}.lstm.h"] = y
            x = y
        h_last = x
        b = image.shape[0]
        # P5 CDNA kernel head
        k = spec.cdna_ksize
        nt = spec.num_transformed
        feat = smallest.permute(0, 2, 3, 1).reshape(b, -1)                       # NHWC flatten
        kern = feat @ w["cdna.dense.w"] + w["cdna.dense.b"]
        kern = kern.view(b, k, k, nt)
        ident = torch.zeros(k, k, 1, dtype=self.dtype)
        ident[k // 2, k // 2, 0] = 1.0
        kern = kern + ident
        kern = F.relu(kern - 1e-12) + 1e-12
        kern = kern / kern.sum(dim=(1, 2), keepdim=True)
        kern = kern.permute(0, 3, 1, 2).contiguous()                             # (B,N,k,k)
        if debug is not None:
            debug["cdna.kernels"] = kern
        # P6
        t_img = _cdna_apply(image, kern, k)
        # P7 scratch
        s = self._conv_norm_relu("scratch.conv0", h_last)
        s = torch.sigmoid(_conv_same(s, w["scratch.conv1.w"], w["scratch.conv1.b"]))
        # P8 masks + composite
        layers = t_img + [image, first, s]
        hm = self._conv_norm_relu("masks.conv0", h_last)
        logits = _conv_same(torch.cat([hm] + layers, dim=1), w["masks.conv1.w"], w["masks.conv1.b"])
        masks = torch.softmax(logits, dim=1)
        if debug is not None:
            debug["scratch"] = s
            debug["mask_logits"] = logits
        gen_image = sum(masks[:, i:i + 1] * layers[i] for i in range(len(layers)))
        gen_distrib = None
        if distrib is not None:
            t_d = _cdna_apply(distrib, kern, k)
            layers_d = t_d + [distrib, first_distrib, distrib]
            gen_distrib = sum(masks[:, i:i + 1] * layers_d[i] for i in range(len(layers_d)))
            if debug is not None:
                debug["gen_distrib_raw"] = gen_distrib
            gen_distrib = gen_distrib / gen_distrib.sum(dim=(2, 3), keepdim=True)
        return gen_image, gen_distrib, new_states

    # -- rollout -----------------------------------------------------------------------------
    @torch.no_grad()
    def rollout(self, ctx_frames, ctx_states, ctx_distrib, step_actions, zs=None, debug_steps=None):
        """ctx_frames (C,H,W,3) float in [0,1]; ctx_states (C,sdim) or None; ctx_distrib (C,H,W,nd) or None;
        step_actions (M, S-1, adim) — one action per cell step (context actions already prepended).
        returns gen_images (M,P,H,W,3), gen_distrib (M,P,H,W,nd)|None, gen_states (M,P,sdim)|None."""
        spec = self.spec
        C, S = spec.context_frames, spec.seq_len
        acts = _t(step_actions, self.dtype)
        M = acts.shape[0]
        assert acts.shape[1] >= S - 1, "need one action per cell step"
        frames = _t(ctx_frames, self.dtype).permute(0, 3, 1, 2)                  # (C,3,H,W)
        dist = None if ctx_distrib is None else _t(ctx_distrib, self.dtype).permute(0, 3, 1, 2)
        st_ctx = None if (ctx_states is None or spec.sdim == 0) else _t(ctx_states, self.dtype)
        first = frames[0:1].expand(M, -1, -1, -1)
        first_d = None if dist is None else dist[0:1].expand(M, -1, -1, -1)
        states: List = [None] * (2 * len(spec.encoder))
        gen_image = gen_distrib = gen_state = None
        zc = zh = torch.zeros(M, spec.nz, dtype=self.dtype) if spec.rnn_z else None   # dense LSTM over the latent (use_rnn_z)
        out_i, out_d, out_s = [], [], []
        for tau in range(S - 1):
            if tau < C:
                image = frames[tau:tau + 1].expand(M, -1, -1, -1)
                distrib = None if dist is None else dist[tau:tau + 1].expand(M, -1, -1, -1)
                state = None if st_ctx is None else st_ctx[tau:tau + 1].expand(M, -1)
            else:
                image, distrib, state = gen_image, gen_distrib, gen_state
            parts = [acts[:, tau]]
            if state is not None:
                parts.append(state)
            if spec.nz > 0:
                z = _t(zs[:, tau], self.dtype)
                if spec.rnn_z:                                                       # BasicLSTMCell, forget_bias 1.0, gates i, j, f, o
                    g = torch.cat([z, zh], dim=1) @ self.w["zrnn.w"] + self.w["zrnn.b"]
                    gi_, gj_, gf_, go_ = torch.chunk(g, 4, dim=1)
                    zc = zc * torch.sigmoid(gf_ + 1.0) + torch.sigmoid(gi_) * torch.tanh(gj_)
                    zh = torch.tanh(zc) * torch.sigmoid(go_)
                    z = zh
                parts.append(z)
            sa = torch.cat(parts, dim=1)
            dbg = None
            if debug_steps is not None and tau in debug_steps:
                dbg = debug_steps[tau]
            gen_image, gen_distrib, states = self.cell(image, distrib, first, first_d, sa, states, dbg)
            if state is not None:                                                  # P9
                gen_state = torch.cat([acts[:, tau], state], dim=1) @ self.w["state.dense.w"] + self.w["state.dense.b"]
            if tau >= C - 1:
                out_i.append(gen_image)
                out_d.append(gen_distrib)
                out_s.append(gen_state)
        gi = torch.stack(out_i, 1).permute(0, 1, 3, 4, 2).contiguous().numpy()
        gd = None if out_d[0] is None else torch.stack(out_d, 1).permute(0, 1, 3, 4, 2).contiguous().numpy()
        gs = None if out_s[0] is None else torch.stack(out_s, 1).numpy()
        return gi, gd, gs


class OracleMultiViewPredictor:
    """ncam independent weight sets sharing actions/states (IndepMultiSAVP; reference
    ``vpred_model_interface.py:75-88`` stacks the per-view outputs on axis 2)."""

    def __init__(self, spec: PredictorSpec, weights_per_view: Sequence[Dict[str, np.ndarray]], dtype=torch.float32):
        assert len(weights_per_view) == spec.ncam
        self.spec = spec
        self.views = [OraclePredictor(spec, w, dtype) for w in weights_per_view]

    def rollout(self, ctx_frames, ctx_states, ctx_distrib, step_actions, zs=None):
        """ctx_frames (C,ncam,H,W,3) in [0,1]; ctx_distrib (C,ncam,H,W,nd)|None ->
        (M,P,ncam,H,W,3), (M,P,ncam,H,W,nd)|None, (M,P,sdim)|None"""
        gi, gd, gs = [], [], None
        for v, p in enumerate(self.views):
            d = None if ctx_distrib is None else np.asarray(ctx_distrib)[:, v]
            i_, d_, s_ = p.rollout(np.asarray(ctx_frames)[:, v], ctx_states, d, step_actions, zs)
            gi.append(i_)
            gd.append(d_)
            if v == 0:
                gs = s_
        gen_i = np.stack(gi, axis=2)
        gen_d = None if gd[0] is None else np.stack(gd, axis=2)
        return gen_i, gen_d, gs

"""TF1 SAVP checkpoint -> engine weights, without TensorFlow.

The reference restores its predictor with ``Saver.restore`` after mapping graph variables onto checkpoint variables by
path SUFFIX (``video_prediction/checkpoint_matcher.py:4-38``), picks the newest ``model*`` file of a directory
(``video_prediction/setup_predictor.py:12-28``) and reads the architecture from ``model_hparams.json`` /
``dataset_hparams.json`` next to the checkpoint (``video_prediction/vpred_model_interface.py:20-58``).  TensorFlow is not
available where the engine runs, so the input here is a dump of the checkpoint as ``{variable name: ndarray}`` (an
``.npz``; one line on any TF1 box: ``np.savez(out, **{n: r.get_tensor(n) for n in r.get_variable_to_shape_map()})`` with
``r = tf.train.NewCheckpointReader(ckpt)``).

What is pinned: the matching rule, the newest-checkpoint rule and the hparams ingestion restate reference code.
What is NOT pinned: ``TF_NAMES`` — the variable names of the external ``video_prediction`` package (un-vendored, un-pinned,
SURVEY.md 8c) are written from memory of the public upstream and can be overridden with a JSON table; the layer
semantics behind each name are spec P.  Shapes are checked tensor by tensor, so a wrong guess fails loudly.

Composite-layer ORDER (must be settled before a real checkpoint is loaded; irrelevant for random weights): spec P feeds the
mask-logit conv ``concat(h_masks, [T_0..T_{n-1}, prev image, first image, scratch])`` and softmaxes the masks in that order.
Upstream SAVP (from memory) orders the composited layers ``[prev image, first image, scratch, T_0..]``.  If that is what
the checkpoint was trained with, ``masks.conv1`` needs (a) its INPUT channels [ngf + 0 .. ngf + 3*n_layers) permuted from
the upstream layer order to spec P's and (b) its OUTPUT channels (one mask per layer) permuted the same way.
``permute_mask_layers`` does exactly that; pass ``--upstream-layer-order`` to the CLI.  Unverifiable in this environment
(no TF1, no checkpoint) — hence an explicit switch, not a silent guess.

    python -m visual_foresight_b200.checkpoint dump.npz[,dump_view1.npz] model_hparams.json out_weights.npz
           [--dataset-hparams dataset_hparams.json] [--conf conf.json] [--names table.json] [--upstream-layer-order]
"""
from __future__ import annotations

import json
import re
import sys
from typing import Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np

from . import spec as specmod
from .spec import PredictorSpec


# ---- reference rules, restated ---------------------------------------------------------------------------------------
def match_variables(graph_names: Iterable[str], checkpoint_names: Sequence[str], ignore_varname_firstag: bool = False) -> Dict[str, str]:
    """``variable_checkpoint_matcher`` (checkpoint_matcher.py:20-38): for every graph variable, the FIRST checkpoint
    variable (in checkpoint order) whose '/'-separated path ends with the graph variable's path; ``:0`` output suffixes
    are dropped; optionally the first path component of the graph name is ignored.  Returns {checkpoint name: graph
    name}; raises ValueError("did not find variable ...") like the reference."""
    out: Dict[str, str] = {}
    for var in graph_names:
        varname = var.split(":")[0]
        parts = varname.split("/")
        if ignore_varname_firstag:
            parts = parts[1:]
        for ck in checkpoint_names:
            if parts == ck.split("/")[-len(parts):]:
                out[ck] = varname
                break
        else:
            raise ValueError("did not find variable {}".format(varname))
    return out


def newest_checkpoint(filenames: Sequence[str]) -> Optional[str]:
    """``get_maxiter_weights`` (setup_predictor.py:12-28) on a directory listing of ``model*`` files: the file whose name
    ends in the largest integer (names without one count as -1), cut at the first '.'."""
    if not filenames:
        return None
    nums = []
    for f in filenames:
        m = re.match(r".*?([0-9]+)$", f)
        nums.append(int(m.group(1)) if m else -1)
    return filenames[int(np.argmax(np.array(nums)))].split(".")[0]


def spec_from_hparams(model_hparams: Mapping, dataset_hparams: Optional[Mapping] = None, conf: Optional[Mapping] = None) -> PredictorSpec:
    """``VPred_Model_Interface.__init__`` (vpred_model_interface.py:20-58): architecture from model_hparams.json (with
    conf['override_json'] applied, 'num_gpus' dropped), frame size / view count / action-state sizes from the net conf."""
    mh = dict(model_hparams)
    mh.pop("num_gpus", None)
    conf = dict(conf or {})
    mh.update(conf.get("override_json", {}))
    dh = dict(dataset_hparams or {})
    H, W = conf.get("orig_size", (mh.get("height", 64), mh.get("width", 64)))
    adim = dh["autograsp"] if "autograsp" in dh else conf.get("adim", 4)      # vpred_model_interface.py:28-30
    # the states are model inputs only when dataset_hparams.json says so (vpred_model_interface.py:61: use_state =
    # dataset_hparams.get('use_state', False)); conf['sdim'] alone does NOT make the model stateful
    use_state = bool(dh.get("use_state", False))
    kw = dict(height=int(H), width=int(W), ncam=int(conf.get("ncam", 1)), ndesig=int(conf.get("ndesig", 1)),
              adim=int(adim), sdim=int(conf.get("sdim", 0)) if use_state else 0,
              seq_len=int(mh.get("sequence_length", conf.get("sequence_length", 15))),
              context_frames=int(mh.get("context_frames", conf.get("context_frames", 2))),
              ngf=int(mh.get("ngf", 32)), num_transformed=int(mh.get("num_transformed_images", 4)),
              nz=int(mh.get("nz", 0)), rnn_z=bool(mh.get("use_rnn_z", False)) and int(mh.get("nz", 0)) > 0)
    ks = mh.get("kernel_size", (5, 5))
    kw["cdna_ksize"] = int(ks[0] if isinstance(ks, (list, tuple)) else ks)
    family = specmod.spec_128 if int(H) >= 128 else specmod.spec_64
    if family is specmod.spec_128:
        kw.pop("height"), kw.pop("width")
        return specmod.spec_128(height=int(H), width=int(W), **kw)
    return specmod.spec_64(**kw)


# ---- name table (UPSTREAM-FROM-MEMORY, overridable) ---------------------------------------------------------------------
def default_tf_names(spec: PredictorSpec) -> Dict[str, str]:
    """engine weight name -> TF variable path SUFFIX of the SAVP generator cell (scopes above it — 'generator/rnn/...',
    per-view model scopes — are absorbed by suffix matching).  Layer scopes h0..h{n-1} number the encoder then the decoder
    convs, 'conv{L}_rnn' the conv-LSTM behind layer L."""
    t: Dict[str, str] = {}
    n = len(spec.encoder)
    layers = [("enc%d" % i, i, rnn) for i, (_, rnn) in enumerate(spec.encoder)] + \
             [("dec%d" % i, n + i, rnn) for i, (_, rnn) in enumerate(spec.decoder)]
    for pre, L, rnn in layers:
        t[pre + ".conv.w"] = "h%d/conv2d/kernel" % L
        t[pre + ".conv.b"] = "h%d/conv2d/bias" % L
        t[pre + ".conv.gamma"] = "h%d/InstanceNorm/gamma" % L
        t[pre + ".conv.beta"] = "h%d/InstanceNorm/beta" % L
        if rnn:
            t[pre + ".lstm.w"] = "conv%d_rnn/gates/kernel" % L
            t[pre + ".lstm.gates_gamma"] = "conv%d_rnn/gates/InstanceNorm/gamma" % L
            t[pre + ".lstm.gates_beta"] = "conv%d_rnn/gates/InstanceNorm/beta" % L
            t[pre + ".lstm.cell_gamma"] = "conv%d_rnn/state/InstanceNorm/gamma" % L
            t[pre + ".lstm.cell_beta"] = "conv%d_rnn/state/InstanceNorm/beta" % L
    t["cdna.dense.w"] = "cdna_kernels/dense/kernel"
    t["cdna.dense.b"] = "cdna_kernels/dense/bias"
    for pre, scope in (("scratch.conv0", "h%d_scratch" % (2 * n)), ("masks.conv0", "h%d_masks" % (2 * n))):
        t[pre + ".w"] = scope + "/conv2d/kernel"
        t[pre + ".b"] = scope + "/conv2d/bias"
        t[pre + ".gamma"] = scope + "/InstanceNorm/gamma"
        t[pre + ".beta"] = scope + "/InstanceNorm/beta"
    t["scratch.conv1.w"], t["scratch.conv1.b"] = "scratch_image/conv2d/kernel", "scratch_image/conv2d/bias"
    t["masks.conv1.w"], t["masks.conv1.b"] = "masks/conv2d/kernel", "masks/conv2d/bias"
    if spec.sdim > 0:
        t["state.dense.w"], t["state.dense.b"] = "state_pred/dense/kernel", "state_pred/dense/bias"
    if spec.rnn_z:
        t["zrnn.w"], t["zrnn.b"] = "rnn_z/basic_lstm_cell/kernel", "rnn_z/basic_lstm_cell/bias"
    return t


def convert_checkpoint(arrays: Mapping[str, np.ndarray], spec: PredictorSpec, names: Optional[Mapping[str, str]] = None,
                       ignore_varname_firstag: bool = False) -> Dict[str, np.ndarray]:
    """{TF variable name: array} -> engine weights of ONE view.  Every engine tensor is located with the reference's suffix
    rule and checked against ``spec.weight_shapes``; TF conv kernels are HWIO with the input-channel order of the graph's
    concats ([x, tiled action/state/z] and [x, sa, h] for the conv-LSTM), which is the engine's layout, so no transpose."""
    table = dict(default_tf_names(spec))
    if names:
        table.update(names)
    shapes = specmod.weight_shapes(spec)
    missing = [k for k in shapes if k not in table]
    if missing:
        raise ValueError("no TF name for engine tensors: %s" % ", ".join(missing))
    ck_names = list(arrays.keys())
    matched = match_variables([table[k] for k in shapes], ck_names, ignore_varname_firstag)
    by_suffix = {v: k for k, v in matched.items()}                    # graph suffix -> checkpoint name
    out: Dict[str, np.ndarray] = {}
    for k, shp in shapes.items():
        a = np.asarray(arrays[by_suffix[table[k]]])
        if tuple(a.shape) != tuple(shp):
            raise ValueError("%s: checkpoint tensor %s has shape %s, spec P wants %s" % (k, by_suffix[table[k]], a.shape, shp))
        out[k] = np.ascontiguousarray(a, dtype=np.float32)
    return out


def permute_mask_layers(weights: Dict[str, np.ndarray], spec: PredictorSpec) -> Dict[str, np.ndarray]:
    """masks.conv1 trained with the upstream layer order [prev, first, scratch, T_0..T_{n-1}] -> spec P's order
    [T_0..T_{n-1}, prev, first, scratch]: permutes the layer blocks (3 input channels each, behind the ngf h_masks channels)
    and the per-layer output channels of ``masks.conv1`` (see the module docstring)."""
    nt, g = spec.num_transformed, spec.ngf
    nm = nt + 3
    src_of = [3 + i for i in range(nt)] + [0, 1, 2]                 # spec-P layer j comes from upstream layer src_of[j]
    w, b = np.array(weights["masks.conv1.w"]), np.array(weights["masks.conv1.b"])
    cin = np.concatenate([np.arange(g)] + [g + 3 * s + np.arange(3) for s in src_of])
    out = dict(weights)
    out["masks.conv1.w"] = np.ascontiguousarray(w[:, :, cin][..., src_of])
    out["masks.conv1.b"] = np.ascontiguousarray(b[src_of])
    assert out["masks.conv1.w"].shape == w.shape and len(src_of) == nm
    return out


def export_as_tf(weights: Mapping[str, np.ndarray], spec: PredictorSpec, scope: str = "generator/rnn/dna_cell",
                 names: Optional[Mapping[str, str]] = None) -> Dict[str, np.ndarray]:
    """inverse of convert_checkpoint (used by the round-trip test and to hand engine weights back to a TF1 graph)."""
    table = dict(default_tf_names(spec))
    if names:
        table.update(names)
    return {scope + "/" + table[k]: np.asarray(v) for k, v in weights.items()}


def main(argv: List[str]) -> int:
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("dump"), ap.add_argument("model_hparams"), ap.add_argument("out")
    ap.add_argument("--dataset-hparams"), ap.add_argument("--conf", help="JSON of the net conf keys (orig_size, ncam, adim, sdim, ndesig)")
    ap.add_argument("--names", help="JSON {engine name: TF suffix} overriding the built-in table")
    ap.add_argument("--upstream-layer-order", action="store_true",
                    help="masks.conv1 was trained with the layer order [prev, first, scratch, T_0..]: permute it to spec P's")
    a = ap.parse_args(argv)
    load = lambda p: json.load(open(p)) if p else None
    spec = spec_from_hparams(load(a.model_hparams), load(a.dataset_hparams), load(a.conf))
    dumps = a.dump.split(",")                       # one checkpoint dump per view (IndepMultiSAVP: independent weight sets,
    if len(dumps) != spec.ncam:                     # experiments/sawyer/pixel_cost/conf.py:12-15)
        raise SystemExit("the net conf has ncam=%d: pass %d comma-separated checkpoint dumps, one per view (got %d)"
                         % (spec.ncam, spec.ncam, len(dumps)))
    views = []
    for d in dumps:
        w = convert_checkpoint(dict(np.load(d)), spec, load(a.names))
        views.append(permute_mask_layers(w, spec) if a.upstream_layer_order else w)
    from .predictor import save_weights
    save_weights(a.out, spec, views)
    print("wrote %d tensors x %d view(s) -> %s" % (len(views[0]), len(views), a.out))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))

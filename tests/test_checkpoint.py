"""TF1 checkpoint ingestion without TensorFlow: the reference's suffix-matching rule (checkpoint_matcher.py:20-38), its
newest-checkpoint rule (setup_predictor.py:12-28), model_hparams.json ingestion (vpred_model_interface.py:20-58) and the
engine <-> TF name table round trip.  The TF names themselves are from memory of the un-vendored package (unpinned)."""
import numpy as np
import pytest

from visual_foresight_b200 import checkpoint as CK
from visual_foresight_b200 import spec as S


def test_match_variables_suffix_rule():
    ck = ["model/generator/rnn/cell/h0/conv2d/kernel", "other/h0/conv2d/kernel", "model/generator/rnn/cell/h0/conv2d/bias"]
    m = CK.match_variables(["h0/conv2d/kernel:0", "h0/conv2d/bias"], ck)
    assert m == {"model/generator/rnn/cell/h0/conv2d/kernel": "h0/conv2d/kernel",       # FIRST match in checkpoint order wins
                 "model/generator/rnn/cell/h0/conv2d/bias": "h0/conv2d/bias"}
    # whole path components only: 'xh0/conv2d/kernel' does not match 'h0/conv2d/kernel'
    with pytest.raises(ValueError, match="did not find variable h0/conv2d/kernel"):
        CK.match_variables(["h0/conv2d/kernel"], ["a/xh0/conv2d/kernel"])
    # ignore_varname_firstag drops the graph name's first component (towers: 'tower_1/h0/...')
    assert CK.match_variables(["tower_1/h0/conv2d/kernel"], ck, ignore_varname_firstag=True) == \
        {"model/generator/rnn/cell/h0/conv2d/kernel": "tower_1/h0/conv2d/kernel"}


def test_newest_checkpoint_rule():
    files = ["d/model-100.index", "d/model-300000", "d/model-20000", "d/modelfoo"]
    assert CK.newest_checkpoint(files) == "d/model-300000"
    assert CK.newest_checkpoint(["d/model.savp.None/model-30.meta7"]) == "d/model"        # cut at the first '.'
    assert CK.newest_checkpoint([]) is None


def test_spec_from_hparams():
    mh = {"sequence_length": 13, "context_frames": 2, "ngf": 32, "num_transformed_images": 4, "num_gpus": 4, "kernel_size": [5, 5]}
    conf = {"orig_size": [48, 64], "ncam": 2, "ndesig": 2, "adim": 5, "sdim": 5, "override_json": {"sequence_length": 15}}
    sp = CK.spec_from_hparams(mh, {"autograsp": 4, "use_state": True}, conf)
    assert (sp.seq_len, sp.context_frames, sp.height, sp.width, sp.ncam, sp.ndesig, sp.adim, sp.sdim) == (15, 2, 48, 64, 2, 2, 4, 5)
    # the reference takes use_state from dataset_hparams.json (vpred_model_interface.py:61): conf['sdim'] > 0 alone, or a
    # model_hparams key, does not make the model stateful
    assert CK.spec_from_hparams(dict(mh, use_state=True), {"autograsp": 4}, conf).sdim == 0
    assert CK.spec_from_hparams(mh, None, conf).sdim == 0
    assert len(CK.spec_from_hparams({}, None, {"orig_size": [128, 128]}).encoder) == 4     # 128-px family


@pytest.mark.parametrize("family,kw", [("64", dict(height=48, width=64, sdim=5)), ("128", dict(seq_len=6, nz=8, rnn_z=True))])
def test_round_trip_and_shape_checks(family, kw):
    sp = (S.spec_128 if family == "128" else S.spec_64)(**kw)
    w = S.init_weights(sp, seed=3)
    tf = CK.export_as_tf(w, sp, scope="tower_0/generator/rnn/dna_cell")
    tf = dict(reversed(list(tf.items())))                   # checkpoint order is arbitrary
    tf["global_step"] = np.zeros((), np.int64)              # unrelated checkpoint variables are ignored
    back = CK.convert_checkpoint(tf, sp)
    assert set(back) == set(w)
    for k in w:
        np.testing.assert_array_equal(back[k], w[k])
    bad = dict(tf)
    k0 = next(k for k in bad if k.endswith("h0/conv2d/kernel"))
    bad[k0] = bad[k0][..., :-1]
    with pytest.raises(ValueError, match="enc0.conv.w"):
        CK.convert_checkpoint(bad, sp)
    del bad[k0]
    with pytest.raises(ValueError, match="did not find variable h0/conv2d/kernel"):
        CK.convert_checkpoint(bad, sp)


def test_cli(tmp_path):
    import json
    sp = S.spec_64(height=32, width=32, seq_len=6, sdim=4)
    w = S.init_weights(sp, seed=1)
    np.savez(tmp_path / "dump.npz", **CK.export_as_tf(w, sp))
    (tmp_path / "mh.json").write_text(json.dumps({"sequence_length": 6, "context_frames": 2}))
    (tmp_path / "dh.json").write_text(json.dumps({"use_state": True}))
    (tmp_path / "conf.json").write_text(json.dumps({"orig_size": [32, 32], "adim": 4, "sdim": 4}))
    assert CK.main([str(tmp_path / "dump.npz"), str(tmp_path / "mh.json"), str(tmp_path / "out.npz"), "--conf", str(tmp_path / "conf.json"),
                    "--dataset-hparams", str(tmp_path / "dh.json")]) == 0
    from visual_foresight_b200.predictor import load_weights
    sp2, views = load_weights(str(tmp_path / "out.npz"))
    assert sp2 == sp
    for k in w:
        np.testing.assert_array_equal(views[0][k], w[k])
    # two views = two checkpoint dumps (independent weight sets); one dump for ncam=2 is refused
    sp_mv = S.spec_64(height=32, width=32, seq_len=6, sdim=4, ncam=2)
    w1 = S.init_weights(sp_mv, seed=2, view=1)
    np.savez(tmp_path / "dump1.npz", **CK.export_as_tf(w1, sp_mv))
    (tmp_path / "conf2.json").write_text(json.dumps({"orig_size": [32, 32], "adim": 4, "sdim": 4, "ncam": 2}))
    args = [str(tmp_path / "mh.json"), str(tmp_path / "out2.npz"), "--conf", str(tmp_path / "conf2.json"), "--dataset-hparams", str(tmp_path / "dh.json")]
    with pytest.raises(SystemExit):
        CK.main([str(tmp_path / "dump.npz")] + args)
    assert CK.main([str(tmp_path / "dump.npz") + "," + str(tmp_path / "dump1.npz")] + args) == 0
    sp3, views = load_weights(str(tmp_path / "out2.npz"))
    assert sp3.ncam == 2 and len(views) == 2
    np.testing.assert_array_equal(views[1]["masks.conv1.w"], w1["masks.conv1.w"])


def test_permute_mask_layers_is_the_stated_permutation():
    """upstream order [prev, first, scratch, T0..T3] -> spec P [T0..T3, prev, first, scratch] on masks.conv1 (input layer blocks
    behind the ngf hidden channels, and the per-layer outputs)."""
    sp = S.spec_64(height=32, width=32, seq_len=6)
    g, nm = sp.ngf, sp.num_transformed + 3
    w = S.init_weights(sp, seed=5)
    cin = w["masks.conv1.w"].shape[2]
    tagged = dict(w)
    up = np.zeros_like(w["masks.conv1.w"])
    for ci in range(cin):                                  # encode (input channel, output channel) in the values
        for co in range(nm):
            up[:, :, ci, co] = 100 * ci + co
    tagged["masks.conv1.w"], tagged["masks.conv1.b"] = up, np.arange(nm, dtype=np.float32)
    out = CK.permute_mask_layers(tagged, sp)
    src = [3, 4, 5, 6, 0, 1, 2]
    np.testing.assert_array_equal(out["masks.conv1.b"], np.array(src, np.float32))
    for j, s_ in enumerate(src):
        for c in range(3):
            assert out["masks.conv1.w"][0, 0, g + 3 * j + c, j] == 100 * (g + 3 * s_ + c) + s_
    assert out["masks.conv1.w"][0, 0, 5, 2] == 100 * 5 + src[2]          # hidden channels keep their place

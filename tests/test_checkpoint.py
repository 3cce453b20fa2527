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
    sp = CK.spec_from_hparams({"sequence_length": 13, "context_frames": 2, "ngf": 32, "num_transformed_images": 4, "num_gpus": 4,
                               "kernel_size": [5, 5], "use_state": True},
                              {"autograsp": 4}, {"orig_size": [48, 64], "ncam": 2, "ndesig": 2, "adim": 5, "sdim": 5,
                                                 "override_json": {"sequence_length": 15}})
    assert (sp.seq_len, sp.context_frames, sp.height, sp.width, sp.ncam, sp.ndesig, sp.adim, sp.sdim) == (15, 2, 48, 64, 2, 2, 4, 5)
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
    (tmp_path / "mh.json").write_text(json.dumps({"sequence_length": 6, "context_frames": 2, "use_state": True}))
    (tmp_path / "conf.json").write_text(json.dumps({"orig_size": [32, 32], "adim": 4, "sdim": 4}))
    assert CK.main([str(tmp_path / "dump.npz"), str(tmp_path / "mh.json"), str(tmp_path / "out.npz"), "--conf", str(tmp_path / "conf.json")]) == 0
    from visual_foresight_b200.predictor import load_weights
    sp2, views = load_weights(str(tmp_path / "out.npz"))
    assert sp2 == sp
    for k in w:
        np.testing.assert_array_equal(views[0][k], w[k])

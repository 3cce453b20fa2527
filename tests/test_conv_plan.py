"""Resource invariants of the tcgen05 convolution's tiling for every layer shape of the supported predictor families, checked
on the CPU through the host-only C-ABI entry vf_debug_conv_plan (no device): dynamic shared memory within the 227 KB
opt-in limit, accumulator sets within the 512 TMEM columns, legal MMA N, at least two weight stages, the shared-memory
plane covering every pixel row an item's MMAs read, and the expected tiling per layer class (row groups for 8- and
16-pixel-wide gate convolutions, row-stacked tiling for the thin 3x3 layers)."""
import ctypes as C

import numpy as np
import pytest

from visual_foresight_b200 import spec as S
from visual_foresight_b200.engine import load_library

F = ["swap", "rg", "G", "npass", "v_cnt", "units", "ncols", "nmax", "acc_cols", "nacc", "nbuf", "nstage", "stage_bytes",
     "plane_bytes", "smem", "nitems", "R", "Wp", "img_pix", "box_bytes", "last_read", "plane_rows", "nchunk", "n_mt"]


@pytest.fixture(scope="module", autouse=True)
def _built():
    from visual_foresight_b200 import build
    build.build()                                      # in-tree nvcc build if the library is missing or stale (no GPU needed)


def plan(k, kw, cin, cout, H, W, B, passes=3):
    lib = load_library()
    out = (C.c_int32 * 24)()
    rc = lib.vf_debug_conv_plan(k, kw, cin, cout, H, W, B, passes, C.cast(out, C.c_void_p))
    return None if rc else dict(zip(F, list(out)))


def layer_shapes(sp):
    """(name, k, Cin executed, Cout, H, W) of one cell step on the tensor-core path (action/state channels folded into the
    bias; the first conv + pool as the 3x3 conv over 2x2 pixel blocks)."""
    h, w, cprev, out, enc_out = sp.height, sp.width, 6, [], []
    for i, (oc, rnn) in enumerate(sp.encoder):
        if i == 0:
            out.append(("enc0", 3, 32, oc, h // 2, w // 2))
        else:
            out.append(("enc%d" % i, 3, cprev, oc, h, w))
        h, w = h // 2, w // 2
        if rnn:
            out.append(("enc%d.lstm" % i, sp.lstm_ksize, 2 * oc, 4 * oc, h, w))
        enc_out.append(oc)
        cprev = oc
    n = len(sp.encoder)
    for i, (oc, rnn) in enumerate(sp.decoder):
        cin = cprev + (enc_out[n - 1 - i] if i > 0 else 0)
        h, w = h * 2, w * 2
        out.append(("dec%d" % i, 3, cin, oc, h, w))
        if rnn:
            out.append(("dec%d.lstm" % i, sp.lstm_ksize, 2 * oc, 4 * oc, h, w))
        cprev = oc
    g = sp.ngf
    out += [("scratch0", 3, g, g, h, w), ("scratch1", 3, g, 3, h, w), ("masks0", 3, g, g, h, w),
            ("masks1", 3, g + 24, sp.n_masks, h, w)]
    return out


SPECS = {"c2 64x64": S.spec_64(height=64, width=64), "c1/c3 48x64": S.spec_64(height=48, width=64),
         "32x32": S.spec_64(height=32, width=32), "c5 128x128": S.spec_128()}


@pytest.mark.parametrize("tag", list(SPECS))
@pytest.mark.parametrize("B", [1, 5, 200, 512])
def test_every_layer_has_a_legal_plan(tag, B):
    sp = SPECS[tag]
    for name, k, cin, cout, H, W in layer_shapes(sp):
        p = plan(k, k, cin, cout, H, W, B)
        if p is None:                                  # tiny maps may fall back to the FFMA convolution; never the gate convs of c2 / c5
            assert H * W <= 16, (tag, name)
            continue
        where = "%s %s B=%d %r" % (tag, name, B, p)
        assert p["smem"] <= 227 * 1024, where
        assert p["acc_cols"] * p["nacc"] <= 512 and p["nacc"] in (1, 2), where
        assert 16 <= p["nmax"] <= 256 and p["nmax"] % 16 == 0, where
        assert p["nstage"] >= 2 and p["nbuf"] in (1, 2) and p["nitems"] >= 1, where
        assert p["last_read"] < p["plane_rows"], where
        assert p["R"] <= 256 and p["Wp"] <= 256, where       # TMA box extents
        if k == 5 and cout >= 128:                            # gate convolutions: wide tiling, row groups on narrow maps
            assert p["swap"] == 0, where
            strip = W % 8 == 0 and H % 4 == 0 and 8 * H <= 256            # 8-pixel column strips (row-group mode 3)
            want = 1 if W == 8 else (0 if not strip else (2 if (W == 16 and cout % 256 == 0) else (3 if W >= 16 else 0)))
            assert p["rg"] == want, where
            assert p["nacc"] == 2, where                      # the epilogue overlaps the next item's MMAs
        if k == 3 and cout <= 64:
            assert p["swap"] == 2 and p["units"] <= 4, where


def test_headline_layer_plans_are_the_documented_ones():
    """lstm0 (32x32) = four 8-pixel column strips per image, one N = 256 MMA each (row-group mode 3), lstm1 (16x16)
    = two N = 128 row-group MMAs, lstm2 (8x8) = one N = 224 MMA over 3 stacked images; work-item counts at M = 200."""
    p0 = plan(5, 5, 64, 128, 32, 32, 200)
    # 800 strips on 148 CTAs: 5 whole rounds (740) + the last 60 strips as 120 half strips
    assert (p0["npass"], p0["v_cnt"], p0["nitems"], p0["rg"], p0["Wp"]) == (4, 256, 740 + 120, 3, 12)
    p1 = plan(5, 5, 128, 256, 16, 16, 200)
    assert (p1["rg"], p1["nmax"], p1["acc_cols"], p1["nitems"]) == (2, 128, 256, 400)
    p2 = plan(5, 5, 256, 512, 8, 8, 200)
    assert (p2["rg"], p2["G"], p2["nmax"], p2["nitems"]) == (1, 3, 224, 4 * 67)
    pf = plan(5, 1, 40, 32, 64, 64, 200)                      # the dx-folded first conv (VF_ENC0=fold): per-tap thin tiling
    assert pf is not None and pf["swap"] == 1


def test_random_shapes_keep_the_resource_invariants():
    """Seeded sweep over layer shapes outside the shipped specs (odd heights, wide maps, many channel counts): whatever tiling
    plan_geometry picks — generic, row groups 1 / 2 / 3, row-stacked, stacked weight halves — stays inside shared memory, TMEM,
    the legal MMA N and its staged pixel rows; the strip tilings are picked exactly where their preconditions hold."""
    rng = np.random.default_rng(7)
    seen = set()
    for _ in range(400):
        k = int(rng.choice([3, 5]))
        cin = int(rng.choice([8, 16, 24, 32, 40, 56, 64, 96, 128, 192, 256]))
        cout = int(rng.choice([3, 7, 16, 32, 48, 64, 96, 128, 160, 256, 512]))
        H = int(rng.choice([4, 6, 8, 12, 16, 20, 24, 32, 48, 64]))
        W = int(rng.choice([8, 12, 16, 24, 32, 40, 64, 96]))
        B = int(rng.choice([1, 3, 25, 200]))
        p = plan(k, k, cin, cout, H, W, B)
        if p is None:
            continue
        where = "k%d %d->%d %dx%d B=%d %r" % (k, cin, cout, H, W, B, p)
        assert p["smem"] <= 227 * 1024, where
        assert p["acc_cols"] * p["nacc"] <= 512 and p["nacc"] in (1, 2), where
        assert 16 <= p["nmax"] <= 256 and p["nmax"] % 16 == 0, where
        assert p["nstage"] >= 2 and p["nbuf"] in (1, 2) and p["nitems"] >= 1, where
        assert p["last_read"] < p["plane_rows"], where
        assert p["R"] <= 256 and p["Wp"] <= 256, where
        if p["rg"] == 3:
            assert W % 8 == 0 and W >= 16 and H % 4 == 0 and 8 * H <= 256 and p["Wp"] == 8 + k - 1 and p["npass"] == W // 8, where
            strips = p["n_mt"] * B * (W // 8)                     # + the strips of a split last round, counted twice
            assert p["v_cnt"] == 8 * H and strips <= p["nitems"] <= strips + 74, where
            if p["nitems"] > strips:
                assert (4 * H) % 64 == 0 and strips > 148 and 2 * (strips % 148) <= 148 and p["nitems"] == strips + strips % 148, where
        if p["rg"] == 2:
            assert W == 16 and cout % 256 == 0 and p["v_cnt"] == 16 * H, where
        seen.add((p["swap"], p["rg"]))
    assert {(0, 0), (0, 1), (0, 2), (0, 3), (2, 0)} <= seen, seen

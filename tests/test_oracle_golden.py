"""C oracle vs the committed golden vectors (tests/golden/*.npz), which were produced by an
independent Python restatement on top of real cv2 primitives (tests/golden/gen_golden.py).
Also checks the closed-form constants of SURVEY.md 8(d).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_names


def _load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def _mk(oracle, g):
    nf, sc, nl, it, mt = g["params"]
    img = g["image"]
    return oracle.Oracle(img.shape[1], img.shape[0], int(nf), float(sc), int(nl), int(it), int(mt)), img


@pytest.mark.parametrize("name", golden_names())
def test_geometry_and_pyramid(oracle, name):
    g = _load(name)
    o, img = _mk(oracle, g)
    assert np.array_equal(o.lw, g["level_w"]) and np.array_equal(o.lh, g["level_h"])
    assert np.array_equal(o.nfeat, g["nfeat"])
    assert np.array_equal(o.umax, g["umax"])
    o.compute_pyramid(img)
    for l in range(o.nlevels):
        sha = hashlib.sha256(o.level_padded(l).tobytes()).hexdigest()
        assert sha == str(g["pyr_sha256"][l]), f"padded level {l}"
    assert np.array_equal(o.level_padded(o.nlevels - 1), g["pyr_last_padded"])


@pytest.mark.parametrize("name", golden_names())
def test_candidates(oracle, name):
    g = _load(name)
    o, img = _mk(oracle, g)
    o.compute_pyramid(img)
    for l in range(o.nlevels):
        c = o.level_candidates(l)
        got = np.stack([c["x"], c["y"], c["response"]], 1).astype(np.int16) if c.size else np.zeros((0, 3), np.int16)
        assert np.array_equal(got, g[f"cand{l}"]), f"level {l}"  # same set AND upstream order


@pytest.mark.parametrize("name", golden_names())
def test_extract_end_to_end(oracle, name):
    g = _load(name)
    o, img = _mk(oracle, g)
    kp, desc = o.extract(img)
    assert kp.shape == g["kp"].shape
    for f in kp.dtype.names:  # bit-exact incl. float fields and list order
        assert np.array_equal(kp[f], g["kp"][f]), f
    assert np.array_equal(desc, g["desc"])


def test_survey_constants(oracle):
    o = oracle.Oracle(640, 480)
    assert o.nfeat.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(zip(o.lw.tolist(), o.lh.tolist())) == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                       (257, 193), (214, 161), (179, 134)]
    assert o.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    o = oracle.Oracle(1280, 720, nfeatures=2000)
    assert o.nfeat.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert o.lw.tolist() == [1280, 1067, 889, 741, 617, 514, 429, 357]
    o = oracle.Oracle(848, 800)
    assert o.lh.tolist() == [800, 667, 556, 463, 386, 322, 268, 223]


def test_determinism(oracle, synth):
    img = synth.textured_frame(320, 240, 123)
    o = oracle.Oracle(320, 240, 400)
    a = o.extract(img)
    b = o.extract(img)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


def test_rejects_too_small(oracle):
    with pytest.raises(ValueError):
        oracle.Oracle(120, 100)  # level 7 would be 33x28: upstream divides by nCols == 0

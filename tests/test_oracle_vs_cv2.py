"""Pin every OpenCV primitive of the CPU oracle bit-exactly against the cv2 build in this image
(OpenCV 4.13).  SURVEY.md 8(c) / Appendix B.  CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _rand(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w), dtype=np.uint8)


SIZES = [(640, 480), (848, 480), (848, 800), (1280, 720), (333, 251), (100, 77)]


@pytest.mark.parametrize("w,h", SIZES)
def test_resize_chain_matches_cv2(oracle, w, h):
    """7 level transitions with the upstream float32 scale chain, each from the previous level."""
    sf = np.float32(1.0)
    cur = _rand(h, w, w * 7 + h)
    for lvl in range(1, 8):
        sf = np.float32(sf * np.float32(1.2))
        inv = np.float32(np.float32(1.0) / sf)
        dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        got = oracle.resize_linear(cur, dw, dh)
        assert np.array_equal(ref, got), f"level {lvl} {cur.shape}->{(dh, dw)}"
        cur = ref


@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 480, 427, 320), (640, 480, 320, 240), (300, 200, 299, 199),
                                         (97, 131, 64, 90), (500, 375, 250, 187), (64, 64, 80, 80)])
def test_resize_misc_ratios(oracle, sw, sh, dw, dh):
    src = _rand(sh, sw, sw + dw)
    assert np.array_equal(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR), oracle.resize_linear(src, dw, dh))


@pytest.mark.parametrize("w,h", [(640, 480), (179, 134), (40, 21)])
def test_border_reflect101(oracle, w, h):
    roi = _rand(h, w, 3)
    ref = cv2.copyMakeBorder(roi, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, oracle.border_reflect101(roi, 19))


def _cv_fast(win, thr, nms=True):
    det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=nms, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    return [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(np.ascontiguousarray(win))]


@pytest.mark.parametrize("thr", [20, 7])
@pytest.mark.parametrize("kind", ["noise", "tex", "checker", "steps"])
def test_fast_window_matches_cv2(oracle, synth, thr, kind):
    if kind == "noise":
        img = _rand(120, 160, 5)
    elif kind == "tex":
        img = synth.textured_frame(320, 240, 42)
    elif kind == "checker":
        img = synth.checkerboard_frame(160, 120, 8)
    else:
        img = (np.add.outer(np.arange(120) // 5, np.arange(160) // 7) * 23 % 256).astype(np.uint8)
    for (x0, y0, cw, ch) in [(0, 0, img.shape[1], img.shape[0]), (16, 16, 36, 37), (50, 40, 7, 7), (3, 9, 65, 33)]:
        win = img[y0:y0 + ch, x0:x0 + cw]
        ref = _cv_fast(win, thr, True)
        got = oracle.fast9_window(win, thr, True)
        assert ref == [(int(c["x"]), int(c["y"]), int(c["response"])) for c in got]  # incl. order
    # without NMS the corner SET must agree (cv2 reports response 0 there)
    ref = {(x, y) for x, y, _ in _cv_fast(img, thr, False)}
    got = oracle.fast9_window(img, thr, False)
    assert ref == {(int(c["x"]), int(c["y"])) for c in got}


def test_score_map_consistent_with_fast(oracle, synth):
    """threshold-free score m: corner at t <=> m > t, response = m - 1."""
    img = synth.textured_frame(200, 150, 77)
    m = oracle.fast_score_map(img).astype(np.int32)
    for thr in (7, 20, 35):
        ref = _cv_fast(img, thr, False)
        mask = np.zeros_like(m, bool)
        for x, y, _ in ref:
            mask[y, x] = True
        assert np.array_equal(mask, m > thr)
    for x, y, r in _cv_fast(img, 7, True):
        assert m[y, x] - 1 == r


@pytest.mark.parametrize("w,h", [(640, 480), (179, 134), (237, 223), (64, 62)])
def test_gaussian_blur7(oracle, w, h):
    img = _rand(h, w, w)
    ref = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, oracle.gaussian_blur7(img))


def test_fast_atan2(oracle):
    rng = np.random.default_rng(0)
    ys = rng.integers(-60000, 60001, size=20000)
    xs = rng.integers(-60000, 60001, size=20000)
    ys[:8] = [0, 0, 0, 1, -1, 5, -5, 0]
    xs[:8] = [0, 1, -1, 0, 0, 5, -5, 0]
    for y, x in zip(ys.tolist(), xs.tolist()):
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(oracle.fast_atan2(float(y), float(x)))


def test_match_knn_vs_bfmatcher(oracle):
    rng = np.random.default_rng(1)
    t = rng.integers(0, 256, size=(700, 32), dtype=np.uint8)
    q = t[rng.integers(0, 700, size=300)].copy()
    flip = rng.integers(0, 256, size=q.shape, dtype=np.uint8) & rng.integers(0, 256, size=q.shape, dtype=np.uint8) \
        & rng.integers(0, 256, size=q.shape, dtype=np.uint8)
    q ^= flip
    idx, dist, acc = oracle.match_knn(q, t, k=2, ratio=0.7)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    for i, (m1, m2) in enumerate(knn):
        assert int(m1.distance) == dist[i, 0] and int(m2.distance) == dist[i, 1]
        assert m1.trainIdx == idx[i, 0]
        assert acc[i] == (m1.distance < 0.7 * m2.distance)

"""Pin the ORB-specific half of the CPU oracle to third-party code in this image.

tests/test_oracle_vs_cv2.py pins the OpenCV *primitives* (resize, border, FAST, blur, fastAtan2).  The ORB glue on top
of them -- IC_Angle (umax disc, integer moments), the rBRIEF pattern, the steering (cos/sin, cvRound) and the bit
packing -- is checked here against OpenCV's own ORB implementation, `cv2.ORB_create(...).detectAndCompute`, which
shares those four pieces with upstream ORB-SLAM2 (ORBextractor.cc was derived from it):

  * `Oracle.ic_angle` == `cv2.KeyPoint.angle` for every cv2.ORB keypoint,
  * `oracle.descriptor(blur, kp.pt, kp.angle)` == the cv2.ORB descriptor, bit for bit.

One documented difference: cv2.ORB blurs a SUB-MATRIX of its border-extended pyramid buffer, which sends
cv::GaussianBlur down OpenCV's float path (sepFilter2D with the CV_32F 7-tap kernel); ORB-SLAM2 blurs a CLONE of the
level (`Mat workingMat = mvImagePyramid[level].clone()`), a whole matrix, which OpenCV 4.13 serves with the 8-bit
fixed-point kernel [18,34,48,56,48,34,18]/256 -- the oracle's choice, pinned in test_oracle_vs_cv2.py::test_gaussian_blur7.
So the descriptor comparison feeds the oracle's steering code the float-kernel blur cv2.ORB used; the last test shows
that the two blurs really differ (otherwise this distinction would be untested folklore).

What stays unpinned by third-party code: DistributeOctTree and the cell loop (no implementation of them exists in
this image); those are compared between the two independent restatements in test_octree_fuzz.py / test_oracle_golden.py.
CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

GEOMETRIES = [(640, 480, 9100), (848, 480, 9101), (333, 251, 9102), (1280, 720, 9103)]


def _float_blur(img):
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    return cv2.sepFilter2D(img, -1, k, k, borderType=cv2.BORDER_REFLECT_101)


def _cv_orb(img):
    orb = cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0, WTA_K=2,
                         scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=20)
    return orb.detectAndCompute(img, None)


def test_ic_angle_and_descriptor_match_cv2_orb(oracle, synth):
    total = 0
    for w, h, seed in GEOMETRIES:
        img = synth.textured_frame(w, h, seed)
        kps, desc = _cv_orb(img)
        assert len(kps) > 500
        o = oracle.Oracle(w, h, 1000, 1.2, 1, 20, 7)
        o.compute_pyramid(img)
        blur = _float_blur(img)
        bad_angle = bad_desc = 0
        for kp, d in zip(kps, desc):
            x, y = kp.pt
            assert x == int(x) and y == int(y)  # level 0: integer pixel positions
            a = o.ic_angle(0, x, y)
            if np.float32(a) != np.float32(kp.angle):
                bad_angle += 1
            if not np.array_equal(oracle.descriptor(blur, x, y, kp.angle), d):
                bad_desc += 1
        assert bad_angle == 0, f"{w}x{h}: {bad_angle} of {len(kps)} IC_Angle values differ from cv2.ORB"
        assert bad_desc == 0, f"{w}x{h}: {bad_desc} of {len(kps)} descriptors differ from cv2.ORB"
        total += len(kps)
    assert total >= 3000


def test_degenerate_patches_match_cv2_orb(oracle, synth):
    """low contrast / checkerboard: many zero or tied moments (fastAtan2(0, 0) = 0, axis-aligned angles)"""
    for img in (synth.low_contrast_frame(320, 240, 4), synth.checkerboard_frame(320, 240, 8)):
        orb = cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0, WTA_K=2,
                             scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=7)
        kps, desc = orb.detectAndCompute(img, None)
        if not kps:
            continue
        o = oracle.Oracle(320, 240, 1000, 1.2, 1, 20, 7)
        o.compute_pyramid(img)
        blur = _float_blur(img)
        for kp, d in zip(kps, desc):
            assert np.float32(o.ic_angle(0, *kp.pt)) == np.float32(kp.angle)
            assert np.array_equal(oracle.descriptor(blur, kp.pt[0], kp.pt[1], kp.angle), d)


def test_submatrix_blur_is_the_float_path(oracle, synth):
    """The finding recorded in DESIGN.md section 2: on OpenCV 4.13 cv::GaussianBlur(7x7, sigma 2) of a whole 8-bit
    matrix is the fixed-point kernel (== oracle.gaussian_blur7), while the blur cv2.ORB applies to a sub-matrix equals
    the float kernel; the two differ on a noticeable share of pixels, and by at most 1 grey level."""
    img = synth.textured_frame(640, 480, 9100)
    fixed = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(fixed, oracle.gaussian_blur7(img))
    flt = _float_blur(img)
    diff = np.abs(fixed.astype(np.int16) - flt.astype(np.int16))
    assert 0 < (diff != 0).mean() < 0.5 and diff.max() == 1
    # and cv2.ORB's descriptors follow the float blur, not the fixed-point one
    kps, desc = _cv_orb(img)
    n_fixed_mismatch = sum(not np.array_equal(oracle.descriptor(fixed, kp.pt[0], kp.pt[1], kp.angle), d)
                           for kp, d in zip(kps, desc))
    assert n_fixed_mismatch > 0

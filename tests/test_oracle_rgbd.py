"""CPU tier: the RGB-D association oracle (oracle/rgbd_oracle.c) against independent known answers.

The reference has no test for these stages, so the oracle is pinned here against closed-form
cases derived by hand: with identical power-of-two intrinsics and identity extrinsics every float operation of
the deproject -> transform -> project chain is exact, so the aligned depth has a closed form."""
import numpy as np
import pytest


def _closed_form_identity(depth):
    """depth pixel (x, y) covers [x, x+1] x [y, y+1] of the other image; pixels whose rectangle leaves the image
    are skipped whole (reference src/cuda/cuda-align.cu:236-237); min over non-zero contributors, else 0."""
    h, w = depth.shape
    out = np.full((h, w), 0xFFFFFFFF, np.uint64)
    for y in range(h - 1):
        for x in range(w - 1):
            d = int(depth[y, x])
            if d == 0:
                continue
            out[y:y + 2, x:x + 2] = np.minimum(out[y:y + 2, x:x + 2], d)
    out[out == 0xFFFFFFFF] = 0
    return out.astype(np.uint32)


def test_align_identity_closed_form(oracle):
    rng = np.random.default_rng(7)
    w, h = 64, 48
    depth = rng.integers(300, 4000, (h, w)).astype(np.uint16)
    depth[rng.random((h, w)) < 0.2] = 0
    intr = oracle.make_intrinsics(w, h, 32.0, 24.0, 64.0, 64.0)
    got = oracle.align_depth_to_other(depth, 2.0 ** -10, intr, intr, oracle.make_extrinsics())
    assert np.array_equal(got, _closed_form_identity(depth))


def test_align_translation_shifts_by_disparity(oracle):
    """A pure x translation t at constant depth Z moves everything by fx*t/Z pixels: pick numbers where that is an
    exact integer (fx=64, t=0.5, Z=4 -> 8 px)."""
    w, h = 96, 32
    depth = np.zeros((h, w), np.uint16)
    depth[8:20, 10:40] = 4096  # * 2^-10 = 4.0
    intr = oracle.make_intrinsics(w, h, 48.0, 16.0, 64.0, 64.0)
    ex = oracle.make_extrinsics(translation=(0.5, 0, 0))
    got = oracle.align_depth_to_other(depth, 2.0 ** -10, intr, intr, ex)
    ref = np.zeros((h, w), np.uint32)
    ref[8:21, 18:49] = 4096  # rectangle grows by one row/column (both corners are rounded outwards)
    assert np.array_equal(got, ref)


def test_align_all_zero_and_out_of_view(oracle):
    w, h = 40, 30
    intr = oracle.make_intrinsics(w, h, 20.0, 15.0, 32.0, 32.0)
    assert not oracle.align_depth_to_other(np.zeros((h, w), np.uint16), 0.001, intr, intr, oracle.make_extrinsics()).any()
    far = oracle.make_extrinsics(translation=(1000.0, 0, 0))  # everything projects right of the image
    d = np.full((h, w), 1000, np.uint16)
    assert not oracle.align_depth_to_other(d, 0.001, intr, intr, far).any()


def test_keypoint_pixel_to_point_gate_order_and_values(oracle):
    w, h = 64, 48
    intr = oracle.make_intrinsics(w, h, 32.0, 24.0, 64.0, 64.0)
    aligned = np.zeros((h, w), np.uint32)
    aligned[10, 20] = 2048; aligned[11, 21] = 1; aligned[30, 40] = 512
    kp = np.zeros(5, oracle.KEYPOINT_DTYPE)
    kp["x"] = [20.4, 21.0, 40.0, 5.0, 19.5]
    kp["y"] = [9.6, 11.0, 30.0, 5.0, 10.49]
    kp["response"] = [30, 30, 30, 30, 1.0]
    desc = np.arange(5 * 32, dtype=np.uint8).reshape(5, 32)
    k2, d2, pts = oracle.keypoint_pixel_to_point(aligned, intr, kp, desc)
    # kp0 -> (20,10) depth 2048 kept; kp1 depth 1 (not > 1) dropped; kp2 kept; kp3 no depth; kp4 response 1.0 dropped
    assert np.array_equal(k2["x"], kp["x"][[0, 2]]) and np.array_equal(d2, desc[[0, 2]])
    x0 = np.float64((np.float32(20.4) - np.float32(32)) / np.float32(64))
    y0 = np.float64((np.float32(9.6) - np.float32(24)) / np.float32(64))
    assert np.array_equal(pts[0], [2048.0 * x0, 2048.0 * y0, 2048.0])
    assert np.array_equal(pts[1], [512.0 * 0.125, 512.0 * 0.09375, 512.0])


def test_reproject_identity_and_transform(oracle):
    intr = oracle.make_intrinsics(640, 480, 320.0, 240.0, 512.0, 512.0)
    pts = np.array([[0.25, -0.5, 2.0], [1.0, 1.0, 4.0], [0.0, 0.0, 1.0]])
    pos = oracle.reproject_points(pts, None, intr)
    assert np.array_equal(pos, np.array([[384, 112], [448, 368], [320, 240]], np.float32))
    T = np.eye(4); T[0, 3] = 0.5; T[2, 3] = 1.0  # x += 0.5, z += 1
    pos = oracle.reproject_points(pts, T, intr)
    ref = np.stack([(pts[:, 0] + 0.5) / (pts[:, 2] + 1) * 512 + 320, pts[:, 1] / (pts[:, 2] + 1) * 512 + 240], 1)
    assert np.allclose(pos, ref, rtol=0, atol=1e-3)


def test_distortion_round_trip(oracle):
    """project(MODIFIED_BROWN_CONRADY) of deproject(INVERSE_BROWN_CONRADY) with the same coefficients is the identity
    up to the model's approximation error: aligned depth through a distorted pair stays within a pixel of the
    undistorted result."""
    rng = np.random.default_rng(3)
    w, h = 80, 60
    depth = rng.integers(800, 1200, (h, w)).astype(np.uint16)
    plain = oracle.make_intrinsics(w, h, 40.0, 30.0, 70.0, 70.0)
    c = (0.05, -0.02, 0.001, -0.001, 0.003)
    di = oracle.make_intrinsics(w, h, 40.0, 30.0, 70.0, 70.0, 2, c)
    oi = oracle.make_intrinsics(w, h, 40.0, 30.0, 70.0, 70.0, 1, c)
    a = oracle.align_depth_to_other(depth, 0.001, plain, plain, oracle.make_extrinsics())
    b = oracle.align_depth_to_other(depth, 0.001, di, oi, oracle.make_extrinsics())
    inner = (slice(8, h - 8), slice(8, w - 8))
    assert (b[inner] > 0).all()
    # each output is a min over a 2x2 neighbourhood; a sub-pixel model error moves at most one contributor
    lo = np.minimum.reduce([np.roll(np.roll(depth, dy, 0), dx, 1) for dy in (-1, 0, 1, 2) for dx in (-1, 0, 1, 2)])
    hi = np.maximum.reduce([np.roll(np.roll(depth, dy, 0), dx, 1) for dy in (-1, 0, 1, 2) for dx in (-1, 0, 1, 2)])
    assert (b[inner] >= lo[inner]).all() and (b[inner] <= hi[inner]).all() and (a[inner] >= lo[inner]).all()


def test_compact_pairs(oracle):
    idx = np.array([2, -1, 0, -1, 1], np.int32)
    qp = np.arange(15, dtype=np.float64).reshape(5, 3)
    tp = 100 + np.arange(9, dtype=np.float64).reshape(3, 3)
    txy = np.array([[10.9, 20.1], [30.5, 40.5], [65535.0, 0.2]], np.float32)
    prev, curr, xs, ys = oracle.compact_pairs(idx, qp, tp, txy)
    assert np.array_equal(prev, qp[[0, 2, 4]]) and np.array_equal(curr, tp[[2, 0, 1]])
    assert xs.tolist() == [65535, 10, 30] and ys.tolist() == [0, 20, 40]


def test_rgb_to_grayscale_matches_float64_formula(oracle):
    """Independent restatement with numpy float64 (IEEE, un-fused): floor((B*0.07 + G*0.72 + R*0.21) + 0.5), over a
    frame that contains every boundary case 7B + 72G + 21R = 50 (mod 100)."""
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (64, 96, 3)).astype(np.uint8)
    tie = [(r, g, b) for r in range(0, 256, 5) for g in range(0, 256, 7) for b in range(256)
           if (7 * b + 72 * g + 21 * r) % 100 == 50][:64 * 96 // 2]
    rgb.reshape(-1, 3)[:len(tie)] = np.array(tie, np.uint8)
    R, G, B = (rgb[..., i].astype(np.float32).astype(np.float64) for i in range(3))
    ref = np.floor((B * 0.07 + G * 0.72 + R * 0.21) + 0.5).astype(np.uint8)
    assert np.array_equal(oracle.rgb_to_grayscale(rgb), ref)
    assert oracle.rgb_to_grayscale(np.full((4, 4, 3), 255, np.uint8)).max() == 255


# ---------------------------------------------------------------------------------------------------------------
# Independent restatement in vectorised numpy: every np.float32 / np.float64 ufunc is one IEEE operation, so numpy
# reproduces "each operation rounded on its own" without sharing a line of code with the C oracle.
def _np_radial(k, r2, one):
    return one + k[0] * r2 + k[1] * r2 * r2 + k[4] * r2 * r2 * r2


def _np_align(depth, scale, di, oi, ex):
    f32 = np.float32
    h, w = depth.shape
    raw = depth.astype(np.int64)
    metres = depth.astype(np.int32).astype(f32) * f32(scale)
    R = [f32(v) for v in ex.rotation]
    T = [f32(v) for v in ex.translation]
    kd = [f32(v) for v in di.coeffs]
    ko = [f32(v) for v in oi.coeffs]
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    corners = []
    for shift in (f32(-0.5), f32(0.5)):
        x = ((xs.astype(f32) + shift) - f32(di.ppx)) / f32(di.fx)
        y = ((ys.astype(f32) + shift) - f32(di.ppy)) / f32(di.fy)
        if di.model == 2:
            r2 = x * x + y * y
            f = _np_radial(kd, r2, f32(1))
            ux = x * f + f32(2) * kd[2] * x * y + kd[3] * (r2 + f32(2) * x * x)
            uy = y * f + f32(2) * kd[3] * x * y + kd[2] * (r2 + f32(2) * y * y)
            x, y = ux, uy
        p = (metres * x, metres * y, metres)
        q = [R[r] * p[0] + R[3 + r] * p[1] + R[6 + r] * p[2] + T[r] for r in range(3)]
        with np.errstate(divide="ignore", invalid="ignore"):
            a, b = q[0] / q[2], q[1] / q[2]
        if oi.model == 1:
            r2 = a * a + b * b
            f = _np_radial(ko, r2, f32(1))
            a = a * f
            b = b * f
            da = a + f32(2) * ko[2] * a * b + ko[3] * (r2 + f32(2) * a * a)
            db = b + f32(2) * ko[3] * a * b + ko[2] * (r2 + f32(2) * b * b)
            a, b = da, db
        u = a * f32(oi.fx) + f32(oi.ppx) + f32(0.5)
        v = b * f32(oi.fy) + f32(oi.ppy) + f32(0.5)
        with np.errstate(invalid="ignore"):
            corners.append((np.trunc(np.nan_to_num(u, nan=0.0)).astype(np.int64), np.trunc(np.nan_to_num(v, nan=0.0)).astype(np.int64)))
    (x0, y0), (x1, y1) = corners
    ok = (metres != 0) & (x0 >= 0) & (y0 >= 0) & (x1 < oi.width) & (y1 < oi.height)
    out = np.full(oi.width * oi.height, 2 ** 32 - 1, np.int64)
    span_x, span_y = int((x1 - x0)[ok].max(initial=-1)), int((y1 - y0)[ok].max(initial=-1))
    for oy in range(span_y + 1):
        for ox in range(span_x + 1):
            m = ok & (x0 + ox <= x1) & (y0 + oy <= y1)
            np.minimum.at(out, ((y0 + oy) * oi.width + (x0 + ox))[m], raw[m])
    out[out == 2 ** 32 - 1] = 0
    return out.reshape(oi.height, oi.width).astype(np.uint32)


@pytest.mark.parametrize("dist_depth,dist_other", [(False, False), (False, True), (True, True)])
def test_align_matches_numpy_float32(oracle, dist_depth, dist_other):
    rng = np.random.default_rng(21)
    w, h = 160, 120
    yy, xx = np.mgrid[0:h, 0:w]
    depth = (900 + 400 * np.sin(xx / 23.0) * np.cos(yy / 17.0) + rng.integers(-9, 10, (h, w))).astype(np.uint16)
    depth[rng.random((h, w)) < 0.1] = 0
    c = (0.11, -0.23, 0.0009, -0.0006, 0.07)
    di = oracle.make_intrinsics(w, h, 81.3, 58.9, 95.7, 96.1, 2 if dist_depth else 4, c if dist_depth else (0,) * 5)
    oi = oracle.make_intrinsics(w, h, 78.2, 61.4, 131.5, 132.25, 1 if dist_other else 0, c if dist_other else (0,) * 5)
    a, b = 0.01, -0.007
    ex = oracle.make_extrinsics((1, a * b, -b, 0, 1, a, b, -a, 1), (0.0148, -0.0003, 0.0011))
    got = oracle.align_depth_to_other(depth, 0.001, di, oi, ex)
    ref = _np_align(depth, 0.001, di, oi, ex)
    assert (ref > 0).mean() > 0.3
    assert np.array_equal(got, ref), f"{(got != ref).sum()} pixels differ"


def test_points_and_reprojection_match_numpy(oracle):
    rng = np.random.default_rng(8)
    w, h, n = 320, 240, 400
    c = (0.09, -0.2, 0.0011, -0.0004, 0.05)
    intr = oracle.make_intrinsics(w, h, 158.3, 121.7, 211.5, 212.75, 2, c)
    aligned = rng.integers(0, 3000, (h, w)).astype(np.uint32)
    kp = np.zeros(n, oracle.KEYPOINT_DTYPE)
    kp["x"] = rng.uniform(0, w - 1, n).astype(np.float32); kp["y"] = rng.uniform(0, h - 1, n).astype(np.float32)
    kp["response"] = rng.integers(0, 60, n)
    desc = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    k2, d2, pts = oracle.keypoint_pixel_to_point(aligned, intr, kp, desc)
    xi = (kp["x"].astype(np.float64) + 0.5).astype(np.int64); yi = (kp["y"].astype(np.float64) + 0.5).astype(np.int64)
    dep = aligned[yi, xi].astype(np.int64)
    keep = (dep > 1) & (kp["response"] > 1.0)
    assert np.array_equal(k2["x"], kp["x"][keep]) and np.array_equal(d2, desc[keep])
    f32, f64 = np.float32, np.float64
    kc = [f64(f32(v)) for v in c]
    x = ((kp["x"][keep] - f32(intr.ppx)) / f32(intr.fx)).astype(f64)
    y = ((kp["y"][keep] - f32(intr.ppy)) / f32(intr.fy)).astype(f64)
    r2 = x * x + y * y
    f = _np_radial(kc, r2, f64(1))
    ux = x * f + 2 * kc[2] * x * y + kc[3] * (r2 + 2 * x * x)
    uy = y * f + 2 * kc[3] * x * y + kc[2] * (r2 + 2 * y * y)
    z = dep[keep].astype(f32).astype(f64)
    assert np.array_equal(pts, np.stack([z * ux, z * uy, z], 1))
    # reprojection through a forward-distorted camera with a rigid transform
    cam = oracle.make_intrinsics(w, h, 158.3, 121.7, 211.5, 212.75, 1, c)
    T = np.eye(4); T[:3, :3] = [[1, -0.01, 0.02], [0.01, 1, -0.015], [-0.02, 0.015, 1]]; T[:3, 3] = [4.0, -2.5, 7.0]
    pos = oracle.reproject_points(pts, T, cam)
    e = [(T[r, 0] * pts[:, 0] + T[r, 1] * pts[:, 1]) + (T[r, 2] * pts[:, 2] + T[r, 3] * 1.0) for r in range(3)]
    a = (e[0] / e[2]).astype(f32); b = (e[1] / e[2]).astype(f32)
    k32 = [f32(v) for v in c]
    r2 = a * a + b * b
    f = _np_radial(k32, r2, f32(1))
    a = a * f; b = b * f
    da = a + f32(2) * k32[2] * a * b + k32[3] * (r2 + f32(2) * a * a)
    db = b + f32(2) * k32[3] * a * b + k32[2] * (r2 + f32(2) * b * b)
    ref = np.stack([da * f32(cam.fx) + f32(cam.ppx), db * f32(cam.fy) + f32(cam.ppy)], 1)
    assert np.array_equal(pos.view(np.uint32), ref.astype(f32).view(np.uint32))


def _np_search_by_projection(q_desc, q_uv, q_kp, t_desc, t_kp, sf, th, th_high, check):
    """Vectorised numpy restatement of ORBmatcher::SearchByProjection's front-end gates + ComputeThreeMaxima."""
    nq = len(q_desc)
    lut = np.array([bin(i).count("1") for i in range(256)], np.int32)
    dist = lut[q_desc[:, None, :] ^ t_desc[None, :, :]].sum(2)
    radius = (np.float32(th) * sf[np.clip(q_kp["octave"], 0, len(sf) - 1)]).astype(np.float32)
    dx = np.abs(t_kp["x"][None, :] - q_uv[:, 0:1]); dy = np.abs(t_kp["y"][None, :] - q_uv[:, 1:2])
    oq = np.clip(q_kp["octave"], 0, len(sf) - 1)[:, None]; ot = t_kp["octave"][None, :]
    ok = (dx < radius[:, None]) & (dy < radius[:, None]) & (ot >= oq - 1) & (ot <= oq + 1)
    d = np.where(ok, dist, 10 ** 6)
    idx = d.argmin(1).astype(np.int32)  # first minimum = lowest train index
    best = d[np.arange(nq), idx]
    good = best <= th_high
    idx = np.where(good, idx, -1).astype(np.int32); best = np.where(good, best, -1).astype(np.int32)
    if check:
        rot = (q_kp["angle"] - t_kp["angle"][np.maximum(idx, 0)]).astype(np.float32)
        rot = np.where(rot < 0, rot + np.float32(360), rot).astype(np.float32)
        x = (rot * np.float32(1.0 / 30)).astype(np.float64)
        bins = (np.sign(x) * np.floor(np.abs(x) + 0.5)).astype(np.int64)  # C round(): half away from zero
        bins[bins == 30] = 0
        hist = np.bincount(bins[idx >= 0], minlength=30)
        m = [0, 0, 0]; ind = [-1, -1, -1]
        for b in range(30):
            s = int(hist[b])
            if s > m[0]:
                m = [s, m[0], m[1]]; ind = [b, ind[0], ind[1]]
            elif s > m[1]:
                m = [m[0], s, m[1]]; ind = [ind[0], b, ind[1]]
            elif s > m[2]:
                m[2] = s; ind[2] = b
        if np.float32(m[1]) < np.float32(0.1) * np.float32(m[0]):
            ind[1] = ind[2] = -1
        elif np.float32(m[2]) < np.float32(0.1) * np.float32(m[0]):
            ind[2] = -1
        drop = (idx >= 0) & ~np.isin(bins, [i for i in ind if i >= 0])
        idx[drop] = -1; best[drop] = -1
    return idx, best, int((idx >= 0).sum())


@pytest.mark.parametrize("check", [False, True])
def test_search_by_projection_matches_numpy(oracle, check):
    rng = np.random.default_rng(17)
    nq, nt, nl = 300, 350, 8
    sf = (np.float32(1.2) ** np.arange(nl)).astype(np.float32)
    t_kp = np.zeros(nt, oracle.KEYPOINT_DTYPE)
    t_kp["x"] = rng.uniform(20, 620, nt).astype(np.float32); t_kp["y"] = rng.uniform(20, 460, nt).astype(np.float32)
    t_kp["octave"] = rng.integers(0, nl, nt); t_kp["angle"] = rng.uniform(0, 360, nt).astype(np.float32)
    t_desc = rng.integers(0, 256, (nt, 32)).astype(np.uint8)
    src = rng.integers(0, nt, nq)  # every query is a noisy copy of some train keypoint
    q_kp = t_kp[src].copy()
    q_kp["octave"] = np.clip(q_kp["octave"] + rng.integers(-2, 3, nq), 0, nl - 1)
    q_kp["angle"] = np.mod(q_kp["angle"] + np.where(rng.random(nq) < 0.7, 12.0, rng.uniform(0, 360, nq)), 360).astype(np.float32)
    q_uv = np.stack([t_kp["x"][src] + rng.normal(0, 6, nq), t_kp["y"][src] + rng.normal(0, 6, nq)], 1).astype(np.float32)
    q_desc = t_desc[src] ^ (rng.random((nq, 32)) < 0.25).astype(np.uint8) * rng.integers(0, 256, (nq, 32)).astype(np.uint8)
    got = oracle.search_by_projection(q_desc, q_uv, q_kp, t_desc, t_kp, sf, 7.0, 100, check)
    ref = _np_search_by_projection(q_desc, q_uv, q_kp, t_desc, t_kp, sf, 7.0, 100, check)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
    assert 40 < got[2] < nq
    if check:
        assert got[2] < oracle.search_by_projection(q_desc, q_uv, q_kp, t_desc, t_kp, sf, 7.0, 100, False)[2]


def _py_stereo(ol, orr, kl, dl, kr, dr, mbf, fx):
    """Frame::ComputeStereoMatches restated with numpy float32 scalars / vector ops, independent of the C oracle
    (which only supplies the two image pyramids)."""
    f32 = np.float32
    sf, isf = ol.scale, ol.inv_scale
    lut = np.array([bin(i).count("1") for i in range(256)], np.int32)
    mb = f32(mbf) / f32(fx); maxD = f32(mbf) / mb; minD = f32(0)
    r = (f32(2.0) * sf[kr["octave"]]).astype(f32)
    maxr = np.ceil((kr["y"] + r).astype(f32)).astype(np.int64); minr = np.floor((kr["y"] - r).astype(f32)).astype(np.int64)
    ur = np.full(len(kl), -1, f32); z = np.full(len(kl), -1, f32)
    found = []
    for iL in range(len(kl)):
        uL, vL, lev = kl["x"][iL], kl["y"][iL], int(kl["octave"][iL])
        row = int(vL)
        minU, maxU = f32(uL - maxD), f32(uL - minD)
        if maxU < 0:
            continue
        cand = np.nonzero((minr <= row) & (row <= maxr) & (kr["octave"] >= lev - 1) & (kr["octave"] <= lev + 1) &
                          (kr["x"] >= minU) & (kr["x"] <= maxU))[0]
        if len(cand) == 0:
            continue
        d = lut[dl[iL][None, :] ^ dr[cand]].sum(1)
        j = int(d.argmin())
        if not (d[j] < 100 and d[j] < 75):
            continue
        uR0 = kr["x"][cand[j]]
        su, sv, sr = (f32(np.sign(v) * np.floor(np.abs(np.float64(f32(a * isf[lev]))) + 0.5)) for v, a in
                      ((uL, uL), (vL, vL), (uR0, uR0)))  # C round(): half away from zero
        imL = ol.level_padded(lev).astype(np.int64); imR = orr.level_padded(lev).astype(np.int64)
        cols = int(ol.lw[lev])
        if sr + 5 - 5 < 0 or sr + 5 + 5 + 1 >= cols:
            continue
        xl, yl, xr = int(su) + 19, int(sv) + 19, int(sr) + 19
        IL = imL[yl - 5:yl + 6, xl - 5:xl + 6] - imL[yl, xl]
        dists = []
        for inc in range(-5, 6):
            IR = imR[yl - 5:yl + 6, xr + inc - 5:xr + inc + 6] - imR[yl, xr + inc]
            dists.append(int(np.abs(IL - IR).sum()))
        b = int(np.argmin(dists))  # first minimum, like the strict '<' scan
        if b == 0 or b == 10:
            continue
        d1, d2, d3 = f32(dists[b - 1]), f32(dists[b]), f32(dists[b + 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
        if delta < -1 or delta > 1:
            continue
        best_u = f32(sf[lev] * f32(f32(sr + f32(b - 5)) + delta))
        disp = f32(uL - best_u)
        if disp >= minD and disp < maxD:
            if disp <= 0:
                disp = f32(0.01); best_u = f32(np.float64(uL) - 0.01)
            z[iL] = f32(f32(mbf) / disp); ur[iL] = best_u
            found.append((dists[b], iL))
    if found:
        found.sort()
        th = f32(f32(f32(1.5) * f32(1.4)) * f32(found[len(found) // 2][0]))
        for dist, iL in found:
            if not (f32(dist) < th):
                ur[iL] = -1; z[iL] = -1
    return ur, z, int((ur >= 0).sum() if not found else sum(1 for dist, _ in found if f32(dist) < th))


def test_stereo_matches_numpy_restatement_and_known_disparity(oracle, synth):
    w, h = 320, 240
    left = synth.textured_frame(w, h, 21)
    right = synth.shifted_frame(left, -9, 0, 5)  # every scene point appears 9 px further left in the right image
    ol, orr = oracle.Oracle(w, h, 400, 1.2, 6), oracle.Oracle(w, h, 400, 1.2, 6)
    kl, dl = ol.extract(left); kr, dr = orr.extract(right)
    bf, fx = 18.0, 200.0
    ur, z, n = oracle.compute_stereo_matches(ol, orr, kl, dl, kr, dr, bf, fx)
    rur, rz, rn = _py_stereo(ol, orr, kl, dl, kr, dr, bf, fx)
    assert n == rn and np.array_equal(ur.view(np.uint32), rur.view(np.uint32)) and np.array_equal(z.view(np.uint32), rz.view(np.uint32))
    m = ur >= 0
    assert m.sum() == n and n > 100
    disp = kl["x"][m] - ur[m]
    assert np.abs(np.median(disp) - 9.0) < 0.1 and (np.abs(disp - 9.0) < 1.5).mean() > 0.9
    assert np.allclose(z[m], bf / disp, rtol=1e-6)

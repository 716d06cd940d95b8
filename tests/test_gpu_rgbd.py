"""GPU parity, RGB-D association stages: align_depth_to_other, keypoint_pixel_to_point,
reprojection and the batched windowed matcher with pair compaction, through the C ABI, bit-exact against
oracle/rgbd_oracle.c (integers, float32 and float64 alike: every operation is rounded on its own on both sides)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orbb():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import __graft_entry__ as g
    g.build()
    return importlib.import_module("jetracer-orbslam2_b200.orbb")


def synth_depth(w, h, seed):
    """Smooth surface 0.4-4 m in 1 mm units + steps + speckle + holes (zeros), like a D435 depth frame."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    z = 1500 + 900 * np.sin(xx / 97.0 + seed) * np.cos(yy / 61.0) + 3.0 * xx
    for _ in range(12):
        x0, y0 = rng.integers(0, w - 40), rng.integers(0, h - 40)
        z[y0:y0 + rng.integers(10, 120), x0:x0 + rng.integers(10, 160)] = rng.integers(400, 4000)
    z += rng.integers(-8, 9, (h, w))
    d = np.clip(z, 0, 65535).astype(np.uint16)
    d[rng.random((h, w)) < 0.08] = 0
    d[:, :6] = 0
    return d


def d435_pair(orbb_or_oracle, w, h, distorted):
    mk_i, mk_e = orbb_or_oracle.make_intrinsics, orbb_or_oracle.make_extrinsics
    di = mk_i(w, h, w * 0.5 + 3.7, h * 0.5 - 2.2, 0.502 * w, 0.502 * w, 4)  # BROWN_CONRADY, zero coeffs (D435 depth)
    if distorted:
        oi = mk_i(w, h, w * 0.5 - 5.1, h * 0.5 + 4.3, 0.72 * w, 0.725 * w, 1, (0.12, -0.25, 0.0007, -0.0004, 0.09))
    else:
        oi = mk_i(w, h, w * 0.5 - 5.1, h * 0.5 + 4.3, 0.72 * w, 0.725 * w, 2, (0, 0, 0, 0, 0))
    a, b = 0.004, -0.003  # small rotation about x and y, column-major
    R = (1, a * b, -b, 0, 1, a, b, -a, 1)
    ex = mk_e(R, (0.0148, 0.0002, 0.0003))
    return di, oi, ex


@pytest.mark.parametrize("w,h,distorted", [(848, 480, False), (848, 480, True), (640, 480, True), (333, 251, True)])
def test_align_depth_to_other(orbb, oracle, w, h, distorted):
    import torch
    n = 3
    depth = np.stack([synth_depth(w, h, 10 + i) for i in range(n)])
    ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=w, height=h, max_batch=1)
    di, oi, e = d435_pair(orbb, w, h, distorted)
    odi, ooi, oe = d435_pair(oracle, w, h, distorted)
    d_depth = torch.from_numpy(depth.view(np.int16)).cuda()
    d_out = torch.zeros((n, h, w), dtype=torch.int32, device="cuda")
    ex.align_depth_to_other(d_depth, n, 0.001, di, oi, e, d_out, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint32)
    for i in range(n):
        ref = oracle.align_depth_to_other(depth[i], 0.001, odi, ooi, oe)
        assert (ref > 0).mean() > 0.3
        assert np.array_equal(got[i], ref), f"frame {i}: {(got[i] != ref).sum()} pixels differ"


def test_align_different_sizes_and_empty(orbb, oracle):
    """depth 424x240 onto a 640x480 image (rectangles of ~2x3 pixels), plus an all-zero frame."""
    import torch
    dw, dh, ow, oh = 424, 240, 640, 480
    depth = np.stack([synth_depth(dw, dh, 77), np.zeros((dh, dw), np.uint16)])
    di = orbb.make_intrinsics(dw, dh, 212.3, 119.1, 213.0, 213.0, 4)
    oi = orbb.make_intrinsics(ow, oh, 322.0, 241.5, 460.0, 461.0, 1, (0.1, -0.2, 0.001, 0.0005, 0.05))
    e = orbb.make_extrinsics((1, 0, 0, 0, 1, 0, 0, 0, 1), (0.015, 0, 0))
    odi = oracle.make_intrinsics(dw, dh, 212.3, 119.1, 213.0, 213.0, 4)
    ooi = oracle.make_intrinsics(ow, oh, 322.0, 241.5, 460.0, 461.0, 1, (0.1, -0.2, 0.001, 0.0005, 0.05))
    oe = oracle.make_extrinsics((1, 0, 0, 0, 1, 0, 0, 0, 1), (0.015, 0, 0))
    ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=ow, height=oh, max_batch=1)
    d_depth = torch.from_numpy(depth.view(np.int16)).cuda()
    d_out = torch.full((2, oh, ow), 5, dtype=torch.int32, device="cuda")
    ex.align_depth_to_other(d_depth, 2, 0.001, di, oi, e, d_out, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[0], oracle.align_depth_to_other(depth[0], 0.001, odi, ooi, oe))
    assert not got[1].any()
    # unsupported models are refused loudly, not approximated
    bad = orbb.make_intrinsics(ow, oh, 322.0, 241.5, 460.0, 461.0, orbb.DISTORTION_FTHETA)
    with pytest.raises(orbb.OrbbError):
        ex.align_depth_to_other(d_depth, 2, 0.001, di, bad, e, d_out)


def _frame_pipeline(orbb, oracle, synth, w, h, n, nfeat, distorted, seed):
    """extract -> align -> keypoint_pixel_to_point on the GPU; returns host copies + the oracle's inputs."""
    import torch
    frames = np.stack([synth.textured_frame(w, h, seed + i) for i in range(n)])
    depth = np.stack([synth_depth(w, h, seed + 50 + i) for i in range(n)])
    ex = orbb.ORBextractor(nfeat, 1.2, 8, 20, 7, width=w, height=h, max_batch=n)
    mk = ex.max_kp
    st = torch.cuda.current_stream()
    d_frames = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((n, mk, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((n, mk, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_frames, n, d_kp, d_desc, d_counts, stream=st)
    di, oi, e = d435_pair(orbb, w, h, distorted)
    if distorted:  # the keypoint image must be deprojectable: inverse model on the colour side
        oi = orbb.make_intrinsics(w, h, oi.ppx, oi.ppy, oi.fx, oi.fy, 2, (0.12, -0.25, 0.0007, -0.0004, 0.09))
    d_depth = torch.from_numpy(depth.view(np.int16)).cuda()
    d_al = torch.zeros((n, h, w), dtype=torch.int32, device="cuda")
    ex.align_depth_to_other(d_depth, n, 0.001, di, oi, e, d_al, stream=st)
    d_kp2 = torch.zeros_like(d_kp); d_desc2 = torch.zeros_like(d_desc)
    d_pts = torch.zeros((n, mk, 3), dtype=torch.float64, device="cuda")
    d_valid = torch.zeros(n, dtype=torch.int32, device="cuda")
    ex.keypoint_pixel_to_point(d_al, oi, n, d_kp, d_desc, d_counts, d_kp2, d_desc2, d_pts, d_valid, stream=st)
    torch.cuda.synchronize()
    ooi = oracle.make_intrinsics(w, h, oi.ppx, oi.ppy, oi.fx, oi.fy, oi.model, tuple(oi.coeffs))
    dev = dict(ex=ex, d_kp2=d_kp2, d_desc2=d_desc2, d_pts=d_pts, d_valid=d_valid, oi=oi)
    host = dict(kp=d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(n, mk), desc=d_desc.cpu().numpy(),
                counts=d_counts.cpu().numpy(), aligned=d_al.cpu().numpy().view(np.uint32),
                kp2=d_kp2.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(n, mk), desc2=d_desc2.cpu().numpy(),
                pts=d_pts.cpu().numpy(), valid=d_valid.cpu().numpy(), ooi=ooi)
    return dev, host


@pytest.mark.parametrize("distorted", [False, True])
def test_keypoint_pixel_to_point(orbb, oracle, synth, distorted):
    n = 4
    dev, hst = _frame_pipeline(orbb, oracle, synth, 640, 480, n, 1000, distorted, 9100)
    for f in range(n):
        c = int(hst["counts"][f])
        okp, odesc, opts = oracle.keypoint_pixel_to_point(hst["aligned"][f], hst["ooi"], hst["kp"][f, :c], hst["desc"][f, :c])
        m = int(hst["valid"][f])
        assert m == len(okp) and 0.5 * c < m < c  # holes in the depth drop some keypoints, not all
        assert np.array_equal(hst["kp2"][f, :m].view(np.uint8), okp.view(np.uint8))
        assert np.array_equal(hst["desc2"][f, :m], odesc)
        assert np.array_equal(hst["pts"][f, :m].view(np.uint64), opts.view(np.uint64)), "float64 points differ"


@pytest.mark.parametrize("with_T", [False, True])
def test_reproject_match_compact(orbb, oracle, synth, with_T):
    """Frame f is matched against frame f+1 of a slowly translating sequence: previous points reprojected with T,
    windowed match (reference gate arguments), matched 3-D pairs compacted in query order."""
    import torch
    n, w, h = 4, 640, 480
    dev, hst = _frame_pipeline(orbb, oracle, synth, w, h, n, 800, False, 9300)
    ex, mk = dev["ex"], dev["ex"].max_kp
    st = torch.cuda.current_stream()
    # previous frame = frame f, current = frame (f+1) % n (different random textures: matches are rare, so loosen the
    # gates enough to get a few hundred candidates through the descriptor compare)
    perm = [(f + 1) % n for f in range(n)]
    cur_kp = dev["d_kp2"][perm].contiguous(); cur_desc = dev["d_desc2"][perm].contiguous()
    cur_pts = dev["d_pts"][perm].contiguous(); cur_valid = dev["d_valid"][perm].contiguous()
    T = None
    if with_T:
        rng = np.random.default_rng(5)
        Ts = []
        for f in range(n):
            a = rng.normal(0, 0.01, 3)
            R = np.array([[1, -a[2], a[1]], [a[2], 1, -a[0]], [-a[1], a[0], 1]])
            M = np.eye(4); M[:3, :3] = R; M[:3, 3] = rng.normal(0, 8.0, 3)
            Ts.append(M)
        T = torch.from_numpy(np.stack([M.T.copy() for M in Ts])).cuda()  # column-major per frame
    d_pos = torch.zeros((n, mk, 2), dtype=torch.float32, device="cuda")
    ex.reproject_points(dev["d_pts"], dev["d_valid"], n, T, dev["oi"], d_pos, stream=st)
    d_idx = torch.full((n, mk), -9, dtype=torch.int32, device="cuda"); d_dist = torch.full_like(d_idx, -9)
    d_prev = torch.zeros((n, mk, 3), dtype=torch.float64, device="cuda"); d_curr = torch.zeros_like(d_prev)
    d_xy = torch.zeros((n, 2, mk), dtype=torch.int16, device="cuda")
    d_nm = torch.zeros(n, dtype=torch.int32, device="cuda")
    ex.match_keypoints_windowed_batch(dev["d_desc2"], d_pos, dev["d_valid"], cur_desc, cur_kp, 28, cur_valid, n, 40.0, 90,
                                      d_idx, d_dist, dev["d_pts"], cur_pts, d_prev, d_curr, d_xy, d_nm, stream=st)
    torch.cuda.synchronize()
    pos = d_pos.cpu().numpy(); idx = d_idx.cpu().numpy(); dist = d_dist.cpu().numpy()
    prev = d_prev.cpu().numpy(); curr = d_curr.cpu().numpy(); xy = d_xy.cpu().numpy().view(np.uint16); nm = d_nm.cpu().numpy()
    total = 0
    for f in range(n):
        m, g = int(hst["valid"][f]), perm[f]
        mc = int(hst["valid"][g])
        opos = oracle.reproject_points(hst["pts"][f, :m], Ts[f] if with_T else None, hst["ooi"])
        assert np.array_equal(pos[f, :m].view(np.uint32), opos.view(np.uint32)), "reprojected positions differ"
        txy = np.stack([hst["kp2"][g, :mc]["x"], hst["kp2"][g, :mc]["y"]], 1)
        oidx, odist, onm = oracle.match_windowed(hst["desc2"][f, :m], opos, hst["desc2"][g, :mc], txy, 40.0, 90)
        assert np.array_equal(idx[f, :m], oidx) and np.array_equal(dist[f, :m], odist)
        oprev, ocurr, oxs, oys = oracle.compact_pairs(oidx, hst["pts"][f, :m], hst["pts"][g, :mc], txy)
        assert int(nm[f]) == onm == len(oprev)
        assert np.array_equal(prev[f, :onm], oprev) and np.array_equal(curr[f, :onm], ocurr)
        assert np.array_equal(xy[f, 0, :onm], oxs) and np.array_equal(xy[f, 1, :onm], oys)
        total += onm
    assert total > 50, "test too weak: hardly any match went through"


@pytest.mark.parametrize("plan,max_batch,graph,fused", [
    ((3, 1, 4), 4, "1", "1"),    # small batches: every submit replays a captured graph (partial and full-block D2H forms)
    ((3, 1, 4), 4, "0", "1"),    # the same through the streamed path
    ((1,) * 8, 1, "1", "1"),     # the reference's operating mode: one frame per wake-up (gate + lift + match in one launch)
    ((1,) * 8, 1, "1", "0"),     # the same with the lone frame's tail as separate launches
    ((3, 1, 4), 4, "1", "0"),
    ((5, 1, 2), 8, "1", "1"),    # streamed batch -> graph -> graph: the hand-over between the two paths, carry row 5 -> 1
    ((1, 6, 1), 8, "1", "1"),    # graph -> streamed -> graph
])
def test_rgbd_frame_stage_sequence(orbb, oracle, synth, plan, max_batch, graph, fused, monkeypatch):
    """The host stage over a moving sequence of 8 frames fed as the batches of `plan` (two in flight):
    every frame's gated keypoints / descriptors / 3-D points and its matches against the previous frame -- across
    batch boundaries -- equal the oracle chain run frame by frame on the GPU extractor's raw output."""
    import torch
    monkeypatch.setenv("ORBB_STAGE_GRAPH", graph)
    monkeypatch.setenv("ORBB_STAGE_FUSED", fused)
    w, h, nfeat = 640, 480, 600
    base = synth.textured_frame(w, h, 4242)
    gray = [base]
    for i in range(7):
        gray.append(synth.shifted_frame(gray[-1], 2, 1, 100 + i))
    gray = np.stack(gray)
    depth = np.stack([synth_depth(w, h, 300 + i // 3) for i in range(8)])
    di, oi, e = d435_pair(orbb, w, h, False)
    odi, ooi, oe = d435_pair(oracle, w, h, False)
    params = orbb.Params(nfeat, 1.2, 8, 20, 7)
    stage = orbb.RgbdFrameStage(params, di, oi, e, depth_scale=0.001, max_pixel_distance=6.0, max_hamming_distance=60,
                                max_batch=max_batch)
    rng = np.random.default_rng(11)
    T = np.tile(np.eye(4), (8, 1, 1))
    T[:, :3, 3] = rng.normal(0, 1.5, (8, 3))  # millimetres: depth units are raw
    assert sum(plan) == 8
    results, pending, f0 = [], None, 0
    for n in plan:  # submit batch k + 1 before waiting for batch k
        t = stage.submit(gray[f0:f0 + n], depth[f0:f0 + n], T[f0:f0 + n])
        if pending is not None:
            results.append({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in stage.wait(pending).items()})
        pending, f0 = t, f0 + n
    results.append({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in stage.wait(pending).items()})
    got = {k: np.concatenate([r[k] for r in results]) for k in results[0] if isinstance(results[0][k], np.ndarray)}

    ex = orbb.ORBextractor(nfeat, 1.2, 8, 20, 7, width=w, height=h, max_batch=8)
    kp, desc, counts = ex.extract_batch(gray)  # raw GPU output, GPU order (deterministic)
    assert np.array_equal(got["keypoints_count"], counts)
    prev = None
    total_matched = 0
    for f in range(8):
        c = int(counts[f])
        aligned = oracle.align_depth_to_other(depth[f], 0.001, odi, ooi, oe)
        okp, odesc, opts = oracle.keypoint_pixel_to_point(aligned, ooi, kp[f, :c], desc[f, :c])
        m = len(okp)
        assert int(got["valid_keypoints_num"][f]) == m
        assert np.array_equal(got["keypoints"][f, :m].view(np.uint8), okp.view(np.uint8))
        assert np.array_equal(got["descriptors"][f, :m], odesc)
        assert np.array_equal(got["points"][f, :m], opts)
        if prev is None:
            assert int(got["matched_keypoints_num"][f]) == 0
        else:
            pkp, pdesc, ppts = prev
            pos = oracle.reproject_points(ppts, T[f], ooi)
            txy = np.stack([okp["x"], okp["y"]], 1)
            oidx, odist, onm = oracle.match_windowed(pdesc, pos, odesc, txy, 6.0, 60)
            oprev, ocurr, oxs, oys = oracle.compact_pairs(oidx, ppts, opts, txy)
            assert int(got["matched_keypoints_num"][f]) == onm, f"frame {f}"
            assert np.array_equal(got["previous_matched_points"][f, :onm], oprev)
            assert np.array_equal(got["current_matched_points"][f, :onm], ocurr)
            assert np.array_equal(got["matched_xy"][f, 0, :onm], oxs) and np.array_equal(got["matched_xy"][f, 1, :onm], oys)
            total_matched += onm
        prev = (okp, odesc, opts)
    assert total_matched > 300, f"sequence too weak: {total_matched} matches"
    # reset drops the carried frame
    stage.reset()
    r = stage.wait(stage.submit(gray[0:1], depth[0:1]))
    assert int(r["matched_keypoints_num"][0]) == 0
    if max_batch < 8:
        with pytest.raises(orbb.OrbbError):
            stage.submit(gray[0:max_batch + 1], depth[0:max_batch + 1])  # over the stage's batch capacity


@pytest.mark.parametrize("w,h", [(848, 480), (333, 251), (5, 3)])
def test_rgb_to_grayscale(orbb, oracle, w, h):
    import torch
    rng = np.random.default_rng(w)
    n = 3
    rgb = rng.integers(0, 256, (n, h, w, 3)).astype(np.uint8)
    tie = [(r, g, b) for r in range(0, 256, 3) for g in range(0, 256, 5) for b in range(256)
           if (7 * b + 72 * g + 21 * r) % 100 == 50][:w * h // 2]
    rgb[0].reshape(-1, 3)[:len(tie)] = np.array(tie, np.uint8)
    ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=640, height=480, max_batch=1)
    d_rgb = torch.from_numpy(rgb).cuda()
    d_gray = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
    ex.rgb_to_grayscale(d_rgb, n, d_gray, width=w, height=h, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    got = d_gray.cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], oracle.rgb_to_grayscale(rgb[i])), f"frame {i}"


@pytest.mark.parametrize("check", [False, True])
def test_match_projection_batch(orbb, oracle, synth, check):
    """ORB-SLAM2 SearchByProjection gates on real keypoints of a translating sequence: frame f -> frame f+1, projected
    positions = reprojected 3-D points; per-frame results equal the oracle (idx, dist, survivor count)."""
    import torch
    n, w, h = 4, 640, 480
    base = synth.textured_frame(w, h, 777)
    gray = [base]
    for i in range(n):
        gray.append(synth.shifted_frame(gray[-1], 3, 2, 900 + i))
    gray = np.stack(gray)
    ex = orbb.ORBextractor(800, 1.2, 8, 20, 7, width=w, height=h, max_batch=n + 1)
    mk = ex.max_kp
    st = torch.cuda.current_stream()
    d_frames = torch.from_numpy(gray).cuda()
    d_kp = torch.zeros((n + 1, mk, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((n + 1, mk, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_frames, n + 1, d_kp, d_desc, d_cnt, stream=st)
    # projected positions: the previous keypoint moved by the known shift + a little noise (stands in for the
    # reprojection of its 3-D point)
    torch.cuda.synchronize()
    kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(n + 1, mk)
    desc = d_desc.cpu().numpy(); cnt = d_cnt.cpu().numpy()
    rng = np.random.default_rng(3)
    uv = np.zeros((n, mk, 2), np.float32)
    uv[..., 0] = kp[:n]["x"] + 3 + rng.normal(0, 1.5, (n, mk)); uv[..., 1] = kp[:n]["y"] + 2 + rng.normal(0, 1.5, (n, mk))
    d_uv = torch.from_numpy(uv).cuda()
    q_kp, q_desc, q_cnt = d_kp[:n].contiguous(), d_desc[:n].contiguous(), d_cnt[:n].contiguous()
    t_kp, t_desc, t_cnt = d_kp[1:].contiguous(), d_desc[1:].contiguous(), d_cnt[1:].contiguous()
    d_idx = torch.full((n, mk), -9, dtype=torch.int32, device="cuda"); d_dist = torch.full_like(d_idx, -9)
    d_nm = torch.zeros(n, dtype=torch.int32, device="cuda")
    ex.match_keypoints_projection_batch(q_desc, d_uv, q_kp, q_cnt, t_desc, t_kp, t_cnt, n, 7.0, 100, check, d_idx, d_dist,
                                        d_nm, stream=st)
    torch.cuda.synchronize()
    idx, dist, nm = d_idx.cpu().numpy(), d_dist.cpu().numpy(), d_nm.cpu().numpy()
    sf = ex.GetScaleFactors()
    total = 0
    for f in range(n):
        a, b = int(cnt[f]), int(cnt[f + 1])
        oidx, odist, onm = oracle.search_by_projection(desc[f, :a], uv[f, :a], kp[f, :a], desc[f + 1, :b], kp[f + 1, :b], sf,
                                                       7.0, 100, check)
        assert np.array_equal(idx[f, :a], oidx) and np.array_equal(dist[f, :a], odist) and int(nm[f]) == onm, f"frame {f}"
        total += onm
    assert total > 400, f"test too weak: {total} matches"


def test_compute_stereo_matches(orbb, oracle, synth):
    """Frame::ComputeStereoMatches over a batch of rectified pairs (left = frame 2p, right = 2p+1): uRight, depth and
    the surviving-match count are bit-exact against the oracle fed with the GPU extractor's own keypoint order."""
    import torch
    w, h, npairs = 640, 480, 3
    frames = []
    for p in range(npairs):
        left = synth.textured_frame(w, h, 40 + p)
        frames += [left, synth.shifted_frame(left, -(7 + 5 * p), 0, 60 + p)]
    frames = np.stack(frames)
    ex = orbb.ORBextractor(1000, 1.2, 8, 20, 7, width=w, height=h, max_batch=2 * npairs)
    mk = ex.max_kp
    st = torch.cuda.current_stream()
    d_in = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((2 * npairs, mk, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((2 * npairs, mk, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2 * npairs, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_in, 2 * npairs, d_kp, d_desc, d_cnt, stream=st)
    d_ur = torch.full((npairs, mk), -7.0, dtype=torch.float32, device="cuda"); d_z = torch.full_like(d_ur, -7.0)
    d_ns = torch.zeros(npairs, dtype=torch.int32, device="cuda")
    bf, fx = 40.0, 400.0
    ex.compute_stereo_matches(d_kp, d_desc, d_cnt, npairs, bf, fx, d_ur, d_z, d_ns, stream=st)
    torch.cuda.synchronize()
    kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(2 * npairs, mk); desc = d_desc.cpu().numpy(); cnt = d_cnt.cpu().numpy()
    ur, z, ns = d_ur.cpu().numpy(), d_z.cpu().numpy(), d_ns.cpu().numpy()
    for p in range(npairs):
        ol, orr = oracle.Oracle(w, h, 1000), oracle.Oracle(w, h, 1000)
        ol.compute_pyramid(frames[2 * p]); orr.compute_pyramid(frames[2 * p + 1])
        a, b = int(cnt[2 * p]), int(cnt[2 * p + 1])
        our, oz, on = oracle.compute_stereo_matches(ol, orr, kp[2 * p, :a], desc[2 * p, :a], kp[2 * p + 1, :b], desc[2 * p + 1, :b], bf, fx)
        assert int(ns[p]) == on and on > 300, (p, int(ns[p]), on)
        assert np.array_equal(ur[p, :a].view(np.uint32), our.view(np.uint32)), f"pair {p}: uRight"
        assert np.array_equal(z[p, :a].view(np.uint32), oz.view(np.uint32)), f"pair {p}: depth"
        m = our >= 0
        assert abs(np.median(kp[2 * p, :a]["x"][m] - our[m]) - (7 + 5 * p)) < 0.15
    with pytest.raises(orbb.OrbbError):
        ex.compute_stereo_matches(d_kp, d_desc, d_cnt, npairs + 1, bf, fx, d_ur, d_z)  # more pairs than resident frames


def test_jpeg_preview(orbb, synth):
    """SURVEY 8f-4, the nvJPEG preview of the reference (buildStream.cpp:491-521, 613-621; overlay kernel
    post_processing.cu:45-70): the three planes handed to the encoder equal a numpy restatement of "gray x 3, 2x2 block
    per keypoint set to 255 in G" exactly; the JPEG decodes (cv2) to that image within JPEG error; the keypoint count
    can live on the device; the bytes slot into the frame message."""
    import torch
    cv2 = pytest.importorskip("cv2")
    w, h = 848, 480
    img = synth.textured_frame(w, h, 9900)
    ex = orbb.ORBextractor(400, 1.2, 4, 20, 7, width=w, height=h, max_batch=1)
    d_in = torch.from_numpy(img[None]).cuda()
    d_kp = torch.zeros((1, ex.max_kp, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((1, ex.max_kp, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    ex.extract_batch_device(d_in, 1, d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    n = int(d_cnt.item())
    kp = d_kp.cpu().numpy().reshape(-1, 7)[:n]
    pv = orbb.Preview(w, h, 90)
    # expected planes: reference loop bounds for (int x = pos.x - 1; x < pos.x + 1; x++), clamped to the image
    exp = np.stack([img, img.copy(), img])
    for x, y in kp[:, :2]:
        xs = [v for v in range(int(np.float32(x) - np.float32(1)), int(np.ceil(np.float32(x) + np.float32(1)))) if v < np.float32(x) + np.float32(1)]
        ys = [v for v in range(int(np.float32(y) - np.float32(1)), int(np.ceil(np.float32(y) + np.float32(1)))) if v < np.float32(y) + np.float32(1)]
        for xx in xs:
            for yy in ys:
                if 0 <= xx < w and 0 <= yy < h:
                    exp[1, yy, xx] = 255
    planes = pv.debug_planes(d_in[0], d_kp, 28, ex.max_kp, d_cnt)  # count read on the device
    assert np.array_equal(planes, exp)
    assert (exp[1] != img).sum() > 1000
    assert np.array_equal(pv.debug_planes(d_in[0]), np.stack([img] * 3))  # no overlay
    jpg = pv.encode(d_in[0], d_kp, 28, ex.max_kp, d_cnt, stream=st)
    assert jpg[:2] == b"\xff\xd8" and jpg[-2:] == b"\xff\xd9" and 10_000 < len(jpg) < w * h
    dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)  # BGR
    assert dec.shape == (h, w, 3)
    err = dec[:, :, 1].astype(np.float32) - exp[1].astype(np.float32)  # G plane; 4:2:0 chroma smears the overlay a little
    assert 10 * np.log10(255.0 ** 2 / np.mean(err ** 2)) > 24.0
    gray_err = dec[:, :, 2].astype(np.float32)[exp[1] == img] - img.astype(np.float32)[exp[1] == img]
    assert np.abs(gray_err).mean() < 12.0
    # a second frame re-uses planes and encoder state; a strided frame (pitch > width) works too
    wide = torch.zeros((h, w + 64), dtype=torch.uint8, device="cuda")
    wide[:, :w] = d_in[0]
    jpg2 = pv.encode(wide, None, gray_pitch=w + 64, stream=st)
    dec2 = cv2.imdecode(np.frombuffer(jpg2, np.uint8), cv2.IMREAD_GRAYSCALE)
    assert np.abs(dec2.astype(np.float32) - img.astype(np.float32)).mean() < 6.0
    msg = orbb.slam_frame_to_bson(1, 2, 3, w, h, kp[:10, 0].astype(np.uint16), kp[:10, 1].astype(np.uint16), image=jpg, channels=3)
    assert jpg in msg
    pv.close(); ex.close()

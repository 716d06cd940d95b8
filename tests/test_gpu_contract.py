"""GPU tier, part 2: the paths behind the headline numbers and the boundary contract of include/orbb200.h.

  * the bench workloads themselves -- 256 x 640x480 through the two-stream device path and through the double-buffered
    host path, 64 x 848x480 (cfg 2) -- against the oracle on sampled frames of BOTH batch parts;
  * Jetracer::compute_fast_angle / calc_orb as separate calls (reference src/cuda/orb.cuh:9-27) == the fused kernel;
  * stage calls are re-runnable; no entry point other than *_host / wait / debug_* blocks the calling thread;
  * matcher scratch is fixed-size: query sets larger than it are chunked, results unchanged.
"""
import importlib
import os
import time

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


def canon(kp, desc):
    order = np.lexsort((kp["x"], kp["y"], kp["octave"]))
    return kp[order], desc[order]


def _check_frames(oracle, frames, sample, kp, desc, cnt, w, h, nf):
    o = oracle.Oracle(w, h, nf)
    for f in sample:
        okp, odesc = canon(*o.extract(frames[f]))
        gkp, gdesc = canon(kp[f, :cnt[f]], desc[f, :cnt[f]])
        assert len(gkp) == len(okp), f"frame {f}: {len(gkp)} vs {len(okp)} keypoints"
        assert gkp.tobytes() == okp.tobytes(), f"frame {f}: keypoint fields differ"
        assert np.array_equal(gdesc, odesc), f"frame {f}: descriptor bits differ"


def test_headline_batch_256x640x480_vs_oracle(orbb, oracle, synth):
    """bench.py's timed call: 256 frames of 640x480 / 1000 kp in ONE orbb_extract_batch_device (two batch parts on two
    streams, throughput kernel variants) and in ONE orbb_extract_batch_host_async (chunked H2D -> compute -> D2H).
    Sampled frames of both parts, first / last frame of each part included, must equal the oracle bit for bit."""
    import torch
    w, h, nf, nb = 640, 480, 1000, 256
    frames = synth.rolled_batch(w, h, nb, 1000)
    ex = orbb.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=nb)
    st = torch.cuda.current_stream()
    d_in = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((nb, ex.max_kp, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((nb, ex.max_kp, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(nb, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_in, nb, d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(nb, ex.max_kp)
    desc, cnt = d_desc.cpu().numpy(), d_cnt.cpu().numpy()
    sample = [0, 1, 63, 127, 128, 129, 200, 255]
    _check_frames(oracle, frames, sample, kp, desc, cnt, w, h, nf)
    # the end-to-end path of the headline e2e number: same bytes as the device path for EVERY frame
    pin = torch.from_numpy(frames).pin_memory()
    hk = torch.zeros(nb * ex.max_kp * 28, dtype=torch.uint8).pin_memory()
    hd = torch.zeros(nb * ex.max_kp * 32, dtype=torch.uint8).pin_memory()
    hc = torch.zeros(nb, dtype=torch.int32).pin_memory()
    for _ in range(2):  # second submission: same chunk layout, streams not cross-serialised
        t = ex.extract_batch_host_async(pin.data_ptr(), w, w * h, nb, hk.data_ptr(), hd.data_ptr(), hc.data_ptr(), stream=st)
        ex.wait(t)
        assert np.array_equal(hc.numpy(), cnt)
        k2 = np.frombuffer(hk.numpy().tobytes(), orbb.KEYPOINT_DTYPE).reshape(nb, ex.max_kp)
        d2 = hd.numpy().reshape(nb, ex.max_kp, 32)
        for f in range(nb):
            n = int(cnt[f])
            assert k2[f, :n].tobytes() == kp[f, :n].tobytes() and d2[f, :n].tobytes() == desc[f, :n].tobytes(), f
        hk.zero_(); hd.zero_(); hc.zero_()
    ex.close()


def test_cfg2_batch_64x848x480_vs_oracle(orbb, oracle, synth):
    """BASELINE cfg 2: 64 frames of 848x480, 1200 kp, one batch (the multi-stream split starts at 64 frames)."""
    w, h, nf, nb = 848, 480, 1200, 64
    frames = synth.rolled_batch(w, h, nb, 2000, n_base=8)
    ex = orbb.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=nb)
    kp, desc, cnt = ex.extract_batch(frames)
    _check_frames(oracle, frames, [0, 31, 32, 47, 63], kp, desc, cnt, w, h, nf)
    ex.close()


def test_separate_angle_and_orb_entry_points(orbb, oracle, synth):
    """The reference calls compute_fast_angle and calc_orb one after the other on caller-owned SoA arrays
    (buildStream.cpp:442-460).  orbb_detect_export + orbb_compute_fast_angle + orbb_calc_orb per level must give the
    angles and descriptor bits of the fused kernel (and so of the oracle)."""
    import torch
    w, h, nf = 640, 480, 1000
    frames = np.stack([synth.textured_frame(w, h, 8100), synth.low_contrast_frame(w, h, 8101)])
    ex = orbb.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=2)
    st = torch.cuda.current_stream()
    d_in = torch.from_numpy(frames).cuda()
    mk = ex.max_kp
    ex.stage_upload(d_in, 2, stream=st)
    ex.pyramid_create_levels(stream=st)
    ex.detect(stream=st)
    ex.gaussian_blur(stream=st)
    d_kp = torch.zeros((2, mk, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((2, mk, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    ex.compute_fast_angle_and_orb(d_kp, d_desc, d_cnt, stream=st)
    # the two-call form
    d_pos = torch.zeros((2, mk, 2), dtype=torch.float32, device="cuda")
    d_score = torch.zeros((2, mk), dtype=torch.float32, device="cuda")
    d_level = torch.full((2, mk), -1, dtype=torch.int32, device="cuda")
    d_lc = torch.zeros((2, 8), dtype=torch.int32, device="cuda")
    d_c2 = torch.zeros(2, dtype=torch.int32, device="cuda")
    ex.detect_export(d_pos, d_score, d_level, d_lc, d_c2, stream=st)
    torch.cuda.synchronize()
    lc = d_lc.cpu().numpy()
    assert np.array_equal(d_c2.cpu().numpy(), d_cnt.cpu().numpy()) and np.array_equal(lc.sum(1), d_cnt.cpu().numpy())
    d_angle = torch.full((2, mk), -7.0, dtype=torch.float32, device="cuda")
    d_desc2 = torch.zeros((2, mk, 32), dtype=torch.uint8, device="cuda")
    for f in range(2):
        off = 0
        for l in range(8):
            n = int(lc[f, l])
            li = ex.level_info(l, frame=f)
            roi = li.padded + 19 * li.pitch + 19
            ex.compute_fast_angle(d_angle[f, off:], d_pos[f, off:], roi, li.pitch, li.width, li.height, n, stream=st)
            ex.calc_orb(d_angle[f, off:], d_pos[f, off:], d_desc2[f, off:], li.blurred, li.pitch, li.width, li.height, n, stream=st)
            off += n
    torch.cuda.synchronize()
    kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(2, mk)
    cnt = d_cnt.cpu().numpy()
    for f in range(2):
        n = int(cnt[f])
        assert n > 500
        assert np.array_equal(d_angle[f, :n].cpu().numpy(), kp[f, :n]["angle"])
        assert np.array_equal(d_desc2[f, :n].cpu().numpy(), d_desc[f, :n].cpu().numpy())
        assert np.array_equal(d_level[f, :n].cpu().numpy(), kp[f, :n]["octave"])
        assert np.array_equal(d_score[f, :n].cpu().numpy(), kp[f, :n]["response"])
        sc = ex.GetScaleFactors()[kp[f, :n]["octave"]]
        pos = d_pos[f, :n].cpu().numpy()
        lvl0 = kp[f, :n]["octave"] == 0
        assert np.array_equal(np.where(lvl0, pos[:, 0], pos[:, 0] * sc), kp[f, :n]["x"])
        okp, odesc = canon(*oracle.Oracle(w, h, nf).extract(frames[f]))
        gk, gd = canon(kp[f, :n], d_desc2[f, :n].cpu().numpy())
        assert gk.tobytes() == okp.tobytes() and np.array_equal(gd, odesc)
    # keypoints too close to the edge: angle -1 / zero descriptor, never an out-of-bounds read
    li = ex.level_info(0)
    edge = torch.tensor([[3.0, 3.0], [w - 2.0, 100.0], [100.0, h - 1.0], [320.0, 240.0]], dtype=torch.float32, device="cuda")
    a = torch.zeros(4, dtype=torch.float32, device="cuda")
    dd = torch.full((4, 32), 9, dtype=torch.uint8, device="cuda")
    ex.compute_fast_angle(a, edge, li.padded + 19 * li.pitch + 19, li.pitch, w, h, 4, stream=st)
    ex.calc_orb(a, edge, dd, li.blurred, li.pitch, w, h, 4, stream=st)
    torch.cuda.synchronize()
    assert a[:3].cpu().tolist() == [-1.0, -1.0, -1.0] and 0.0 <= float(a[3]) < 360.0
    assert int(dd[:3].sum()) == 0 and int(dd[3].sum()) > 0
    ex.close()


def test_stage_calls_are_rerunnable(orbb, synth):
    """orbb_detect twice on the resident batch: same selection, no duplicated candidates; also
    detect_fast twice WITHOUT a distribute in between (the cell tables are then stale and the quadtree kernel must
    notice and take its general path)."""
    import torch
    w, h = 424, 240
    frames = np.stack([synth.textured_frame(w, h, 61), synth.sparse_frame(w, h, 62)])
    ex = orbb.ORBextractor(600, 1.2, 6, 20, 7, width=w, height=h, max_batch=2)
    rk, rd, rc = ex.extract_batch(frames)
    st = torch.cuda.current_stream()
    d_in = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((2, ex.max_kp, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((2, ex.max_kp, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    ex.stage_upload(d_in, 2, stream=st)
    ex.pyramid_create_levels(stream=st)
    ex.detect(stream=st)
    ex.detect(stream=st)             # second detect on the same batch
    ncand = [len(ex.debug_candidates(l, f)) for f in range(2) for l in range(6)]
    ex.detect_fast(stream=st)
    ex.detect_fast(stream=st)        # stale cell tables now
    assert ncand == [len(ex.debug_candidates(l, f)) for f in range(2) for l in range(6)]
    ex.detect_distribute(stream=st)
    ex.gaussian_blur(stream=st)
    ex.gaussian_blur(stream=st)
    ex.compute_fast_angle_and_orb(d_kp, d_desc, d_cnt, stream=st)
    ex.compute_fast_angle_and_orb(d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy()
    assert np.array_equal(cnt, rc)
    kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(2, ex.max_kp)
    for f in range(2):
        n = int(cnt[f])
        assert kp[f, :n].tobytes() == rk[f, :n].tobytes() and np.array_equal(d_desc[f, :n].cpu().numpy(), rd[f, :n])
    # and the whole-extractor call afterwards starts from clean tables
    k3, d3, c3 = ex.extract_batch(frames)
    assert np.array_equal(c3, rc) and k3.tobytes() == rk.tobytes() and d3.tobytes() == rd.tobytes()
    ex.close()


def test_no_device_entry_point_blocks(orbb, synth):
    """orbb200.h: "all work is enqueued on the stream given, no hidden synchronisation".  A long spin kernel is put
    on the stream first; every device-side entry point must return while it is still running (stream.query() False
    right after the call).  Only *_host, orbb_wait and debug_* may block."""
    import torch
    w, h, nb = 320, 240, 4
    frames = np.stack([synth.textured_frame(w, h, 70 + i) for i in range(nb)])
    ex = orbb.ORBextractor(400, 1.2, 4, 20, 7, width=w, height=h, max_batch=nb)
    mk = ex.max_kp
    st = torch.cuda.Stream()
    d_in = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((nb, mk, 7), dtype=torch.float32, device="cuda")
    d_kp2 = torch.zeros((nb, mk, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((nb, mk, 32), dtype=torch.uint8, device="cuda")
    d_desc2 = torch.zeros((nb, mk, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(nb, dtype=torch.int32, device="cuda")
    d_valid = torch.zeros(nb, dtype=torch.int32, device="cuda")
    d_idx = torch.zeros((nb * mk, 2), dtype=torch.int32, device="cuda")
    d_dist = torch.zeros((nb * mk, 2), dtype=torch.int32, device="cuda")
    d_acc = torch.zeros(nb * mk, dtype=torch.uint8, device="cuda")
    d_nacc = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_pos = torch.zeros((nb, mk, 2), dtype=torch.float32, device="cuda")
    d_angle = torch.zeros((nb, mk), dtype=torch.float32, device="cuda")
    d_off = torch.tensor([0, mk, 2 * mk], dtype=torch.int32, device="cuda")
    d_rgb = torch.zeros((nb, h, w, 3), dtype=torch.uint8, device="cuda")
    d_gray = torch.zeros((nb, h, w), dtype=torch.uint8, device="cuda")
    d_depth = torch.full((nb, h, w), 1500, dtype=torch.int16, device="cuda")
    d_al = torch.zeros((nb, h, w), dtype=torch.int32, device="cuda")
    d_pts = torch.zeros((nb, mk, 3), dtype=torch.float64, device="cuda")
    d_ur = torch.zeros((nb // 2, mk), dtype=torch.float32, device="cuda")
    d_dp = torch.zeros((nb // 2, mk), dtype=torch.float32, device="cuda")
    d_ns = torch.zeros(nb // 2, dtype=torch.int32, device="cuda")
    intr = orbb.make_intrinsics(w, h, w / 2, h / 2, 300.0, 300.0)
    extr = orbb.make_extrinsics()
    li = ex.level_info(0)
    calls = [
        ("extract_batch_device", lambda: ex.extract_batch_device(d_in, nb, d_kp, d_desc, d_cnt, stream=st)),
        ("stage_upload", lambda: ex.stage_upload(d_in, nb, stream=st)),
        ("pyramid_create_levels", lambda: ex.pyramid_create_levels(stream=st)),
        ("detect", lambda: ex.detect(stream=st)),
        ("detect_fast", lambda: ex.detect_fast(stream=st)),
        ("detect_distribute", lambda: ex.detect_distribute(stream=st)),
        ("gaussian_blur", lambda: ex.gaussian_blur(stream=st)),
        ("compute_angle_and_orb", lambda: ex.compute_fast_angle_and_orb(d_kp, d_desc, d_cnt, stream=st)),
        ("detect_export", lambda: ex.detect_export(d_pos, None, None, None, None, stream=st)),
        ("compute_fast_angle", lambda: ex.compute_fast_angle(d_angle, d_pos, li.padded + 19 * li.pitch + 19, li.pitch, w, h, 64, stream=st)),
        ("calc_orb", lambda: ex.calc_orb(d_angle, d_pos, d_desc2, li.blurred, li.pitch, w, h, 64, stream=st)),
        ("match_knn", lambda: ex.match_keypoints(d_desc, nb * mk, d_desc, nb * mk, d_idx, d_dist, d_acc, d_nacc, k=2, stream=st)),
        ("match_knn_segmented", lambda: ex.match_keypoints_segmented(d_desc, d_off, d_desc, d_off, 2, 2 * mk, mk, mk, d_idx, d_dist, d_acc, stream=st)),
        ("match_windowed", lambda: ex.match_keypoints_windowed(d_desc, d_kp, 28, mk, d_desc, d_kp, 28, mk, 2.0, 64, d_idx, d_dist, d_nacc, stream=st)),
        ("rgb_to_grayscale", lambda: ex.rgb_to_grayscale(d_rgb, nb, d_gray, stream=st)),
        ("align_depth_to_other", lambda: ex.align_depth_to_other(d_depth, nb, 0.001, intr, intr, extr, d_al, stream=st)),
        ("keypoint_pixel_to_point", lambda: ex.keypoint_pixel_to_point(d_al, intr, nb, d_kp, d_desc, d_cnt, d_kp2, d_desc2, d_pts, d_valid, stream=st)),
        ("reproject_points", lambda: ex.reproject_points(d_pts, d_valid, nb, None, intr, d_pos, stream=st)),
        ("match_windowed_batch", lambda: ex.match_keypoints_windowed_batch(d_desc2, d_pos, d_valid, d_desc2, d_kp2, 28, d_valid, nb, 2.0, 64, d_idx, d_dist, stream=st)),
        ("match_projection_batch", lambda: ex.match_keypoints_projection_batch(d_desc2, d_pos, d_kp2, d_valid, d_desc2, d_kp2, d_valid, nb, 7.0, 100, True, d_idx, d_dist, d_ns, stream=st)),
        ("compute_stereo_matches", lambda: ex.compute_stereo_matches(d_kp, d_desc, d_cnt, nb // 2, 40.0, 300.0, d_ur, d_dp, d_ns, stream=st)),
    ]
    for _, c in calls:   # warm up: module load, first-use attribute calls
        c()
    torch.cuda.synchronize()
    spin = int(0.25 * 1.9e9)  # ~0.25 s at 1.9 GHz
    blocked = []
    for name, c in calls:
        with torch.cuda.stream(st):
            torch.cuda._sleep(spin)
        t0 = time.perf_counter()
        c()
        dt = time.perf_counter() - t0
        still_running = not st.query()
        st.synchronize()
        if not still_running or dt > 0.1:
            blocked.append((name, round(dt, 4)))
    assert not blocked, f"entry points that waited for the stream: {blocked}"
    # the async host form may not block either (its ticket is waited for separately)
    pin = torch.from_numpy(frames).pin_memory()
    hk = torch.zeros(nb * mk * 28, dtype=torch.uint8).pin_memory()
    hd = torch.zeros(nb * mk * 32, dtype=torch.uint8).pin_memory()
    hc = torch.zeros(nb, dtype=torch.int32).pin_memory()
    with torch.cuda.stream(st):
        torch.cuda._sleep(spin)
    t0 = time.perf_counter()
    t = ex.extract_batch_host_async(pin.data_ptr(), w, w * h, nb, hk.data_ptr(), hd.data_ptr(), hc.data_ptr(), stream=st)
    assert time.perf_counter() - t0 < 0.1 and not st.query()
    ex.wait(t)
    assert int(hc.sum()) > 0
    ex.close()


def test_matcher_scratch_is_fixed_and_chunks(orbb, oracle):
    """The split-T scratch is allocated in orbb_create; a query set larger than it holds goes through in chunks and
    the per-chunk accept counts accumulate.  300 k queries against 700 train rows, checked against the oracle."""
    import torch
    rng = np.random.default_rng(17)
    ex = orbb.ORBextractor(300, 1.2, 2, 20, 7, width=200, height=150, max_batch=1)
    nq, nt = 300_000, 700
    t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    q = t[rng.integers(0, nt, size=nq)].copy()
    q[:, :4] ^= rng.integers(0, 256, size=(nq, 4), dtype=np.uint8) & rng.integers(0, 256, size=(nq, 4), dtype=np.uint8)
    free0 = torch.cuda.mem_get_info()[0]
    for k in (1, 2):
        idx, dist, acc, nacc = orbb.match_knn_host(ex, q, t, k=k, ratio=0.7)
        oidx, odist, oacc = oracle.match_knn(q, t, k=k, ratio=0.7, threads=os.cpu_count() or 4)
        if k == 1:
            oidx[:, 1] = -1; odist[:, 1] = -1
        assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
        assert np.array_equal(acc, oacc) and nacc == int(oacc.sum())
    dq = torch.from_numpy(q).cuda(); dt = torch.from_numpy(t).cuda()
    idx = torch.zeros((nq, 2), dtype=torch.int32, device="cuda"); dist = torch.zeros_like(idx)
    free1 = torch.cuda.mem_get_info()[0]
    ex.match_keypoints(dq, nq, dt, nt, idx, dist, k=1)
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] == free1  # nothing was allocated by the call
    del free0
    ex.close()


def test_create_and_stride_validation(orbb):
    with pytest.raises(orbb.OrbbError):
        orbb.ORBextractor(100, 1.2, 2, 20, 7, width=200, height=150, max_batch=70000)  # frame index is a grid dimension
    import torch
    ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=200, height=150, max_batch=2)
    d_in = torch.zeros((2, 150, 200), dtype=torch.uint8, device="cuda")
    d_kp = torch.zeros((2, ex.max_kp, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((2, ex.max_kp, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    with pytest.raises(orbb.OrbbError):
        ex.extract_batch_device(d_in, 2, d_kp, d_desc, d_cnt, stride=200 * 100)  # frames would overlap
    with pytest.raises(orbb.OrbbError):
        ex.stage_upload(d_in, 2, stride=100)
    ex.extract_batch_device(d_in, 1, d_kp, d_desc, d_cnt, stride=0)  # a single frame ignores the stride
    torch.cuda.synchronize()
    ex.close()


def test_match_knn_batch_fixed_stride(orbb, oracle, synth):
    """orbb_match_knn_batch (cfg 5: every frame of an extraction output against one map): rows below each frame's
    device-side count equal the oracle's 1-NN / 2-NN, rows past it report -1, nothing is compacted."""
    import torch
    w, h, nb = 320, 240, 5
    frames = np.stack([synth.textured_frame(w, h, 900 + i) for i in range(3)] + [synth.sparse_frame(w, h, 5), synth.flat_frame(w, h)])
    ex = orbb.ORBextractor(500, 1.2, 8, 20, 7, width=w, height=h, max_batch=nb)
    mk = ex.max_kp
    st = torch.cuda.current_stream()
    d_in = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((nb, mk, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.full((nb, mk, 32), 0xAB, dtype=torch.uint8, device="cuda")  # rows past the counts hold junk
    d_cnt = torch.zeros(nb, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_in, nb, d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy()
    assert cnt[4] == 0 and 0 < cnt[3] < 300 and cnt[0] > 400
    desc = d_desc.cpu().numpy()
    rng = np.random.default_rng(2)
    tmap = np.concatenate([desc[0, :cnt[0]], desc[1, :cnt[1]], rng.integers(0, 256, size=(3000, 32), dtype=np.uint8)])
    tmap[7] = tmap[2]  # duplicate rows: ties -> lowest train index
    d_map = torch.from_numpy(tmap).cuda()
    for k in (1, 2):
        d_idx = torch.full((nb * mk, 2), -9, dtype=torch.int32, device="cuda")
        d_dist = torch.full((nb * mk, 2), -9, dtype=torch.int32, device="cuda")
        d_acc = torch.full((nb * mk,), 9, dtype=torch.uint8, device="cuda")
        d_nacc = torch.zeros(1, dtype=torch.int32, device="cuda")
        ex.match_keypoints_batch(d_desc, d_cnt, nb, d_map, len(tmap), d_idx, d_dist, d_acc, d_nacc, k=k, ratio=0.7, stream=st)
        torch.cuda.synchronize()
        idx = d_idx.cpu().numpy().reshape(nb, mk, 2); dist = d_dist.cpu().numpy().reshape(nb, mk, 2)
        acc = d_acc.cpu().numpy().reshape(nb, mk)
        nacc = 0
        for f in range(nb):
            c = int(cnt[f])
            assert (idx[f, c:] == -1).all() and (dist[f, c:] == -1).all() and (acc[f, c:] == 0).all()
            if c == 0:
                continue
            oi, od, oa = oracle.match_knn(desc[f, :c], tmap, k=k, ratio=0.7)
            if k == 1:
                oi[:, 1] = -1; od[:, 1] = -1
            assert np.array_equal(idx[f, :c], oi) and np.array_equal(dist[f, :c], od), (f, k)
            assert np.array_equal(acc[f, :c].astype(bool), oa)
            nacc += int(oa.sum())
        assert int(d_nacc.item()) == nacc
    ex.close()


def test_popc_rate_microbenchmark(orbb):
    """The matcher's roofline denominator is measured, not quoted: POPC lanes per clock per SM."""
    ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=200, height=150, max_batch=1)
    r = ex.debug_popc_rate()
    assert 8.0 < r < 40.0, r
    ex.close()


def test_results_do_not_depend_on_scratch_contents_or_touch_guards(orbb, oracle, synth):
    """Stand-in for compute-sanitizer (closed on this GPU pool, see profiles/r02_sanitizer_unavailable.txt):
      * initcheck: every stateless scratch buffer of the handle is filled with 0xCD / 0x00 / 0xFF before an extraction --
        pyramid pad bytes, candidate lists, sort scratch, staging -- and the result must stay bit-identical;
      * memcheck (writes): all output arrays sit between 64 KB guard bands that must come back untouched, for the
        extractor, the matcher (all three forms) and the two-call angle / descriptor entry points;
      * racecheck: the same batch is extracted 12 times, alternating batch sizes so different kernel variants and the
        one-launch pyramid run, and every run must give the same bytes as the oracle-checked first one."""
    import torch
    w, h, nb = 424, 240, 6
    frames = np.stack([synth.textured_frame(w, h, 7300 + i) for i in range(nb - 2)] + [synth.sparse_frame(w, h, 7), synth.low_contrast_frame(w, h, 8)])
    ex = orbb.ORBextractor(600, 1.2, 6, 20, 7, width=w, height=h, max_batch=nb)
    mk, G = ex.max_kp, 65536
    st = torch.cuda.current_stream()

    def guarded(nbytes, fill):
        buf = torch.full((nbytes + 2 * G,), fill, dtype=torch.uint8, device="cuda")
        return buf, buf[G:G + nbytes]

    def guards_ok(buf, nbytes, fill):
        return bool((buf[:G] == fill).all()) and bool((buf[G + nbytes:] == fill).all())

    d_in = torch.from_numpy(frames).cuda()
    bk, d_kp = guarded(nb * mk * 28, 0x5A)
    bd, d_desc = guarded(nb * mk * 32, 0x5A)
    bc, d_cnt8 = guarded(nb * 4, 0x5A)
    d_cnt = d_cnt8.view(torch.int32)
    ref = None
    for rep, poison in enumerate([None, 0xCD, 0x00, 0xFF, 0xCD, None, 0x37, 0xFF, 0x00, 0xCD, None, 0x11]):
        if poison is not None:
            ex.debug_poison(poison)
        n = nb if rep % 3 == 0 else (1 if rep % 3 == 1 else 3)  # 1 and 3 frames: the small-grid kernel variants
        d_kp.fill_(0x5A); d_desc.fill_(0x5A); d_cnt8.fill_(0x5A)
        ex.extract_batch_device(d_in, n, d_kp, d_desc, d_cnt, stream=st)
        torch.cuda.synchronize()
        assert guards_ok(bk, nb * mk * 28, 0x5A) and guards_ok(bd, nb * mk * 32, 0x5A) and guards_ok(bc, nb * 4, 0x5A)
        cnt = d_cnt.cpu().numpy()[:n]
        kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(nb, mk)
        desc = d_desc.cpu().numpy().reshape(nb, mk, 32)
        if ref is None:
            assert n == nb
            o = oracle.Oracle(w, h, 600, 1.2, 6)
            for f in range(nb):
                okp, od = canon(*o.extract(frames[f]))
                gk, gd = canon(kp[f, :cnt[f]], desc[f, :cnt[f]])
                assert gk.tobytes() == okp.tobytes() and np.array_equal(gd, od)
            ref = (cnt.copy(), kp.copy(), desc.copy())
        for f in range(n):
            c = int(ref[0][f])
            assert cnt[f] == c and kp[f, :c].tobytes() == ref[1][f, :c].tobytes() and desc[f, :c].tobytes() == ref[2][f, :c].tobytes(), (rep, f)
            # rows past the count are never written
            assert (d_desc.view(nb, mk, 32)[f, c:] == 0x5A).all()
    # matcher outputs between guards (plain, batch, segmented, windowed)
    nq = int(ref[0][0])
    q = torch.from_numpy(ref[2][0, :nq].copy()).cuda()
    t = torch.from_numpy(ref[2][1, :int(ref[0][1])].copy()).cuda()
    bi, d_idx8 = guarded(nb * mk * 8, 0xA5); bdst, d_dist8 = guarded(nb * mk * 8, 0xA5); ba, d_acc = guarded(nb * mk, 0xA5)
    d_idx, d_dist = d_idx8.view(torch.int32).view(-1, 2), d_dist8.view(torch.int32).view(-1, 2)
    ex.debug_poison(0xEE)
    ex.match_keypoints(q, nq, t, t.shape[0], d_idx, d_dist, d_acc, None, k=2, stream=st)
    ex.extract_batch_device(d_in, nb, d_kp, d_desc, d_cnt, stream=st)
    ex.match_keypoints_batch(d_desc.view(nb, mk, 32), d_cnt, nb, t, t.shape[0], d_idx, d_dist, d_acc, None, k=1, stream=st)
    torch.cuda.synchronize()
    assert guards_ok(bi, nb * mk * 8, 0xA5) and guards_ok(bdst, nb * mk * 8, 0xA5) and guards_ok(ba, nb * mk, 0xA5)
    oi, od, _ = oracle.match_knn(ref[2][0, :nq], ref[2][1, :int(ref[0][1])], k=1)
    got = d_idx.cpu().numpy().reshape(nb, mk, 2)[0, :nq, 0]
    assert np.array_equal(got, oi[:, 0])
    ex.close()


def test_fused_pyramid_opt_in(orbb, oracle, synth, monkeypatch):
    """ORBB_FUSED_PYRAMID=n: batches of up to n frames build level 0 and every further level in ceil(nlevels / 4)
    launches (k_pyramid_fused: tile pyramids in shared memory, halos recomputed).  Opt-in because it measured slower
    than the per-level kernels; it must stay bit-exact -- every padded pixel of every level, odd sizes, scale 2.0
    (INTER_AREA route) and more than four levels (two groups)."""
    monkeypatch.setenv("ORBB_FUSED_PYRAMID", "4")
    for (w, h, nf, sc, nl, seed) in ((640, 480, 1000, 1.2, 8, 1000), (333, 251, 400, 1.2, 6, 5), (512, 384, 600, 2.0, 3, 21),
                                     (97, 83, 120, 1.2, 2, 33), (848, 480, 405, 1.2, 1, 2100), (400, 300, 400, 1.5, 4, 9)):
        frames = np.stack([synth.textured_frame(w, h, seed), synth.textured_frame(w, h, seed + 1), synth.sparse_frame(w, h, 3)])
        ex = orbb.ORBextractor(nf, sc, nl, 20, 7, width=w, height=h, max_batch=3)
        l0 = ex.launch_count()
        kp, desc, cnt = ex.extract_batch(frames)
        assert ex.launch_count() - l0 == (nl + 3) // 4 + 4  # fused groups + FAST + quadtree + blur + angle/rBRIEF
        o = oracle.Oracle(w, h, nf, sc, nl, 20, 7)
        for f in range(3):
            okp, od = canon(*o.extract(frames[f]))
            for l in range(nl):
                assert np.array_equal(ex.debug_padded(l, frame=f), o.level_padded(l)), (w, h, l, f)
            gk, gd = canon(kp[f, :cnt[f]], desc[f, :cnt[f]])
            assert gk.tobytes() == okp.tobytes() and np.array_equal(gd, od)
        ex.close()


def test_small_batch_graph_cache_and_fallback(orbb, oracle, synth, monkeypatch):
    """Small batches replay a captured graph keyed by the argument set (4 slots).  A caller that hands in fresh buffers on
    every call misses every time: after eight misses in a row the handle must fall back to plain stream launches, and
    every call on the way -- captured, patched in place (cudaGraphExecUpdate) or direct -- must give the oracle's
    bytes.  ORBB_GRAPH=0 (never capture) is checked on the same frames."""
    import torch
    w, h = 424, 240
    frames = np.stack([synth.textured_frame(w, h, 8800 + i) for i in range(2)])
    o = oracle.Oracle(w, h, 500, 1.2, 6)
    ref = [canon(*o.extract(f)) for f in frames]
    st = torch.cuda.current_stream()

    def check(ex, d_kp, d_desc, d_cnt, n):
        cnt = d_cnt.cpu().numpy()
        kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(-1, ex.max_kp)
        desc = d_desc.cpu().numpy()
        for f in range(n):
            gk, gd = canon(kp[f, :cnt[f]], desc[f, :cnt[f]])
            assert gk.tobytes() == ref[f][0].tobytes() and np.array_equal(gd, ref[f][1])

    for graph in ("1", "0"):
        monkeypatch.setenv("ORBB_GRAPH", graph)
        ex = orbb.ORBextractor(500, 1.2, 6, 20, 7, width=w, height=h, max_batch=2)
        per_call, keep = [], []  # `keep` stops the caching allocator from handing the same addresses out again
        for i in range(14):  # a new set of device buffers every call
            n = 1 + (i % 2)
            d_in = torch.from_numpy(frames[:n].copy()).cuda()
            d_kp = torch.zeros((n, ex.max_kp, 7), dtype=torch.float32, device="cuda")
            d_desc = torch.zeros((n, ex.max_kp, 32), dtype=torch.uint8, device="cuda")
            d_cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
            l0 = ex.launch_count()
            ex.extract_batch_device(d_in, n, d_kp, d_desc, d_cnt, stream=st)
            torch.cuda.synchronize()
            per_call.append(ex.launch_count() - l0)
            check(ex, d_kp, d_desc, d_cnt, n)
            keep.append((d_in, d_kp, d_desc, d_cnt))
        if graph == "1":
            assert per_call[0] == 14 and per_call[-1] == 10  # captured (detection in three level groups: 4 more launches), then direct
        else:
            assert set(per_call) == {10}
        ex.close()


@pytest.mark.parametrize("groups", ["1,4", "4", "1,2,5", "3", "2,9", "x"])
def test_lone_frame_level_groups(orbb, oracle, synth, monkeypatch, groups):
    """Inside the small-batch graph the detection runs in level groups on side streams (default {0}{1-3}{rest}).  Any
    grouping -- also boundaries past the level count and a malformed spec, which fall back to usable ones -- must give
    the oracle's bytes: the groups only change which stream a (frame, level)'s FAST + quadtree kernels run on."""
    import torch
    monkeypatch.setenv("ORBB_LONE_GROUPS", groups)
    st = torch.cuda.current_stream()
    for (w, h, nf, nl, n) in ((424, 240, 500, 8, 1), (640, 480, 1000, 5, 2), (333, 251, 300, 4, 3)):
        frames = np.stack([synth.textured_frame(w, h, 6100 + i) for i in range(n)])
        ex = orbb.ORBextractor(nf, 1.2, nl, 20, 7, width=w, height=h, max_batch=n)
        o = oracle.Oracle(w, h, nf, 1.2, nl, 20, 7)
        d_in = torch.from_numpy(frames).cuda()
        d_kp = torch.zeros((n, ex.max_kp, 7), dtype=torch.float32, device="cuda")
        d_desc = torch.zeros((n, ex.max_kp, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
        for rep in range(2):  # captured, then replayed
            d_kp.zero_(); d_desc.zero_(); d_cnt.zero_()
            ex.extract_batch_device(d_in, n, d_kp, d_desc, d_cnt, stream=st)
            torch.cuda.synchronize()
            cnt = d_cnt.cpu().numpy()
            kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(n, ex.max_kp)
            desc = d_desc.cpu().numpy()
            for f in range(n):
                okp, od = canon(*o.extract(frames[f]))
                gk, gd = canon(kp[f, :cnt[f]], desc[f, :cnt[f]])
                assert gk.tobytes() == okp.tobytes() and np.array_equal(gd, od), (groups, w, h, f, rep)
        ex.close()

"""GPU parity tests: the CUDA path (through the C ABI, via ctypes) against the CPU oracle and the committed
golden vectors.  Bit-exact bar for pyramid pixels, FAST scores, candidate sets, selected keypoints
(x, y, octave, response, size), angles and descriptor bits.  Run on the B200 box: pytest -m gpu."""
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orbb():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import __graft_entry__ as g
    g.build()
    return importlib.import_module("jetracer-orbslam2_b200.orbb")


def canon(kp, desc=None):
    order = np.lexsort((kp["x"], kp["y"], kp["octave"]))
    return (kp[order], desc[order]) if desc is not None else kp[order]


def as_set(xyr):
    return {tuple(int(v) for v in r) for r in np.asarray(xyr).reshape(-1, 3)}


CONFIGS = [  # (w, h, nfeatures, scale, nlevels, generator, seed)
    (640, 480, 1000, 1.2, 8, "textured_frame", 1000),
    (848, 480, 1200, 1.2, 8, "textured_frame", 2000),
    (333, 251, 400, 1.2, 6, "textured_frame", 5),
    (320, 240, 500, 1.2, 8, "low_contrast_frame", 11),
    (320, 240, 500, 1.2, 8, "sparse_frame", 5),
    (400, 300, 400, 1.5, 4, "textured_frame", 9),
    (512, 384, 600, 2.0, 3, "textured_frame", 21),
    (848, 800, 1000, 1.2, 8, "textured_frame", 3000),   # cfg 3 geometry (T265-sized)
    (1280, 720, 2000, 1.2, 8, "textured_frame", 4000),  # cfg 4 geometry
    (1920, 1080, 3000, 1.2, 8, "textured_frame", 4100),  # full HD: widest tables, two quadtree roots
    (97, 83, 120, 1.2, 2, "textured_frame", 33),        # smallest useful frame: one or two cells per level
    (752, 480, 1000, 1.3, 6, "textured_frame", 77),     # EuRoC-sized, non-default scale (cells up to 37 px wide)
    (848, 480, 405, 1.2, 1, "textured_frame", 2100),    # the reference's live shape: 1 level (defines.h:2), 405 cells
    (640, 480, 1000, 1.2, 1, "textured_frame", 2101),   # single level carrying the whole quota
    # quotas the 4096-cell table cap cannot serve with 4 cells per keypoint: the quadtree kernel's pyramid path decides by
    # itself whether the selection stays inside the table (table cells become final nodes here) or takes the general path
    (848, 480, 1200, 1.2, 1, "textured_frame", 2102),
    (848, 480, 1200, 1.2, 2, "textured_frame", 2103),
    (1280, 720, 2000, 1.2, 2, "textured_frame", 2104),
]


def make_frame(synth, w, h, gen, seed):
    return getattr(synth, gen)(w, h, seed)


@pytest.mark.parametrize("w,h,nf,sc,nl,gen,seed", CONFIGS)
def test_pyramid_scores_candidates_selection(orbb, oracle, synth, w, h, nf, sc, nl, gen, seed):
    img = make_frame(synth, w, h, gen, seed)
    ex = orbb.ORBextractor(nf, sc, nl, 20, 7, width=w, height=h, max_batch=2)
    o = oracle.Oracle(w, h, nf, sc, nl, 20, 7)
    kp, desc = ex(img)
    okp, odesc = o.extract(img)
    assert np.array_equal(ex.features_per_level, o.nfeat)
    assert np.array_equal(ex.GetScaleFactors(), o.scale)
    for l in range(nl):
        # ComputePyramid: every padded pixel incl. the reflect-101 frame
        assert np.array_equal(ex.debug_padded(l), o.level_padded(l)), f"padded level {l}"
        # 7x7 Gaussian of the ROI
        assert np.array_equal(ex.debug_blurred(l), o.level_blurred(l)), f"blurred level {l}"
        # FAST arc scores: m where m > 7 inside the tested range, else 0
        m = oracle.fast_score_map(o.level_roi(l))
        ref = np.where(m > 7, m, 0).astype(np.uint8)
        ref[:19, :] = 0; ref[-19:, :] = 0; ref[:, :19] = 0; ref[:, -19:] = 0
        assert np.array_equal(ex.debug_scores(l), ref), f"score map level {l}"
        # per-cell FAST(20 -> 7) + NMS candidates (set; GPU order is unspecified)
        oc = o.level_candidates(l)
        got = ex.debug_candidates(l)
        assert len(got) == len(oc)
        assert as_set(got) == as_set(np.stack([oc["x"], oc["y"], oc["response"]], 1)) if len(oc) else len(got) == 0
    # selected keypoints, angles, descriptors
    kp, desc = canon(kp, desc)
    okp, odesc = canon(okp, odesc)
    assert len(kp) == len(okp)
    for f in ("x", "y", "octave", "response", "size", "class_id"):
        assert np.array_equal(kp[f], okp[f]), f
    # north_star tolerance: angles within 1e-3 rad, descriptors <= 8 bits -- we require exact here
    assert np.array_equal(kp["angle"], okp["angle"])
    assert np.array_equal(desc, odesc)
    ex.close()


@pytest.mark.parametrize("name", golden_names())
def test_golden_fixtures(orbb, name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    nf, sc, nl, it, mt = g["params"]
    img = g["image"]
    ex = orbb.ORBextractor(int(nf), float(sc), int(nl), int(it), int(mt), width=img.shape[1], height=img.shape[0])
    kp, desc = canon(*ex(img))
    gkp, gdesc = canon(g["kp"], g["desc"])
    assert len(kp) == len(gkp)
    for f in kp.dtype.names:
        assert np.array_equal(kp[f], gkp[f]), f
    assert np.array_equal(desc, gdesc)
    assert np.array_equal(ex.debug_padded(int(nl) - 1), g["pyr_last_padded"])
    ex.close()


def test_thresholds_other_than_default(orbb, oracle, synth):
    img = synth.textured_frame(320, 240, 31)
    for it, mt in ((35, 12), (12, 12), (10, 25), (60, 3)):
        ex = orbb.ORBextractor(300, 1.2, 5, it, mt, width=320, height=240)
        o = oracle.Oracle(320, 240, 300, 1.2, 5, it, mt)
        kp, desc = canon(*ex(img))
        okp, odesc = canon(*o.extract(img))
        assert len(kp) == len(okp), (it, mt)
        for f in ("x", "y", "octave", "response", "angle"):
            assert np.array_equal(kp[f], okp[f]), (it, mt, f)
        assert np.array_equal(desc, odesc)
        ex.close()


def _canonical_order(cand, W, H):
    """upstream candidate order: cells row-major, inside a cell FAST's row-major order"""
    w_cell = int(np.ceil(np.float32(W) / np.float32(int(np.float32(W) / np.float32(30)))))
    h_cell = int(np.ceil(np.float32(H) / np.float32(int(np.float32(H) / np.float32(30)))))
    i, j = (cand[:, 1] - 3) // h_cell, (cand[:, 0] - 3) // w_cell
    return cand[np.lexsort((cand[:, 0], cand[:, 1], j, i))]


def _octree_case(orbb_ex, oracle, level, cand, quota, rng):
    li = orbb_ex.level_info(level)
    minx, maxx, miny, maxy = 16, li.width - 16, 16, li.height - 16
    cand = _canonical_order(cand, li.width - 32, li.height - 32)
    c = np.zeros(len(cand), oracle.CAND_DTYPE)
    c["x"], c["y"], c["response"] = cand[:, 0], cand[:, 1], cand[:, 2]
    sel = oracle.distribute_octree(c, minx, maxx, miny, maxy, quota)
    ref = as_set(cand[sel])
    got = orbb_ex.debug_distribute(level, cand[rng.permutation(len(cand))], quota)  # GPU order is irrelevant
    assert len(got) == len(sel), (len(got), len(sel), quota, len(cand))
    assert as_set(got) == ref, (quota, len(cand))


@pytest.mark.parametrize("nofast", [False, True])
def test_octree_stress(orbb, oracle, monkeypatch, nofast):
    """DistributeOctTree alone on synthetic candidate clouds: uniform, clustered, collinear, tiny; many quotas
    so every exit (>=N after a full pass, careful phase with 1..k rounds, no growth, n < N) is hit.  Run twice:
    with the cell-table fast path (which falls back to the general path on the clustered / collinear clouds that
    need depths below the table's), and with ORBB_OCT_NOFAST=1 forcing the general radix-sort path everywhere."""
    if nofast:
        monkeypatch.setenv("ORBB_OCT_NOFAST", "1")
    rng = np.random.default_rng(7)
    for (w, h) in ((640, 480), (848, 480), (1280, 720)):
        ex = orbb.ORBextractor(4000, 1.2, 2, 20, 7, width=w, height=h)
        W, H = w - 32, h - 32
        clouds = []
        for n in (1, 2, 3, 17, 150, 1200, 6000):
            pts = set()
            while len(pts) < n:
                pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
            clouds.append(np.array(sorted(pts)))
        # clustered: gaussian blobs
        for n in (300, 3000):
            cx, cy = rng.integers(40, W - 40, 5), rng.integers(40, H - 40, 5)
            pts = set()
            while len(pts) < n:
                k = int(rng.integers(0, 5))
                x, y = int(rng.normal(cx[k], 12)), int(rng.normal(cy[k], 9))
                if 3 <= x < W - 3 and 3 <= y < H - 3:
                    pts.add((x, y))
            clouds.append(np.array(sorted(pts)))
        # a horizontal line, a vertical line, and the root seam columns
        clouds.append(np.array([(x, H // 2) for x in range(3, W - 3, 2)]))
        clouds.append(np.array([(W // 2, y) for y in range(3, H - 3)]))
        clouds.append(np.array([(x, y) for x in range(W // 2 - 3, W // 2 + 4) for y in range(3, H - 3, 5)]))
        for pts in clouds:
            resp = rng.integers(7, 120, size=len(pts))  # many response ties
            cand = np.concatenate([pts, resp[:, None]], 1).astype(np.int32)
            for quota in (1, 5, 60, 217, 434, 1000, 2000):
                _octree_case(ex, oracle, 0, cand, quota, rng)
        ex.close()


def test_batch_equals_single_and_is_deterministic(orbb, synth):
    frames = np.stack([synth.textured_frame(424, 240, 100 + i) for i in range(5)] +
                      [synth.flat_frame(424, 240), synth.checkerboard_frame(424, 240)])
    ex = orbb.ORBextractor(600, 1.2, 6, 20, 7, width=424, height=240, max_batch=8)
    kp, desc, counts = ex.extract_batch(frames)
    kp2, desc2, counts2 = ex.extract_batch(frames)
    assert np.array_equal(counts, counts2)
    assert counts[5] == 0  # flat frame: no corners
    one = orbb.ORBextractor(600, 1.2, 6, 20, 7, width=424, height=240, max_batch=1)
    for f in range(len(frames)):
        n = counts[f]
        assert kp[f, :n].tobytes() == kp2[f, :n].tobytes() and desc[f, :n].tobytes() == desc2[f, :n].tobytes()
        k1, d1 = one(frames[f])
        assert len(k1) == n and k1.tobytes() == kp[f, :n].tobytes() and d1.tobytes() == desc[f, :n].tobytes()
    ex.close(); one.close()


def test_stage_interface_matches_operator(orbb, synth):
    """pyramid_create_levels -> detect -> gaussian_blur -> compute_fast_angle_and_orb == operator()."""
    import torch
    img = synth.textured_frame(640, 480, 77)
    ex = orbb.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=1)
    ref_kp, ref_desc = ex(img)
    d_img = torch.from_numpy(img).cuda()
    d_kp = torch.zeros(ex.max_kp * 28, dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros(ex.max_kp * 32, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    ex.stage_upload(d_img, 1, stream=st)
    ex.pyramid_create_levels(stream=st)
    ex.detect(stream=st)
    ex.gaussian_blur(stream=st)
    ex.compute_fast_angle_and_orb(d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    n = int(d_cnt.item())
    kp = np.frombuffer(d_kp.cpu().numpy().tobytes(), orbb.KEYPOINT_DTYPE)[:n]
    desc = d_desc.cpu().numpy().reshape(-1, 32)[:n]
    assert n == len(ref_kp) and kp.tobytes() == ref_kp.tobytes() and np.array_equal(desc, ref_desc)
    # The quadtree stage consumes (and clears) the cell table the FAST stage filled.  Running it again on the same
    # candidate lists finds an empty table, notices that it does not describe the lists and takes the general
    # sorted-key path: same selection, same bytes.
    d_kp.zero_(); d_desc.zero_()
    ex.detect_distribute(stream=st)
    ex.compute_fast_angle_and_orb(d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    assert int(d_cnt.item()) == n
    assert d_kp.cpu().numpy().tobytes()[:28 * n] == ref_kp.tobytes()
    assert np.array_equal(d_desc.cpu().numpy().reshape(-1, 32)[:n], ref_desc)
    # ... and a fresh run of the whole chain afterwards uses the table again
    kp3, desc3 = ex(img)
    assert kp3.tobytes() == ref_kp.tobytes() and np.array_equal(desc3, ref_desc)
    ex.close()


def test_matcher_vs_oracle(orbb, oracle):
    rng = np.random.default_rng(3)
    ex = orbb.ORBextractor(500, 1.2, 4, 20, 7, width=320, height=240)
    for nq, nt in ((1, 1), (1, 2), (7, 300), (1000, 1000), (257, 5000), (3000, 129), (33, 50000)):
        t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
        q = t[rng.integers(0, nt, size=nq)].copy()
        q ^= (rng.integers(0, 256, size=q.shape, dtype=np.uint8) & rng.integers(0, 256, size=q.shape, dtype=np.uint8)
              & rng.integers(0, 256, size=q.shape, dtype=np.uint8))
        if nt > 10:
            t[5] = t[3]  # exact duplicates: tie -> lowest train index
        for k in (1, 2):
            idx, dist, acc, nacc = orbb.match_knn_host(ex, q, t, k=k, ratio=0.7)
            oidx, odist, oacc = oracle.match_knn(q, t, k=k, ratio=0.7)
            if k == 1:
                oidx[:, 1] = -1; odist[:, 1] = -1
            assert np.array_equal(idx, oidx), (nq, nt, k)
            assert np.array_equal(dist, odist), (nq, nt, k)
            assert np.array_equal(acc, oacc) and nacc == int(oacc.sum())
    ex.close()


def test_matcher_segmented(orbb, oracle):
    import torch
    rng = np.random.default_rng(4)
    ex = orbb.ORBextractor(500, 1.2, 4, 20, 7, width=320, height=240)
    nqs, nts = [100, 0, 257, 31], [90, 50, 400, 1]
    q = rng.integers(0, 256, size=(sum(nqs), 32), dtype=np.uint8)
    t = rng.integers(0, 256, size=(sum(nts), 32), dtype=np.uint8)
    qo = np.concatenate([[0], np.cumsum(nqs)]).astype(np.int32)
    to = np.concatenate([[0], np.cumsum(nts)]).astype(np.int32)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    dqo, dto = torch.from_numpy(qo).cuda(), torch.from_numpy(to).cuda()
    idx = torch.zeros((len(q), 2), dtype=torch.int32, device="cuda")
    dist = torch.zeros((len(q), 2), dtype=torch.int32, device="cuda")
    acc = torch.zeros(len(q), dtype=torch.uint8, device="cuda")
    ex.match_keypoints_segmented(dq, dqo, dt, dto, 4, sum(nqs), max(nqs), max(nts), idx, dist, acc, k=2, ratio=0.8,
                                 stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    idx, dist, acc = idx.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy().astype(bool)
    for s in range(4):
        if nqs[s] == 0:
            continue
        oi, od, oa = oracle.match_knn(q[qo[s]:qo[s + 1]], t[to[s]:to[s + 1]], k=2, ratio=0.8)
        assert np.array_equal(idx[qo[s]:qo[s + 1]], oi) and np.array_equal(dist[qo[s]:qo[s + 1]], od)
        assert np.array_equal(acc[qo[s]:qo[s + 1]], oa)
    ex.close()


def test_errors_are_codes_not_aborts(orbb):
    with pytest.raises(orbb.OrbbError):
        orbb.ORBextractor(1000, 1.2, 8, 20, 7, width=120, height=100)  # level 7 under 62 px
    with pytest.raises(orbb.OrbbError):
        orbb.ORBextractor(1000, 1.0, 8, 20, 7, width=640, height=480)  # scale factor must be > 1
    ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=200, height=150, max_batch=1)
    with pytest.raises(orbb.OrbbError):
        ex.extract_batch(np.zeros((2, 150, 200), np.uint8))  # over batch capacity
    with pytest.raises(orbb.OrbbError):
        ex(np.zeros((100, 100), np.uint8))  # wrong shape
    ex.close()


def test_full_size_properties(orbb, synth):
    """BASELINE config sizes where the oracle is too slow to run per frame: size-independent properties --
    batch determinism, per-level quotas respected, keypoints inside the image, descriptor self-match."""
    base = np.stack([synth.textured_frame(848, 480, 2000 + i) for i in range(4)])
    frames = np.concatenate([base] * 16)  # 64 frames
    ex = orbb.ORBextractor(1200, 1.2, 8, 20, 7, width=848, height=480, max_batch=64)
    kp, desc, counts = ex.extract_batch(frames)
    assert (counts >= 1200).all() and (counts <= 1200 + 2 * 8 + 8).all()
    for f in range(4, 64):  # repeated inputs -> identical outputs regardless of batch slot
        n = counts[f]
        assert n == counts[f % 4] and kp[f, :n].tobytes() == kp[f % 4, :n].tobytes()
        assert desc[f, :n].tobytes() == desc[f % 4, :n].tobytes()
    k0 = kp[0, :counts[0]]
    assert (k0["x"] >= 19).all() and (k0["x"] < 848).all() and (k0["y"] >= 19).all() and (k0["y"] < 480).all()
    per_level = np.bincount(k0["octave"], minlength=8)
    assert (per_level <= ex.features_per_level + 2).all()
    idx, dist, acc, _ = orbb.match_knn_host(ex, desc[0, :counts[0]], desc[4, :counts[4]], k=1)
    assert (dist[:, 0] == 0).all()
    ex.close()


def test_cpp_host_mirror(orbb, oracle, synth, tmp_path):
    """include/orbb200.hpp (the C++ mirror a reference maintainer would include) gives the same bytes."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_mirror"
    libdir = os.path.join(root, "jetracer-orbslam2_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "host_mirror_main.cpp"), "-o", str(exe), "-L" + libdir,
                    "-lorbb200", "-Wl,-rpath," + libdir], check=True)
    img = synth.textured_frame(640, 480, 4242)
    raw, out = tmp_path / "frame.raw", tmp_path / "out.bin"
    raw.write_bytes(img.tobytes())
    subprocess.run([str(exe), str(raw), "640", "480", str(out)], check=True)
    blob = out.read_bytes()
    n = int(np.frombuffer(blob[:4], np.int32)[0])
    kp = np.frombuffer(blob[4:4 + 28 * n], orbb.KEYPOINT_DTYPE)
    desc = np.frombuffer(blob[4 + 28 * n:], np.uint8).reshape(n, 32)
    okp, odesc = canon(*oracle.Oracle(640, 480).extract(img))
    kp, desc = canon(kp, desc)
    assert n == len(okp) and kp.tobytes() == okp.tobytes() and np.array_equal(desc, odesc)


def test_cfg3_stereo_and_temporal_matching(orbb, oracle, synth):
    """cfg 3: left/right (horizontal disparity) and t/t+1 (translation) pairs, 2-NN + 0.7 ratio test through the
    segmented matcher; extraction and matches must equal the oracle's."""
    import torch
    w, h = 848, 800
    left = synth.textured_frame(w, h, 3100)
    right = synth.shifted_frame(left, -24, 0, 3101)
    nxt = synth.shifted_frame(left, 3, 1, 3102)
    frames = np.stack([left, right, nxt])
    ex = orbb.ORBextractor(1000, 1.2, 8, 20, 7, width=w, height=h, max_batch=3)
    kp, desc, counts = ex.extract_batch(frames)
    o = oracle.Oracle(w, h, 1000)
    odesc = []
    for f in range(3):
        okp, od = canon(*o.extract(frames[f]))
        gk, gd = canon(kp[f, :counts[f]], desc[f, :counts[f]])
        assert len(okp) == counts[f] and gk.tobytes() == okp.tobytes() and np.array_equal(gd, od)
        odesc.append(desc[f, :counts[f]].copy())  # GPU order == what the matcher sees
    # segments: (left -> right), (left -> next)
    q = np.concatenate([odesc[0], odesc[0]])
    t = np.concatenate([odesc[1], odesc[2]])
    qo = np.array([0, len(odesc[0]), 2 * len(odesc[0])], np.int32)
    to = np.array([0, len(odesc[1]), len(odesc[1]) + len(odesc[2])], np.int32)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    idx = torch.zeros((len(q), 2), dtype=torch.int32, device="cuda")
    dist = torch.zeros((len(q), 2), dtype=torch.int32, device="cuda")
    acc = torch.zeros(len(q), dtype=torch.uint8, device="cuda")
    ex.match_keypoints_segmented(dq, torch.from_numpy(qo).cuda(), dt, torch.from_numpy(to).cuda(), 2, len(q), len(odesc[0]),
                                 max(len(odesc[1]), len(odesc[2])), idx, dist, acc, k=2, ratio=0.7,
                                 stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    idx, dist, acc = idx.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy().astype(bool)
    for s_ in range(2):
        oi, od, oa = oracle.match_knn(q[qo[s_]:qo[s_ + 1]], t[to[s_]:to[s_ + 1]], k=2, ratio=0.7)
        assert np.array_equal(idx[qo[s_]:qo[s_ + 1]], oi) and np.array_equal(dist[qo[s_]:qo[s_ + 1]], od)
        assert np.array_equal(acc[qo[s_]:qo[s_ + 1]], oa)
        assert oa.sum() > 50  # the synthetic pair really has correspondences
    ex.close()


def test_async_host_api_matches_sync(orbb, synth):
    """orbb_extract_batch_host_async/orbb_wait with two batches in flight == the blocking call."""
    import torch
    frames = [np.stack([synth.textured_frame(424, 240, 700 + 10 * b + i) for i in range(40)]) for b in range(3)]
    ex = orbb.ORBextractor(600, 1.2, 6, 20, 7, width=424, height=240, max_batch=40)
    ref = [ex.extract_batch(f) for f in frames]
    pins = [torch.from_numpy(f).pin_memory() for f in frames]
    outs = [(torch.zeros(40 * ex.max_kp * 28, dtype=torch.uint8).pin_memory(),
             torch.zeros(40 * ex.max_kp * 32, dtype=torch.uint8).pin_memory(),
             torch.zeros(40, dtype=torch.int32).pin_memory()) for _ in range(3)]
    tickets = []
    for rep in range(2):  # 6 submissions: the ticket ring wraps
        for b in range(3):
            k, d, c = outs[b]
            tickets.append(ex.extract_batch_host_async(pins[b].data_ptr(), 424, 424 * 240, 40, k.data_ptr(), d.data_ptr(),
                                                       c.data_ptr()))
            if len(tickets) >= 2:
                ex.wait(tickets[-2])
    ex.wait(tickets[-1])
    assert tickets == list(range(tickets[0], tickets[0] + 6))  # earlier blocking calls consumed tickets too
    for b in range(3):
        k, d, c = outs[b]
        rk, rd, rc = ref[b]
        assert np.array_equal(c.numpy(), rc)
        kk = np.frombuffer(k.numpy().tobytes(), orbb.KEYPOINT_DTYPE).reshape(40, ex.max_kp)
        dd = d.numpy().reshape(40, ex.max_kp, 32)
        for f in range(40):
            n = rc[f]
            assert kk[f, :n].tobytes() == rk[f, :n].tobytes() and dd[f, :n].tobytes() == rd[f, :n].tobytes()
    ex.close()


def test_windowed_matcher_reference_semantics(orbb, oracle, synth):
    """orbb_match_windowed (the reference's match_keypoints gate: +-max_px window, best Hamming < cutoff) vs oracle,
    on real keypoints of frame t / t+1 (keypoint array passed directly, stride 28) and on random data."""
    import torch
    w, h = 640, 480
    f0 = synth.textured_frame(w, h, 5100)
    f1 = synth.shifted_frame(f0, 3, 1, 5101)
    ex = orbb.ORBextractor(1000, 1.2, 8, 20, 7, width=w, height=h, max_batch=2)
    kp, desc, counts = ex.extract_batch(np.stack([f0, f1]))
    n0, n1 = int(counts[0]), int(counts[1])
    st = torch.cuda.current_stream()
    for max_px, max_ham in ((2.0, 4), (6.0, 64), (40.0, 257), (0.0, 30)):
        dq = torch.from_numpy(desc[0, :n0].copy()).cuda()
        dt = torch.from_numpy(desc[1, :n1].copy()).cuda()
        kq = torch.from_numpy(np.frombuffer(kp[0, :n0].tobytes(), np.uint8).copy()).cuda()   # orbb_keypoint[], stride 28
        kt = torch.from_numpy(np.frombuffer(kp[1, :n1].tobytes(), np.uint8).copy()).cuda()
        idx = torch.zeros(n0, dtype=torch.int32, device="cuda")
        dist = torch.zeros(n0, dtype=torch.int32, device="cuda")
        nm = torch.zeros(1, dtype=torch.int32, device="cuda")
        ex.match_keypoints_windowed(dq, kq, 28, n0, dt, kt, 28, n1, max_px, max_ham, idx, dist, nm, stream=st)
        torch.cuda.synchronize()
        qxy = np.stack([kp[0, :n0]["x"], kp[0, :n0]["y"]], 1)
        txy = np.stack([kp[1, :n1]["x"], kp[1, :n1]["y"]], 1)
        oi, od, on = oracle.match_windowed(desc[0, :n0], qxy, desc[1, :n1], txy, max_px, max_ham)
        assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
        assert int(nm.item()) == on
        if max_px == 6.0:
            assert on > 200  # the (3,1) shift really produces windowed matches
    rng = np.random.default_rng(9)
    nq, nt = 777, 1301
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    t[:300] = q[:300] ^ (rng.integers(0, 256, size=(300, 32), dtype=np.uint8) & 0x11)
    qxy = rng.uniform(0, 200, size=(nq, 2)).astype(np.float32)
    txy = rng.uniform(0, 200, size=(nt, 2)).astype(np.float32)
    txy[:300] = qxy[:300] + rng.integers(-3, 4, size=(300, 2)).astype(np.float32)
    idx = torch.zeros(nq, dtype=torch.int32, device="cuda")
    dist = torch.zeros(nq, dtype=torch.int32, device="cuda")
    ex.match_keypoints_windowed(torch.from_numpy(q).cuda(), torch.from_numpy(qxy).cuda(), 8, nq, torch.from_numpy(t).cuda(),
                                torch.from_numpy(txy).cuda(), 8, nt, 3.0, 100, idx, dist, stream=st)
    torch.cuda.synchronize()
    oi, od, _ = oracle.match_windowed(q, qxy, t, txy, 3.0, 100)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    ex.close()


def test_degenerate_inputs(orbb, oracle):
    """Inputs that stress capacities and tie handling: uniform noise (maximal candidate density -> tens of
    thousands of candidates per level, quadtree on the global-scratch path), salt & pepper, smooth gradients
    (threshold-7 fallback everywhere), 2-px and 3-px checkerboards (massive NMS ties), constant frames."""
    rng = np.random.default_rng(2025)
    w, h = 640, 480
    yy, xx = np.mgrid[0:h, 0:w]
    sp = np.full((h, w), 128, np.uint8)
    m = rng.random((h, w))
    sp[m < 0.02] = 0
    sp[m > 0.98] = 255
    frames = np.stack([
        rng.integers(0, 256, size=(h, w), dtype=np.uint8),                       # uniform noise
        sp,                                                                        # salt & pepper
        ((xx * 255) // (w - 1)).astype(np.uint8),                                  # horizontal ramp
        ((xx + 2 * yy) % 256).astype(np.uint8),                                    # diagonal saw-tooth
        np.where(((yy // 2) + (xx // 2)) % 2 == 0, 40, 215).astype(np.uint8),      # 2-px checkerboard
        np.where(((yy // 3) + (xx // 3)) % 2 == 0, 90, 160).astype(np.uint8),      # 3-px checkerboard
        np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8),                # constant
    ])
    ex = orbb.ORBextractor(1000, 1.2, 8, 20, 7, width=w, height=h, max_batch=len(frames))
    kp, desc, counts = ex.extract_batch(frames)
    o = oracle.Oracle(w, h, 1000)
    for f in range(len(frames)):
        okp, odesc = canon(*o.extract(frames[f]))
        gk, gd = canon(kp[f, :counts[f]], desc[f, :counts[f]])
        assert len(okp) == counts[f], (f, len(okp), counts[f])
        assert gk.tobytes() == okp.tobytes(), f
        assert np.array_equal(gd, odesc), f
    # candidate sets of the densest frame, level 0 and last level
    o.compute_pyramid(frames[0])
    for l in (0, 7):
        oc = o.level_candidates(l)
        assert as_set(ex.debug_candidates(l, frame=0)) == as_set(np.stack([oc["x"], oc["y"], oc["response"]], 1))
    ex.close()


def test_noise_1280x720_over_64k_candidates(orbb, oracle):
    """Level 0 of a 1280x720 noise frame yields > 65535 candidates: exercises 32-bit indices in the quadtree sort."""
    rng = np.random.default_rng(77)
    img = rng.integers(0, 256, size=(720, 1280), dtype=np.uint8)
    ex = orbb.ORBextractor(2000, 1.2, 8, 20, 7, width=1280, height=720)
    o = oracle.Oracle(1280, 720, 2000)
    gk, gd = canon(*ex(img))
    okp, od = canon(*o.extract(img))
    assert len(ex.debug_candidates(0)) == len(o.level_candidates(0)) > 65535
    assert len(gk) == len(okp) and gk.tobytes() == okp.tobytes() and np.array_equal(gd, od)
    ex.close()


def test_cuda_graph_replay_opt_in(orbb, oracle, synth, monkeypatch):
    """The small-batch device entry point replays a captured CUDA graph (default since round 2; ORBB_GRAPH=1 forces it):
    results stay bit-exact across replays, across alternating buffer sets (graph cache), after evictions (5 distinct
    argument sets > 4 slots: the evicted executable graph is patched in place) and after the fallback to plain launches."""
    import torch
    monkeypatch.setenv("ORBB_GRAPH", "1")
    w, h = 320, 240
    ex = orbb.ORBextractor(400, 1.2, 6, 20, 7, width=w, height=h, max_batch=2)
    monkeypatch.delenv("ORBB_GRAPH")
    st = torch.cuda.current_stream()
    o = oracle.Oracle(w, h, 400, 1.2, 6, 20, 7)
    sets = []
    for k in range(5):
        frames = np.stack([synth.textured_frame(w, h, 600 + 2 * k), synth.textured_frame(w, h, 601 + 2 * k)])
        sets.append((frames, torch.from_numpy(frames).cuda(), torch.zeros((2, ex.max_kp, 7), dtype=torch.float32, device="cuda"),
                     torch.zeros((2, ex.max_kp, 32), dtype=torch.uint8, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda")))
    l0 = ex.launch_count()
    for rep in range(3):
        for frames, d_in, d_kp, d_desc, d_cnt in sets:
            d_kp.zero_(); d_desc.zero_(); d_cnt.zero_()
            ex.extract_batch_device(d_in, 2, d_kp, d_desc, d_cnt, stream=st)
            torch.cuda.synchronize()
            kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(2, ex.max_kp)
            desc = d_desc.cpu().numpy(); cnt = d_cnt.cpu().numpy()
            for f in range(2):
                okp, odesc = canon(*o.extract(frames[f]))
                gkp, gdesc = canon(kp[f, :cnt[f]], desc[f, :cnt[f]])
                assert len(gkp) == len(okp)
                assert np.array_equal(gkp.view(np.uint8), okp.view(np.uint8)) and np.array_equal(gdesc, odesc)
    # replays are counted like direct launches: level0 + 5 resizes + (FAST + quadtree) x 3 level groups (the captured
    # graph runs the detection of level 0 and of levels 1..3 next to the rest of the pyramid) + blur + angle/rBRIEF = 14.  Five argument
    # sets round-robin over four slots miss every time: eight captures / in-place updates, then the handle falls back
    # to plain launches (10 per call) for the remaining seven calls
    assert ex.launch_count() - l0 == 8 * 14 + 7 * 10


@pytest.mark.parametrize("nb", [12, 30, 64, 71, 130])
def test_large_batch_paths_equal_single_frame(orbb, synth, nb):
    """Batches >= 64 frames take the multi-stream split with the throughput variants of the kernels (large-grid
    quadtree kernel, two batch parts); a single frame takes the small-grid variants (shared-memory per-key arrays,
    8 loads in flight, one keypoint per warp in the angle/rBRIEF kernel); 12 and 30 frames land on the variants in
    between (quadtree grids of 96 and 240 CTAs, 2 and 8 keypoint slots per warp).  All must give the same bytes, in
    the same order."""
    import torch
    w, h = 320, 240
    base = [synth.textured_frame(w, h, 300 + i) for i in range(6)] + [synth.sparse_frame(w, h, 9), synth.low_contrast_frame(w, h, 4)]
    frames = np.stack([np.roll(base[i % 8], (3 * (i // 8), 5 * (i // 8)), axis=(0, 1)) for i in range(nb)])
    ex = orbb.ORBextractor(500, 1.2, 8, 20, 7, width=w, height=h, max_batch=nb)
    st = torch.cuda.current_stream()
    d_in = torch.from_numpy(frames).cuda()
    d_kp = torch.zeros((nb, ex.max_kp, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((nb, ex.max_kp, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(nb, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_in, nb, d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    kp = d_kp.cpu().numpy().view(orbb.KEYPOINT_DTYPE).reshape(nb, ex.max_kp); desc = d_desc.cpu().numpy(); cnt = d_cnt.cpu().numpy()
    one = orbb.ORBextractor(500, 1.2, 8, 20, 7, width=w, height=h, max_batch=1)
    for f in list(range(0, nb, 9)) + [nb - 1]:
        k1, d1 = one(frames[f])
        n = int(cnt[f])
        assert len(k1) == n and k1.tobytes() == kp[f, :n].tobytes() and d1.tobytes() == desc[f, :n].tobytes(), f"frame {f}"
    ex.close(); one.close()


def test_reference_gpu_kernels_baseline_runs(synth, tmp_path):
    """The timing baseline built from the reference's own .cu files (oracle/_ref/ref_gpu_bench, `make -C oracle
    ref_gpu`) runs on this GPU and reports one keypoint per 32x32 cell.  It is a different algorithm (FAST-12, one
    level, 32-bit hashes), so only its well-formedness is checked here; bench.py reports its timings."""
    import json
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_gpu_bench")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gpu_bench not built (needs /root/reference at build time)")
    w, h = 848, 480
    raw = tmp_path / "frame.raw"
    synth.textured_frame(w, h, 2000).tofile(raw)
    r = subprocess.run([exe, str(raw), str(w), str(h), "20"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-400:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cells"] == 27 * 15 and 0 < d["keypoints"] <= d["cells"]
    assert d["frame_latency_us"] > 0 and set(d["stages_us"]) >= {"detect", "calc_orb", "compute_fast_angle"}


@pytest.mark.gpu
def test_every_matcher_kernel_reproduces_the_popc_matcher():
    """The brute-force matcher has four kernels behind one interface, chosen once per process from the environment:
    tcgen05 with the pre-expanded train image (default), tcgen05 expanding its tiles in the CTAs (ORBB_MATCH_PRE=0; also
    what small and segmented calls use), tcgen05 with every pair keyed (ORBB_MATCH_UMMA=2), warp-level int8 MMA
    (ORBB_MATCH_UMMA=0) and XOR / POPC (ORBB_MATCH_POPC=1).  tools/umma_probe.py runs each in a child process on the same
    inputs -- full and partial train tiles, split-T, duplicate train rows, k = 1 and 2, a 262 144-row map -- and compares
    every result array with the POPC kernel's (which the other matcher tests check against the oracle)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, UMMA_PROBE_NOTIME="1")
    for k in ("ORBB_MATCH_UMMA", "ORBB_MATCH_PRE", "ORBB_MATCH_POPC", "ORBB_MATCH_CTAS", "UMMA_PROBE_ONLY"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "umma_probe.py")], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("result arrays identical to the POPC matcher") == 4, r.stdout
    assert "MISMATCH" not in r.stdout and "TIMEOUT" not in r.stdout

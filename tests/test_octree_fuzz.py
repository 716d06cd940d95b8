"""DistributeOctTree: the C oracle (oracle/orb_oracle.c) against the independent Python restatement that generates the
golden fixtures (tests/golden/gen_golden.py::PyOrbExtractor.distribute), fuzzed with hypothesis over random candidate
clouds, quotas, response ties and geometries -- far more cases than the 9 committed fixtures.  Both are restatements of
upstream ORB-SLAM2; no third-party implementation of this function exists in the image, so this is the
strongest pin available: two independently written programs (a literal linked-list walk in Python, an array-based one
in C) must return the same keys IN THE SAME ORDER.  CPU only."""
import importlib.util
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def pyorb():
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(HERE, "golden", "gen_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pytest.importorskip("cv2")
    return mod.PyOrbExtractor(1000, 1.2, 8, 20, 7)


def _upstream_order(xy, W, H):
    """candidate order upstream feeds DistributeOctTree: 30-px cells row-major, row-major inside a cell"""
    w_cell = int(np.ceil(np.float32(W) / np.float32(int(np.float32(W) / np.float32(30)))))
    h_cell = int(np.ceil(np.float32(H) / np.float32(int(np.float32(H) / np.float32(30)))))
    i, j = np.maximum(xy[:, 1] - 3, 0) // h_cell, np.maximum(xy[:, 0] - 3, 0) // w_cell
    return np.lexsort((xy[:, 0], xy[:, 1], j, i))


def _run_both(oracle, pyorb, xy, resp, W, H, quota):
    order = _upstream_order(xy, W, H)
    xy, resp = xy[order], resp[order]
    c = np.zeros(len(xy), oracle.CAND_DTYPE)
    c["x"], c["y"], c["response"] = xy[:, 0], xy[:, 1], resp
    sel = oracle.distribute_octree(c, 16, 16 + W, 16, 16 + H, quota)
    got = [(int(c["x"][s]), int(c["y"][s]), int(c["response"][s])) for s in sel]
    keys = [(int(x), int(y), int(r)) for (x, y), r in zip(xy, resp)]
    ref = pyorb.distribute(keys, 16, 16 + W, 16, 16 + H, quota)
    return got, [tuple(k) for k in ref]


GEOMS = [(608, 448), (816, 448), (1248, 688), (147, 102), (816, 768), (225, 161), (1888, 1048)]


@st.composite
def clouds(draw):
    W, H = draw(st.sampled_from(GEOMS))
    n = draw(st.integers(1, 900))
    kind = draw(st.sampled_from(["uniform", "clustered", "lines", "grid"]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        xy = np.stack([rng.integers(3, W - 3, n), rng.integers(3, H - 3, n)], 1)
    elif kind == "clustered":
        k = draw(st.integers(1, 6))
        cx, cy = rng.integers(3, W - 3, k), rng.integers(3, H - 3, k)
        pick = rng.integers(0, k, n)
        xy = np.stack([np.clip(cx[pick] + rng.normal(0, draw(st.sampled_from([2, 8, 30])), n).astype(int), 3, W - 4),
                       np.clip(cy[pick] + rng.normal(0, 6, n).astype(int), 3, H - 4)], 1)
    elif kind == "lines":
        xy = np.stack([rng.integers(3, W - 3, n), np.full(n, int(rng.integers(3, H - 3)))], 1)
        if rng.random() < 0.5:
            xy = np.stack([np.full(n, int(rng.integers(3, W - 3))), rng.integers(3, H - 3, n)], 1)
    else:
        step = draw(st.sampled_from([2, 7, 16, 31]))
        gx, gy = np.meshgrid(np.arange(3, W - 3, step), np.arange(3, H - 3, step))
        xy = np.stack([gx.ravel(), gy.ravel()], 1)
        xy = xy[rng.permutation(len(xy))[:n]]
    xy = np.unique(xy.astype(np.int64), axis=0)  # candidates are distinct pixels
    n_resp = draw(st.sampled_from([1, 2, 5, 248]))  # few distinct responses => many ties inside a node
    resp = rng.integers(7, 7 + n_resp, len(xy))
    quota = draw(st.one_of(st.integers(1, 40), st.integers(40, 500), st.sampled_from([len(xy), len(xy) + 5, 2 * len(xy) + 1])))
    return W, H, xy, resp, int(quota)


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture,
                                                                  HealthCheck.data_too_large], derandomize=True)
@given(case=clouds())
def test_distribute_octree_c_vs_python(oracle, pyorb, case):
    W, H, xy, resp, quota = case
    got, ref = _run_both(oracle, pyorb, xy, resp, W, H, quota)
    assert got == ref, f"{W}x{H} n={len(xy)} quota={quota}: {len(got)} vs {len(ref)} keys"


def test_distribute_octree_real_candidates(oracle, pyorb, synth):
    """the candidate clouds of real frames at every level, with the real quotas and a few others"""
    for w, h, seed in ((640, 480, 501), (848, 480, 502)):
        img = synth.textured_frame(w, h, seed)
        o = oracle.Oracle(w, h, 1000)
        o.extract(img)
        for lvl in range(8):
            cand = o.level_candidates(lvl)
            W, H = int(o.lw[lvl]) - 32, int(o.lh[lvl]) - 32
            xy = np.stack([cand["x"], cand["y"]], 1).astype(np.int64)
            for quota in (int(o.nfeat[lvl]), 17, 3 * int(o.nfeat[lvl])):
                got, ref = _run_both(oracle, pyorb, xy, cand["response"].astype(np.int64), W, H, quota)
                assert got == ref, (w, h, lvl, quota)

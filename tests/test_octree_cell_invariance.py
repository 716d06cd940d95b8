"""CPU check of the claim behind the quadtree kernel's cell-table fast path ("Why records can stand
in for keys"), made against the ORACLE only -- no GPU, no restatement of the kernel:

    as long as DistributeOctTree never splits a node below depth Dc, its selection depends on the candidates only
    through, per tree cell of depth Dc, (a) the NUMBER of candidates in the cell and (b) the cell's BEST candidate
    (highest response, earliest in upstream candidate order).

So moving every non-best candidate to another pixel of its own cell (and lowering its response) must not change
the selected set.  Tree cells follow upstream's root nodes and DivideNode: nIni = round(W / H) roots of
width hX = W / nIni, halves ceil((hi - lo) / 2), a key goes left/up iff its coordinate is below the split.  The
negative control uses a table that is far too shallow for the quota, where the same move must change the result."""
import numpy as np
import pytest


def _axis_cells(size, depth, n_roots=1):
    """cell index and [lo, hi) of every coordinate 0..size-1 after `depth` DivideNode halvings per root"""
    idx = np.zeros(size, np.int64)
    lo_of = np.zeros(size, np.int64)
    hi_of = np.zeros(size, np.int64)
    hx = np.float32(size) / np.float32(n_roots)
    for x in range(size):
        root = min(int(np.float32(x) / hx), n_roots - 1)
        lo, hi = int(hx * np.float32(root)), int(hx * np.float32(root + 1))
        bits = 0
        for _ in range(depth):
            mid = lo + int(np.ceil(np.float32(hi - lo) / np.float32(2)))
            if x < mid:
                hi, bits = mid, bits << 1
            else:
                lo, bits = mid, (bits << 1) | 1
        idx[x] = (root << depth) | bits
        lo_of[x], hi_of[x] = lo, hi
    return idx, lo_of, hi_of


def _canonical_order(xy, W, H):
    """upstream candidate order: 30-px FAST cells row-major, FAST's row-major order inside a cell"""
    w_cell = int(np.ceil(np.float32(W) / np.float32(int(np.float32(W) / np.float32(30)))))
    h_cell = int(np.ceil(np.float32(H) / np.float32(int(np.float32(H) / np.float32(30)))))
    i, j = (xy[:, 1] - 3) // h_cell, (xy[:, 0] - 3) // w_cell
    return np.lexsort((xy[:, 0], xy[:, 1], j, i))


def _select(oracle, xy, resp, W, H, quota):
    order = _canonical_order(xy, W, H)
    c = np.zeros(len(xy), oracle.CAND_DTYPE)
    c["x"], c["y"], c["response"] = xy[order, 0], xy[order, 1], resp[order]
    sel = oracle.distribute_octree(c, 16, 16 + W, 16, 16 + H, quota)
    return {(int(c["x"][s]), int(c["y"][s]), int(c["response"][s])) for s in sel}


def _shuffle_inside_cells(xy, resp, W, H, depth, rng):
    """every candidate that is not the best of its depth-`depth` tree cell moves to a free pixel of the same cell and
    gets a response below the cell's best"""
    n_roots = int(round(float(np.float32(W) / np.float32(H))))
    cx, xlo, xhi = _axis_cells(W, depth, n_roots)
    cy, ylo, yhi = _axis_cells(H, depth, 1)
    cell = cx[xy[:, 0]] * (1 << depth) + cy[xy[:, 1]]
    order = _canonical_order(xy, W, H)
    rank = np.empty(len(xy), np.int64)
    rank[order] = np.arange(len(xy))
    new_xy, new_resp = xy.copy(), resp.copy()
    moved = 0
    for cid in np.unique(cell):
        members = np.flatnonzero(cell == cid)
        # best: highest response, earliest in upstream order
        best = members[np.lexsort((rank[members], -resp[members]))[0]]
        x0, y0 = xy[best]
        free = [(x, y) for x in range(max(int(xlo[x0]), 3), min(int(xhi[x0]), W - 3))
                for y in range(max(int(ylo[y0]), 3), min(int(yhi[y0]), H - 3)) if (x, y) != (x0, y0)]
        rng.shuffle(free)
        for k, m in enumerate(mm for mm in members if mm != best):
            new_xy[m] = free[k]
            new_resp[m] = int(rng.integers(1, resp[best]))
            moved += 1
    return new_xy, new_resp, moved


CASES = [  # (W, H, quota, table depth): the product's rule picks the shallowest depth with >= 4 x quota cells
    (608, 448, 217, 5),    # 640x480 level 0: 1 root, 1024 cells
    (816, 448, 261, 5),    # 848x480 level 0: 2 roots, 2048 cells
    (1248, 688, 434, 5),   # 1280x720 level 0: 2 roots, 2048 cells
    (412, 302, 126, 5),    # a middle level: 1 root
]


@pytest.mark.parametrize("W,H,quota,depth", CASES)
def test_selection_depends_on_cell_counts_and_cell_bests_only(oracle, W, H, quota, depth):
    rng = np.random.default_rng(W * 7 + quota)
    pts = set()
    while len(pts) < W * H // 45:  # about the candidate density of a textured frame
        pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
    xy = np.array(sorted(pts), np.int64)
    resp = rng.integers(60, 200, size=len(xy))  # many ties
    ref = _select(oracle, xy, resp, W, H, quota)
    assert len(ref) >= quota
    xy2, resp2, moved = _shuffle_inside_cells(xy, resp, W, H, depth, rng)
    assert moved > len(xy) // 2
    assert _select(oracle, xy2, resp2, W, H, quota) == ref


def test_negative_control_shallow_table_changes_the_selection(oracle):
    """the same move inside cells that are far too coarse for the quota (16 cells, quota 217) changes the result:
    the invariance above is a property of the depth bound, not of the test"""
    W, H, quota = 608, 448, 217
    rng = np.random.default_rng(5)
    pts = set()
    while len(pts) < 6000:
        pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
    xy = np.array(sorted(pts), np.int64)
    resp = rng.integers(60, 200, size=len(xy))
    ref = _select(oracle, xy, resp, W, H, quota)
    xy2, resp2, _ = _shuffle_inside_cells(xy, resp, W, H, 2, rng)
    assert _select(oracle, xy2, resp2, W, H, quota) != ref

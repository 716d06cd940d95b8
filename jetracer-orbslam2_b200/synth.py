"""Seeded synthetic inputs for the ORB front-end (numpy only; harness code, not the hot path).

SURVEY.md section 8(d): "textured frame" = multi-octave value noise + random filled
rectangles/discs + uniform pixel noise, which gives several thousand FAST-20 candidates at level 0
so every pyramid level saturates its feature quota.  Degenerate frames (flat, checkerboard)
exercise the threshold-7 fallback and NMS ties.
"""
from __future__ import annotations

import numpy as np


def _upsample_bilinear(grid: np.ndarray, h: int, w: int, cell: int) -> np.ndarray:
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = np.floor(ys).astype(np.int32)
    x0 = np.floor(xs).astype(np.int32)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    g00 = grid[y0][:, x0]
    g01 = grid[y0][:, x0 + 1]
    g10 = grid[y0 + 1][:, x0]
    g11 = grid[y0 + 1][:, x0 + 1]
    return (g00 * (1 - fx) + g01 * fx) * (1 - fy) + (g10 * (1 - fx) + g11 * fx) * fy


def textured_frame(width: int, height: int, seed: int) -> np.ndarray:
    """One u8 gray frame, deterministic in (width, height, seed).

    Shape and speck counts scale with the area so candidate density (about 5k FAST candidates at
    level 0 of a 640x480 frame, every level well above its quota) is resolution independent."""
    rng = np.random.default_rng(seed)
    area = (width * height) / float(640 * 480)
    acc = np.zeros((height, width), np.float32)
    for cell, wgt in ((64, 1.0), (32, 0.5), (16, 0.25), (8, 0.125), (4, 0.125)):
        grid = rng.random((height // cell + 3, width // cell + 3), dtype=np.float32)
        acc += wgt * _upsample_bilinear(grid, height, width, cell)
    acc = (acc - acc.min()) / max(float(acc.max() - acc.min()), 1e-6)
    img = 30.0 + 195.0 * acc
    yy, xx = np.mgrid[0:height, 0:width]
    for _ in range(max(8, int(100 * area))):
        gray = float(rng.integers(0, 256))
        cx, cy = int(rng.integers(0, width)), int(rng.integers(0, height))
        if rng.random() < 0.5:
            rw, rh = int(rng.integers(3, 40)), int(rng.integers(3, 30))
            img[max(0, cy - rh):cy + rh, max(0, cx - rw):cx + rw] = gray
        else:
            r = int(rng.integers(3, 30))
            img[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = gray
    for _ in range(int(6000 * area)):  # small high-contrast specks: dense corner field
        cx, cy = int(rng.integers(2, width - 3)), int(rng.integers(2, height - 3))
        s = int(rng.integers(1, 3))
        img[cy - s:cy + s, cx - s:cx + s] += float(rng.integers(-90, 91))
    img += rng.integers(-4, 5, size=img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def flat_frame(width: int, height: int, value: int = 128) -> np.ndarray:
    return np.full((height, width), value, np.uint8)


def checkerboard_frame(width: int, height: int, cell: int = 8, lo: int = 64, hi: int = 192) -> np.ndarray:
    yy, xx = np.mgrid[0:height, 0:width]
    return np.where(((yy // cell) + (xx // cell)) % 2 == 0, lo, hi).astype(np.uint8)


def low_contrast_frame(width: int, height: int, seed: int) -> np.ndarray:
    """Texture squeezed to a ~24-level range: most cells fail FAST-20 and fall back to FAST-7."""
    t = textured_frame(width, height, seed).astype(np.float32)
    return np.clip(np.rint(116.0 + (t - 127.0) * 0.12), 0, 255).astype(np.uint8)


def sparse_frame(width: int, height: int, seed: int, blobs: int = 12) -> np.ndarray:
    """A few isolated squares on a flat background: fewer candidates than the quota."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 100, np.uint8)
    for _ in range(blobs):
        cx, cy = int(rng.integers(30, width - 30)), int(rng.integers(30, height - 30))
        s = int(rng.integers(3, 9))
        img[cy - s:cy + s, cx - s:cx + s] = int(rng.integers(160, 256))
    return img


def frame_batch(width: int, height: int, n: int, seed0: int) -> np.ndarray:
    """[n, height, width] textured frames with seeds seed0 .. seed0+n-1."""
    return np.stack([textured_frame(width, height, seed0 + i) for i in range(n)])


def shifted_frame(img: np.ndarray, dx: int, dy: int, seed: int, noise: int = 3) -> np.ndarray:
    """img translated by (dx,dy) with edge replication + small noise (frame t+1 / right image)."""
    rng = np.random.default_rng(seed)
    h, w = img.shape
    ys = np.clip(np.arange(h) - dy, 0, h - 1)
    xs = np.clip(np.arange(w) - dx, 0, w - 1)
    out = img[ys][:, xs].astype(np.int16) + rng.integers(-noise, noise + 1, size=img.shape).astype(np.int16)
    return np.clip(out, 0, 255).astype(np.uint8)


def rolled_batch(width: int, height: int, n: int, seed0: int, n_base: int = 16) -> np.ndarray:
    """[n, height, width] distinct textured frames, cheap to make: n_base seeded frames, then cyclic shifts of them
    (frame i = base[i % n_base] rolled by (37k, 53k) pixels, k = i // n_base).  bench.py's workload generator."""
    base = [textured_frame(width, height, seed0 + i) for i in range(min(n, n_base))]
    out = np.empty((n, height, width), np.uint8)
    for i in range(n):
        b = base[i % len(base)]
        k = i // len(base)
        out[i] = np.roll(b, (37 * k, 53 * k), axis=(0, 1)) if k else b
    return out


def rolled_frames(width: int, height: int, indices, seed0: int, n_base: int = 16) -> np.ndarray:
    """frames `indices` of the (unbounded) rolled_batch sequence with n >= n_base: what a rank generates for its shard of
    a global batch, identical to slicing rolled_batch(width, height, n, seed0, n_base) for any n > max(indices)."""
    indices = [int(i) for i in indices]
    need = sorted({i % n_base for i in indices})
    base = {b: textured_frame(width, height, seed0 + b) for b in need}
    out = np.empty((len(indices), height, width), np.uint8)
    for j, i in enumerate(indices):
        k = i // n_base
        out[j] = np.roll(base[i % n_base], (37 * k, 53 * k), axis=(0, 1)) if k else base[i % n_base]
    return out

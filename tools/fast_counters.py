#!/usr/bin/env python3
"""Evidence for the FAST kernel's instruction diet: how many pixels of the bench workload pass
each candidate precheck, are FAST corners, and survive NMS -- per pyramid level.  numpy + cv2 only (no GPU, no oracle):
the pyramid is the cv2.resize chain the oracle is pinned to, the tested range is [19, w-19) x [19, h-19) and the cell
seams are ignored for the NMS count (they change it by a few percent, not the precheck statistics).

  python tools/fast_counters.py [--frames 4] [--w 640 --h 480]
"""
import argparse
import importlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0),
        (-3, 1), (-2, 2), (-1, 3)]


def level_stats(img, thr):
    h, w = img.shape
    c = img[19:h - 19, 19:w - 19].astype(np.int16)
    d = np.stack([img[19 + dy:h - 19 + dy, 19 + dx:w - 19 + dx].astype(np.int16) - c for dx, dy in RING])  # ring - centre
    big = np.abs(d) >= thr            # what the SWAR compare on |d|>>1 accepts (superset of |d| > thr)
    bright, dark = d > thr, d < -thr
    pair = lambda m, k: m[k] | m[k + 8]  # noqa: E731
    st = {"pixels": c.size}
    st["pre2_blind"] = int((pair(big, 0) & pair(big, 4)).sum())
    st["pre4_blind"] = int((pair(big, 0) & pair(big, 4) & pair(big, 2) & pair(big, 6)).sum())
    st["pre8_blind"] = int(np.logical_and.reduce([pair(big, k) for k in range(8)]).sum())
    pol2 = (pair(bright, 0) & pair(bright, 4)) | (pair(dark, 0) & pair(dark, 4))
    pol4 = np.logical_and.reduce([pair(bright, k) for k in (0, 2, 4, 6)]) | np.logical_and.reduce([pair(dark, k) for k in (0, 2, 4, 6)])
    pol8 = np.logical_and.reduce([pair(bright, k) for k in range(8)]) | np.logical_and.reduce([pair(dark, k) for k in range(8)])
    st["pre2_polar"], st["pre4_polar"], st["pre8_polar"] = int(pol2.sum()), int(pol4.sum()), int(pol8.sum())
    # exact: 9 contiguous
    def arc9(m):
        mm = np.concatenate([m, m[:8]])
        run = np.ones_like(m[0])
        out = np.zeros_like(m[0])
        for s in range(16):
            out |= np.logical_and.reduce(mm[s:s + 9])
        return out
    corner = arc9(bright) | arc9(dark)
    st["corners"] = int(corner.sum())
    det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    st["nms_whole_image"] = len(det.detect(np.ascontiguousarray(img[16:h - 16, 16:w - 16])))
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--w", type=int, default=640)
    ap.add_argument("--h", type=int, default=480)
    ap.add_argument("--thr", type=int, default=20)
    a = ap.parse_args()
    synth = importlib.import_module("jetracer-orbslam2_b200.synth")
    tot = {}
    per_level = [dict() for _ in range(8)]
    for f in range(a.frames):
        img = synth.textured_frame(a.w, a.h, 1000 + f)
        sf = np.float32(1.0)
        cur = img
        for l in range(8):
            if l:
                sf = np.float32(sf * np.float32(1.2))
                inv = np.float32(np.float32(1.0) / sf)
                cur = cv2.resize(cur, (int(np.rint(np.float32(a.w) * inv)), int(np.rint(np.float32(a.h) * inv))), interpolation=cv2.INTER_LINEAR)
            st = level_stats(cur, a.thr)
            for k, v in st.items():
                per_level[l][k] = per_level[l].get(k, 0) + v
                tot[k] = tot.get(k, 0) + v
    keys = list(tot)
    print(f"{a.frames} frames {a.w}x{a.h}, threshold {a.thr}; fractions of tested pixels")
    print("level " + " ".join(f"{k:>12}" for k in keys))
    for l in range(8):
        p = per_level[l]["pixels"]
        print(f"{l:5d} " + " ".join(f"{per_level[l][k] / p if k != 'pixels' else per_level[l][k] / a.frames:12.4f}" for k in keys))
    p = tot["pixels"]
    print("  all " + " ".join(f"{tot[k] / p if k != 'pixels' else tot[k] / a.frames:12.4f}" for k in keys))


if __name__ == "__main__":
    main()

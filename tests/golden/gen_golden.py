#!/usr/bin/env python3
"""Golden-vector generator: an INDEPENDENT restatement of upstream ORB-SLAM2 `ORBextractor`
written in Python on top of the *real* OpenCV primitives of this image (cv2 4.13):
cv2.resize / copyMakeBorder / FastFeatureDetector / GaussianBlur / fastAtan2.

Purpose (prompt section 3): the reference repo has no compilable CPU extractor
(src_trash1/orb_extractor.cpp:1-11 is a stub) and no golden vectors, so the C oracle
(oracle/orb_oracle.c) is pinned against (a) cv2 primitives directly (tests/test_oracle_vs_cv2.py)
and (b) the end-to-end outputs of this second implementation, committed as tests/golden/*.npz.
Run in the authoring container:  python tests/golden/gen_golden.py
cv2 is only needed to (re)generate; the committed fixtures are what the tests read.
"""
from __future__ import annotations

import hashlib
import importlib
import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

EDGE = 19
HALF_PATCH = 15
PATCH = 31
f32 = np.float32


def load_pattern() -> np.ndarray:
    txt = open(os.path.join(ROOT, "include", "orb_pattern_31.inc")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    v = np.array([int(t) for t in re.findall(r"-?\d+", txt)], np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)


def cv_round(v) -> int:
    return int(np.rint(f32(v)))  # round-half-even on the float32 value


class PyOrbExtractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        import cv2
        self.cv2 = cv2
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = [f32(1.0)]
        for _ in range(1, nlevels):
            sf.append(f32(sf[-1] * f32(scale_factor)))
        self.sf = sf
        self.inv_sf = [f32(f32(1.0) / s) for s in sf]
        factor = f32(f32(1.0) / f32(scale_factor))
        nd = f32(f32(f32(nfeatures) * f32(f32(1) - factor)) / f32(f32(1) - f32(math.pow(float(factor), float(nlevels)))))
        self.nfeat, s = [], 0
        for _ in range(nlevels - 1):
            self.nfeat.append(cv_round(nd))
            s += self.nfeat[-1]
            nd = f32(nd * factor)
        self.nfeat.append(max(nfeatures - s, 0))
        self.pattern = load_pattern()
        umax = [0] * (HALF_PATCH + 1)
        vmax = int(math.floor(HALF_PATCH * math.sqrt(2.0) / 2 + 1))
        vmin = int(math.ceil(HALF_PATCH * math.sqrt(2.0) / 2))
        for v in range(vmax + 1):
            umax[v] = int(np.rint(math.sqrt(HALF_PATCH * HALF_PATCH - v * v)))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax

    # ---------------- ComputePyramid
    def compute_pyramid(self, image):
        cv2 = self.cv2
        self.padded, self.roi = [], []
        for lvl in range(self.nlevels):
            w = cv_round(f32(image.shape[1]) * self.inv_sf[lvl])
            h = cv_round(f32(image.shape[0]) * self.inv_sf[lvl])
            if lvl == 0:
                r = image
            else:
                r = cv2.resize(self.roi[lvl - 1], (w, h), interpolation=cv2.INTER_LINEAR)
            p = cv2.copyMakeBorder(r, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
            self.padded.append(p)
            self.roi.append(np.ascontiguousarray(p[EDGE:-EDGE, EDGE:-EDGE]))

    # ---------------- per-cell FAST
    def level_candidates(self, lvl):
        cv2 = self.cv2
        im = self.roi[lvl]
        rows, cols = im.shape
        minBX = minBY = EDGE - 3
        maxBX, maxBY = cols - EDGE + 3, rows - EDGE + 3
        width, height = f32(maxBX - minBX), f32(maxBY - minBY)
        nCols, nRows = int(width / f32(30)), int(height / f32(30))
        wCell, hCell = int(math.ceil(f32(width / f32(nCols)))), int(math.ceil(f32(height / f32(nRows))))
        det = {t: cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True,
                                                 type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
               for t in (self.ini_th, self.min_th)}
        out = []
        for i in range(nRows):
            iniY = minBY + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                win = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
                kps = det[self.ini_th].detect(win)
                if len(kps) == 0:
                    kps = det[self.min_th].detect(win)
                for kp in kps:
                    out.append((int(kp.pt[0]) + j * wCell, int(kp.pt[1]) + i * hCell, int(kp.response)))
        return out, (minBX, maxBX, minBY, maxBY)

    # ---------------- DistributeOctTree (literal list algorithm; tie-break = creation sequence)
    def distribute(self, keys, minX, maxX, minY, maxY, N):
        nIni = int(math.floor(float(f32(f32(maxX - minX) / f32(maxY - minY))) + 0.5))
        hX = f32(f32(maxX - minX) / f32(nIni))
        seq = [0]

        def mk(ulx, uly, urx, bry):
            seq[0] += 1
            return {"ulx": ulx, "uly": uly, "urx": urx, "bry": bry, "keys": [], "nomore": False, "seq": seq[0]}

        nodes = []
        ini = []
        for i in range(nIni):
            n = mk(int(f32(hX * f32(i))), 0, int(f32(hX * f32(i + 1))), maxY - minY)
            nodes.append(n)
            ini.append(n)
        for k in keys:
            ini[int(f32(f32(k[0]) / hX))]["keys"].append(k)
        kept = []
        for n in nodes:
            if len(n["keys"]) == 1:
                n["nomore"] = True
                kept.append(n)
            elif len(n["keys"]) > 1:
                kept.append(n)
        nodes = kept

        def divide(p):
            halfX = int(math.ceil(f32(f32(p["urx"] - p["ulx"]) / f32(2))))
            halfY = int(math.ceil(f32(f32(p["bry"] - p["uly"]) / f32(2))))
            mx, my = p["ulx"] + halfX, p["uly"] + halfY
            ch = [mk(p["ulx"], p["uly"], mx, my), mk(mx, p["uly"], p["urx"], my),
                  mk(p["ulx"], my, mx, p["bry"]), mk(mx, my, p["urx"], p["bry"])]
            for k in p["keys"]:
                if k[0] < mx:
                    ch[0 if k[1] < my else 2]["keys"].append(k)
                else:
                    ch[1 if k[1] < my else 3]["keys"].append(k)
            for c in ch:
                if len(c["keys"]) == 1:
                    c["nomore"] = True
            return ch

        finish = False
        while not finish:
            prev_size = len(nodes)
            n_expand = 0
            size_nodes = []
            front = []  # nodes pushed to the front during this pass, newest first
            rest = []
            for n in nodes:
                if n["nomore"]:
                    rest.append(n)
                    continue
                for c in divide(n):
                    if c["keys"]:
                        front.insert(0, c)
                        if len(c["keys"]) > 1:
                            n_expand += 1
                            size_nodes.append(c)
            nodes = front + rest
            if len(nodes) >= N or len(nodes) == prev_size:
                finish = True
            elif len(nodes) + n_expand * 3 > N:
                while not finish:
                    prev_size = len(nodes)
                    prev_nodes = sorted(size_nodes, key=lambda c: (len(c["keys"]), c["seq"]))
                    size_nodes = []
                    for p in reversed(prev_nodes):
                        for c in divide(p):
                            if c["keys"]:
                                nodes.insert(0, c)
                                if len(c["keys"]) > 1:
                                    size_nodes.append(c)
                        nodes.remove(p)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev_size:
                        finish = True
        res = []
        for n in nodes:
            best = n["keys"][0]
            for k in n["keys"][1:]:
                if k[2] > best[2]:
                    best = k
            res.append(best)
        return res

    # ---------------- IC_Angle / descriptors
    def ic_angle(self, lvl, x, y):
        im = self.padded[lvl]
        cy, cx = cv_round(y) + EDGE, cv_round(x) + EDGE
        m01 = m10 = 0
        for u in range(-HALF_PATCH, HALF_PATCH + 1):
            m10 += u * int(im[cy, cx + u])
        for v in range(1, HALF_PATCH + 1):
            d = self.umax[v]
            vs = 0
            for u in range(-d, d + 1):
                vp, vm = int(im[cy + v, cx + u]), int(im[cy - v, cx + u])
                vs += vp - vm
                m10 += u * (vp + vm)
            m01 += v * vs
        return f32(self.cv2.fastAtan2(float(f32(m01)), float(f32(m10))))

    def descriptor(self, blurred, x, y, angle_deg):
        factor_pi = f32(math.pi / float(f32(180.0)))
        ang = f32(f32(angle_deg) * factor_pi)
        a, b = f32(math.cos(float(ang))), f32(math.sin(float(ang)))
        cy, cx = cv_round(y), cv_round(x)
        px = self.pattern[:, 0].astype(f32)
        py = self.pattern[:, 1].astype(f32)
        ry = np.rint((px * b).astype(f32) + (py * a).astype(f32)).astype(np.int32)
        rx = np.rint((px * a).astype(f32) - (py * b).astype(f32)).astype(np.int32)
        vals = blurred[cy + ry, cx + rx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)

    # ---------------- operator()
    def extract(self, image):
        cv2 = self.cv2
        self.compute_pyramid(image)
        kps, descs, cands = [], [], []
        for lvl in range(self.nlevels):
            cand, (minBX, maxBX, minBY, maxBY) = self.level_candidates(lvl)
            cands.append(np.array(cand, np.int32).reshape(-1, 3))
            sel = self.distribute(cand, minBX, maxBX, minBY, maxBY, self.nfeat[lvl]) if cand else []
            if not sel:
                continue
            blurred = cv2.GaussianBlur(self.roi[lvl].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            size = f32(int(f32(PATCH) * self.sf[lvl]))
            for (x, y, r) in sel:
                fx, fy = f32(x + minBX), f32(y + minBY)
                ang = self.ic_angle(lvl, fx, fy)
                descs.append(self.descriptor(blurred, fx, fy, ang))
                if lvl != 0:
                    ox, oy = f32(fx * self.sf[lvl]), f32(fy * self.sf[lvl])
                else:
                    ox, oy = fx, fy
                kps.append((ox, oy, size, ang, f32(r), lvl, -1))
        kp = np.array(kps, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                                  ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
        desc = np.array(descs, np.uint8).reshape(-1, 32)
        return kp, desc, cands


def fixtures():
    synth = importlib.import_module("jetracer-orbslam2_b200.synth")
    return [
        # name, image, params(nfeatures, scale, nlevels, ini, min)
        ("cfg1_640x480_seed1000", synth.textured_frame(640, 480, 1000), (1000, 1.2, 8, 20, 7)),
        ("tex_320x240_seed7", synth.textured_frame(320, 240, 7), (500, 1.2, 8, 20, 7)),
        ("lowc_320x240_seed11", synth.low_contrast_frame(320, 240, 11), (500, 1.2, 8, 20, 7)),
        ("sparse_320x240_seed5", synth.sparse_frame(320, 240, 5), (500, 1.2, 8, 20, 7)),
        ("checker_320x240", synth.checkerboard_frame(320, 240), (300, 1.2, 8, 20, 7)),
        ("flat_320x240", synth.flat_frame(320, 240), (300, 1.2, 8, 20, 7)),
        ("wide_424x240_seed3", synth.textured_frame(424, 240, 3), (600, 1.2, 6, 20, 7)),
        ("s15_400x300_seed9", synth.textured_frame(400, 300, 9), (400, 1.5, 4, 20, 7)),
        # the reference's live shape (Context.h:16-17, defines.h:2): 848x480, ONE level, one keypoint per 32x32 cell's worth
        ("refshape_848x480_1level_seed2100", synth.textured_frame(848, 480, 2100), (405, 1.2, 1, 20, 7)),
    ]


def main():
    only = set(sys.argv[1:])  # optional fixture names: regenerate just those
    for name, img, (nf, sc, nl, it, mt) in fixtures():
        if only and name not in only:
            continue
        ex = PyOrbExtractor(nf, sc, nl, it, mt)
        kp, desc, cands = ex.extract(img)
        pyr_sha = [hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in ex.padded]
        out = {
            "image": img, "params": np.array([nf, sc, nl, it, mt], np.float64),
            "kp": kp, "desc": desc, "pyr_sha256": np.array(pyr_sha),
            "level_w": np.array([r.shape[1] for r in ex.roi], np.int32),
            "level_h": np.array([r.shape[0] for r in ex.roi], np.int32),
            "nfeat": np.array(ex.nfeat, np.int32), "umax": np.array(ex.umax, np.int32),
            "pyr_last_padded": ex.padded[-1],
        }
        for l, c in enumerate(cands):
            out[f"cand{l}"] = c.astype(np.int16)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(kp)} keypoints, cands/level {[len(c) for c in cands]}, {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()

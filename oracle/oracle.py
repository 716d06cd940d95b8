"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product (jetracer-orbslam2_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
CAND_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("response", "<i4")])


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class Intrinsics(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ppx", C.c_float), ("ppy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("model", C.c_int32), ("coeffs", C.c_float * 5)]


class Extrinsics(C.Structure):
    _fields_ = [("rotation", C.c_float * 9), ("translation", C.c_float * 3)]


def make_intrinsics(width, height, ppx, ppy, fx, fy, model=0, coeffs=(0, 0, 0, 0, 0)) -> Intrinsics:
    return Intrinsics(width, height, ppx, ppy, fx, fy, model, (C.c_float * 5)(*coeffs))


def make_extrinsics(rotation=(1, 0, 0, 0, 1, 0, 0, 0, 1), translation=(0, 0, 0)) -> Extrinsics:
    return Extrinsics((C.c_float * 9)(*rotation), (C.c_float * 3)(*translation))


def build(native: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only) and return the .so path."""
    target = "native" if native else "all"
    subprocess.run(["make", "-C", _HERE, target], check=True, capture_output=True)
    return os.path.join(_HERE, "liborb_oracle_native.so" if native else "liborb_oracle.so")


_lib = None


def lib(native: bool = False):
    global _lib
    if _lib is not None and not native:
        return _lib
    path = os.path.join(_HERE, "liborb_oracle_native.so" if native else "liborb_oracle.so")
    if not os.path.exists(path):
        path = build(native)
    L = C.CDLL(path)
    u8p, i32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float)
    vp = C.c_void_p
    L.orbo_resize_linear_u8.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, C.c_int, C.c_size_t]
    L.orbo_resize_linear_u8.restype = None
    L.orbo_border_reflect101.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, C.c_int]
    L.orbo_border_reflect101.restype = None
    L.orbo_fast9_window.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, vp, C.c_int]
    L.orbo_fast9_window.restype = C.c_int
    L.orbo_fast_score_map.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t]
    L.orbo_fast_score_map.restype = None
    L.orbo_gaussian_blur7.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t]
    L.orbo_gaussian_blur7.restype = None
    L.orbo_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.orbo_fast_atan2.restype = C.c_float
    L.orbo_match_knn.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, vp, vp, vp]
    L.orbo_match_knn.restype = C.c_int
    L.orbo_match_windowed.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_int, vp, vp]
    L.orbo_match_windowed.restype = C.c_int
    L.orbo_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.c_int, C.c_int]
    L.orbo_create.restype = C.c_int
    L.orbo_destroy.argtypes = [vp]
    L.orbo_destroy.restype = None
    L.orbo_nlevels.argtypes = [vp]
    L.orbo_nlevels.restype = C.c_int
    L.orbo_get_geometry.argtypes = [vp, vp, vp, vp, vp, vp]
    L.orbo_get_geometry.restype = None
    L.orbo_get_umax.argtypes = [vp, vp]
    L.orbo_get_umax.restype = None
    L.orbo_extract.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_int]
    L.orbo_extract.restype = C.c_int
    L.orbo_compute_pyramid.argtypes = [vp, vp, C.c_size_t]
    L.orbo_compute_pyramid.restype = C.c_int
    L.orbo_level_padded.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
    L.orbo_level_padded.restype = vp
    L.orbo_level_candidates.argtypes = [vp, C.c_int, vp, C.c_int]
    L.orbo_level_candidates.restype = C.c_int
    L.orbo_distribute_octree.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.orbo_distribute_octree.restype = C.c_int
    L.orbo_level_blurred.argtypes = [vp, C.c_int, vp]
    L.orbo_level_blurred.restype = C.c_int
    L.orbo_ic_angle.argtypes = [vp, C.c_int, C.c_float, C.c_float]
    L.orbo_ic_angle.restype = C.c_float
    L.orbo_descriptor.argtypes = [vp, C.c_size_t, C.c_float, C.c_float, C.c_float, vp]
    L.orbo_descriptor.restype = None
    L.orbo_extract_many.argtypes = [C.POINTER(Params), vp, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                    C.c_int, C.c_int, vp]
    L.orbo_extract_many.restype = C.c_long
    L.orbo_match_many.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_int, vp, vp, vp]
    L.orbo_match_many.restype = C.c_long
    L.orbo_align_depth_to_other.argtypes = [vp, C.c_float, C.POINTER(Intrinsics), C.POINTER(Intrinsics),
                                            C.POINTER(Extrinsics), vp]
    L.orbo_align_depth_to_other.restype = None
    L.orbo_keypoint_pixel_to_point.argtypes = [vp, C.POINTER(Intrinsics), vp, vp, C.c_int, vp, vp, vp]
    L.orbo_keypoint_pixel_to_point.restype = C.c_int
    L.orbo_reproject_points.argtypes = [vp, C.c_int, vp, C.POINTER(Intrinsics), vp]
    L.orbo_reproject_points.restype = None
    L.orbo_compact_pairs.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.orbo_compact_pairs.restype = C.c_int
    L.orbo_search_by_projection.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, C.c_float, C.c_int, C.c_int, vp, vp]
    L.orbo_search_by_projection.restype = C.c_int
    L.orbo_compute_stereo_matches.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_float, vp, vp]
    L.orbo_compute_stereo_matches.restype = C.c_int
    L.orbo_rgb_to_grayscale.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, vp, C.c_size_t]
    L.orbo_rgb_to_grayscale.restype = None
    del u8p, i32p, f32p
    if not native:
        _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8c(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


# ---------------------------------------------------------------- primitives
def resize_linear(src, dw: int, dh: int) -> np.ndarray:
    src = _u8c(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orbo_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def border_reflect101(roi, b: int = 19) -> np.ndarray:
    roi = _u8c(roi)
    h, w = roi.shape
    out = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    out[b:b + h, b:b + w] = roi
    lib().orbo_border_reflect101(_p(out), w, h, out.strides[0], b)
    return out


def fast9_window(win, thr: int, nms: bool = True) -> np.ndarray:
    win = _u8c(win)
    h, w = win.shape
    out = np.zeros(max(1, w * h), CAND_DTYPE)
    n = lib().orbo_fast9_window(_p(win), w, h, win.strides[0], thr, int(nms), _p(out), out.size)
    return out[:n].copy()


def fast_score_map(img) -> np.ndarray:
    img = _u8c(img)
    h, w = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().orbo_fast_score_map(_p(img), w, h, img.strides[0], _p(out), w)
    return out


def gaussian_blur7(img) -> np.ndarray:
    img = _u8c(img)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().orbo_gaussian_blur7(_p(img), w, h, img.strides[0], _p(out), w)
    return out


def fast_atan2(y: float, x: float) -> float:
    return float(lib().orbo_fast_atan2(C.c_float(y), C.c_float(x)))


def match_knn(q, t, k: int = 2, ratio: float = 0.7, threads: int = 1, native: bool = False):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.full((q.shape[0], 2), -1, np.int32)
    dist = np.full((q.shape[0], 2), -1, np.int32)
    acc = np.zeros(q.shape[0], np.uint8)
    L = lib(native)
    if threads <= 1:
        L.orbo_match_knn(_p(q), q.shape[0], _p(t), t.shape[0], k, ratio, _p(idx), _p(dist), _p(acc))
    else:
        L.orbo_match_many(_p(q), q.shape[0], _p(t), t.shape[0], k, ratio, threads, _p(idx), _p(dist), _p(acc))
    return idx, dist, acc.astype(bool)


def match_windowed(q, q_xy, t, t_xy, max_px: float, max_hamming: int):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    q_xy = np.ascontiguousarray(q_xy, np.float32).reshape(-1, 2)
    t_xy = np.ascontiguousarray(t_xy, np.float32).reshape(-1, 2)
    idx = np.full(q.shape[0], -1, np.int32)
    dist = np.full(q.shape[0], -1, np.int32)
    n = lib().orbo_match_windowed(_p(q), _p(q_xy), q.shape[0], _p(t), _p(t_xy), t.shape[0], max_px, max_hamming,
                                  _p(idx), _p(dist))
    return idx, dist, int(n)


def distribute_octree(cand: np.ndarray, min_x: int, max_x: int, min_y: int, max_y: int, n: int) -> np.ndarray:
    cand = np.ascontiguousarray(cand, CAND_DTYPE)
    out = np.zeros(max(1, cand.size), np.int32)
    k = lib().orbo_distribute_octree(_p(cand), cand.size, min_x, max_x, min_y, max_y, n, _p(out), out.size)
    if k < 0:
        raise RuntimeError(f"orbo_distribute_octree failed: {k}")
    return out[:k].copy()


# ---------------------------------------------------------------- extractor
class Oracle:
    """Upstream `ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)` on the CPU."""

    def __init__(self, width: int, height: int, nfeatures: int = 1000, scale_factor: float = 1.2,
                 nlevels: int = 8, ini_th_fast: int = 20, min_th_fast: int = 7):
        self.params = Params(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast)
        self.w, self.h = width, height
        self._h = C.c_void_p()
        rc = lib().orbo_create(C.byref(self._h), C.byref(self.params), width, height)
        if rc:
            raise ValueError(f"orbo_create failed: {rc}")
        n = nlevels
        self.lw = np.zeros(n, np.int32)
        self.lh = np.zeros(n, np.int32)
        self.scale = np.zeros(n, np.float32)
        self.inv_scale = np.zeros(n, np.float32)
        self.nfeat = np.zeros(n, np.int32)
        lib().orbo_get_geometry(self._h, _p(self.lw), _p(self.lh), _p(self.scale), _p(self.inv_scale), _p(self.nfeat))
        self.umax = np.zeros(16, np.int32)
        lib().orbo_get_umax(self._h, _p(self.umax))
        self.nlevels = n

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orbo_destroy(self._h)
            self._h = None

    @property
    def max_keypoints(self) -> int:
        return int(self.nfeat.sum()) + 16 * self.nlevels + 64

    def extract(self, img):
        img = _u8c(img)
        assert img.shape == (self.h, self.w)
        cap = self.max_keypoints
        kp = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().orbo_extract(self._h, _p(img), img.strides[0], _p(kp), _p(desc), cap)
        if n < 0:
            raise RuntimeError(f"orbo_extract failed: {n}")
        return kp[:n].copy(), desc[:n].copy()

    def compute_pyramid(self, img):
        img = _u8c(img)
        lib().orbo_compute_pyramid(self._h, _p(img), img.strides[0])

    def level_padded(self, level: int) -> np.ndarray:
        pw, ph, pitch = C.c_int(), C.c_int(), C.c_size_t()
        ptr = lib().orbo_level_padded(self._h, level, C.byref(pw), C.byref(ph), C.byref(pitch))
        buf = (C.c_uint8 * (pitch.value * ph.value)).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(ph.value, pitch.value)[:, :pw.value].copy()

    def level_roi(self, level: int) -> np.ndarray:
        return self.level_padded(level)[19:-19, 19:-19]

    def level_candidates(self, level: int) -> np.ndarray:
        cap = int(self.lw[level]) * int(self.lh[level]) // 4 + 64
        out = np.zeros(cap, CAND_DTYPE)
        n = lib().orbo_level_candidates(self._h, level, _p(out), cap)
        return out[:n].copy()

    def level_blurred(self, level: int) -> np.ndarray:
        out = np.empty((int(self.lh[level]), int(self.lw[level])), np.uint8)
        lib().orbo_level_blurred(self._h, level, _p(out))
        return out

    def ic_angle(self, level: int, x: float, y: float) -> float:
        return float(lib().orbo_ic_angle(self._h, level, C.c_float(x), C.c_float(y)))


def descriptor(blurred, x: float, y: float, angle_deg: float) -> np.ndarray:
    blurred = _u8c(blurred)
    out = np.zeros(32, np.uint8)
    lib().orbo_descriptor(_p(blurred), blurred.strides[0], C.c_float(x), C.c_float(y), C.c_float(angle_deg), _p(out))
    return out


def extract_many(frames: np.ndarray, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                 threads: int = 1, native: bool = False):
    """CPU baseline driver: frames [n,h,w] u8 -> (total keypoints, counts[n])."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    p = Params(nfeatures, scale_factor, nlevels, ini_th, min_th)
    counts = np.zeros(n, np.int32)
    tot = lib(native).orbo_extract_many(C.byref(p), _p(frames), w, h, frames.strides[1], frames.strides[0], n,
                                        threads, _p(counts))
    if tot < 0:
        raise RuntimeError(f"orbo_extract_many failed: {tot}")
    return int(tot), counts


# ---------------------------------------------------------------- RGB-D association (rgbd_oracle.c)
def align_depth_to_other(depth: np.ndarray, depth_scale: float, di: Intrinsics, oi: Intrinsics, ex: Extrinsics):
    depth = np.ascontiguousarray(depth, np.uint16)
    assert depth.shape == (di.height, di.width)
    out = np.empty((oi.height, oi.width), np.uint32)
    lib().orbo_align_depth_to_other(_p(depth), depth_scale, C.byref(di), C.byref(oi), C.byref(ex), _p(out))
    return out


def keypoint_pixel_to_point(aligned: np.ndarray, oi: Intrinsics, kp: np.ndarray, desc: np.ndarray):
    aligned = np.ascontiguousarray(aligned, np.uint32)
    kp = np.ascontiguousarray(kp, KEYPOINT_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = kp.shape[0]
    kp_out = np.zeros(max(n, 1), KEYPOINT_DTYPE)
    desc_out = np.zeros((max(n, 1), 32), np.uint8)
    pts = np.zeros((max(n, 1), 3), np.float64)
    m = lib().orbo_keypoint_pixel_to_point(_p(aligned), C.byref(oi), _p(kp), _p(desc), n, _p(kp_out), _p(desc_out), _p(pts))
    return kp_out[:m].copy(), desc_out[:m].copy(), pts[:m].copy()


def reproject_points(points: np.ndarray, T, intrin: Intrinsics) -> np.ndarray:
    points = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    out = np.zeros((max(points.shape[0], 1), 2), np.float32)
    Tc = None if T is None else np.ascontiguousarray(np.asarray(T, np.float64).reshape(4, 4).T)  # -> column-major
    lib().orbo_reproject_points(_p(points), points.shape[0], _p(Tc) if Tc is not None else None, C.byref(intrin), _p(out))
    return out[:points.shape[0]].copy()


def compact_pairs(idx: np.ndarray, q_points: np.ndarray, t_points: np.ndarray, t_xy: np.ndarray):
    idx = np.ascontiguousarray(idx, np.int32)
    q_points = np.ascontiguousarray(q_points, np.float64).reshape(-1, 3)
    t_points = np.ascontiguousarray(t_points, np.float64).reshape(-1, 3)
    t_xy = np.ascontiguousarray(t_xy, np.float32).reshape(-1, 2)
    n = idx.shape[0]
    prev = np.zeros((max(n, 1), 3), np.float64)
    curr = np.zeros((max(n, 1), 3), np.float64)
    xs = np.zeros(max(n, 1), np.uint16)
    ys = np.zeros(max(n, 1), np.uint16)
    m = lib().orbo_compact_pairs(_p(idx), n, _p(q_points), _p(t_points), _p(t_xy), 2, _p(prev), _p(curr), _p(xs), _p(ys))
    return prev[:m].copy(), curr[:m].copy(), xs[:m].copy(), ys[:m].copy()


def rgb_to_grayscale(rgb: np.ndarray) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    out = np.empty((h, w), np.uint8)
    lib().orbo_rgb_to_grayscale(_p(rgb), rgb.strides[0], w, h, _p(out), w)
    return out


def search_by_projection(q_desc, q_uv, q_kp, t_desc, t_kp, scale_factors, th: float, th_high: int = 100,
                         check_orientation: bool = True):
    q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
    t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
    q_uv = np.ascontiguousarray(q_uv, np.float32).reshape(-1, 2)
    q_kp = np.ascontiguousarray(q_kp, KEYPOINT_DTYPE)
    t_kp = np.ascontiguousarray(t_kp, KEYPOINT_DTYPE)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    nq = q_desc.shape[0]
    idx = np.full(max(nq, 1), -1, np.int32)
    dist = np.full(max(nq, 1), -1, np.int32)
    n = lib().orbo_search_by_projection(_p(q_desc), _p(q_uv), _p(q_kp), nq, _p(t_desc), _p(t_kp), t_desc.shape[0], _p(sf),
                                        sf.shape[0], th, th_high, int(check_orientation), _p(idx), _p(dist))
    return idx[:nq].copy(), dist[:nq].copy(), int(n)


def compute_stereo_matches(left: "Oracle", right: "Oracle", kl, dl, kr, dr, bf: float, fx: float):
    """Frame::ComputeStereoMatches on two Oracle objects whose pyramids hold the rectified pair."""
    kl = np.ascontiguousarray(kl, KEYPOINT_DTYPE); kr = np.ascontiguousarray(kr, KEYPOINT_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8).reshape(-1, 32); dr = np.ascontiguousarray(dr, np.uint8).reshape(-1, 32)
    ur = np.full(max(len(kl), 1), -1, np.float32)
    z = np.full(max(len(kl), 1), -1, np.float32)
    n = lib().orbo_compute_stereo_matches(left._h, right._h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), bf, fx, _p(ur), _p(z))
    return ur[:len(kl)].copy(), z[:len(kl)].copy(), int(n)

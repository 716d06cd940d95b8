"""ctypes binding of liborbb200.so (include/orbb200.h) -- the host-side mirror, in Python, of the
reference's ORB stage interface.

Names follow the reference / upstream operator surface so parity tests read like its own would:
  * ``ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)`` with
    ``__call__(image, mask=None) -> (keypoints, descriptors)`` and the ``GetLevels`` / ``GetScaleFactors``
    getters (upstream ORB-SLAM2 surface named by reference src_trash1/orb_extractor.cpp:6-8);
  * the stage functions of namespace Jetracer (reference src/cuda/{pyramid,fast,nms,orb,post_processing}.cuh):
    ``pyramid_create_levels``, ``detect``, ``gaussian_blur``, ``compute_fast_angle_and_orb``, ``match_keypoints``.

There is NO CPU fallback: if the shared library is missing or no B200 is visible, construction raises.
PyTorch is only plumbing here (device buffers / streams for the device-resident entry points).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborbb200.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)

EXPORTS = [
    "orbb_create", "orbb_destroy", "orbb_strerror", "orbb_last_cuda_error", "orbb_get_levels",
    "orbb_get_scale_factors", "orbb_get_features_per_level", "orbb_max_keypoints_per_frame",
    "orbb_get_launch_count", "orbb_get_level",
    "orbb_extract_batch_device", "orbb_extract_batch_host", "orbb_extract_batch_host_async", "orbb_wait",
    "orbb_stage_upload", "orbb_pyramid_create_levels",
    "orbb_detect", "orbb_detect_fast", "orbb_detect_distribute", "orbb_gaussian_blur", "orbb_compute_angle_and_orb",
    "orbb_compute_fast_angle", "orbb_calc_orb", "orbb_detect_export", "orbb_match_knn", "orbb_match_knn_batch",
    "orbb_match_knn_segmented", "orbb_match_windowed", "orbb_debug_get_padded", "orbb_debug_get_blurred", "orbb_debug_get_scores",
    "orbb_debug_popc_rate", "orbb_debug_imma_rate", "orbb_debug_matcher_kind", "orbb_debug_poison", "orbb_debug_get_candidates", "orbb_debug_get_selected", "orbb_debug_distribute",
    "orbb_align_depth_to_other", "orbb_keypoint_pixel_to_point", "orbb_reproject_points", "orbb_match_windowed_batch",
    "orbb_rgb_to_grayscale", "orbb_match_projection_batch", "orbb_compute_stereo_matches",
    "orbb_slam_frame_bson_size", "orbb_slam_frame_to_bson",
    "orbb_rgbd_stage_create", "orbb_rgbd_stage_destroy", "orbb_rgbd_stage_reset", "orbb_rgbd_stage_handle",
    "orbb_rgbd_stage_submit", "orbb_rgbd_stage_wait",
    "orbb_preview_create", "orbb_preview_destroy", "orbb_preview_last_error", "orbb_preview_encode_host",
    "orbb_preview_debug_planes",
]


class OrbbError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class Level(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32), ("roi_offset", C.c_int32),
                ("padded", C.c_void_p), ("blurred", C.c_void_p), ("scale", C.c_float), ("inv_scale", C.c_float),
                ("nfeatures", C.c_int32)]


class Intrinsics(C.Structure):
    """== rs2_intrinsics (librealsense2 rs_types.h)"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ppx", C.c_float), ("ppy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("model", C.c_int32), ("coeffs", C.c_float * 5)]


class Extrinsics(C.Structure):
    """== rs2_extrinsics: column-major rotation, translation"""
    _fields_ = [("rotation", C.c_float * 9), ("translation", C.c_float * 3)]


class RgbdConfig(C.Structure):
    _fields_ = [("orb", Params), ("max_batch", C.c_int32), ("depth_intrin", Intrinsics), ("image_intrin", Intrinsics),
                ("depth_to_image", Extrinsics), ("depth_scale", C.c_float), ("max_pixel_distance", C.c_float),
                ("max_hamming_distance", C.c_int32)]


class SlamFrames(C.Structure):
    """== orbb_slam_frames: host pointers into the stage's pinned result buffers"""
    _fields_ = [("n_frames", C.c_int32), ("max_kp", C.c_int32), ("keypoints_count", C.c_void_p),
                ("valid_keypoints_num", C.c_void_p), ("matched_keypoints_num", C.c_void_p), ("keypoints", C.c_void_p),
                ("descriptors", C.c_void_p), ("points", C.c_void_p), ("previous_matched_points", C.c_void_p),
                ("current_matched_points", C.c_void_p), ("matched_xy", C.c_void_p)]


DISTORTION_NONE, DISTORTION_MODIFIED_BROWN_CONRADY, DISTORTION_INVERSE_BROWN_CONRADY = 0, 1, 2
DISTORTION_FTHETA, DISTORTION_BROWN_CONRADY, DISTORTION_KANNALA_BRANDT4 = 3, 4, 5


def make_intrinsics(width, height, ppx, ppy, fx, fy, model=DISTORTION_NONE, coeffs=(0, 0, 0, 0, 0)) -> Intrinsics:
    return Intrinsics(width, height, ppx, ppy, fx, fy, model, (C.c_float * 5)(*coeffs))


def make_extrinsics(rotation=(1, 0, 0, 0, 1, 0, 0, 0, 1), translation=(0, 0, 0)) -> Extrinsics:
    return Extrinsics((C.c_float * 9)(*rotation), (C.c_float * 3)(*translation))


_lib = None


def load_library():
    """dlopen liborbb200.so and declare every prototype of include/orbb200.h.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrbbError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, sz, f32 = C.c_void_p, C.c_int, C.c_size_t, C.c_float
    L.orbb_create.argtypes = [C.POINTER(vp), C.POINTER(Params), i32, i32, i32, i32]
    L.orbb_destroy.argtypes = [vp]
    L.orbb_strerror.argtypes = [i32]
    L.orbb_strerror.restype = C.c_char_p
    L.orbb_last_cuda_error.argtypes = [vp]
    L.orbb_last_cuda_error.restype = C.c_char_p
    L.orbb_get_levels.argtypes = [vp]
    L.orbb_get_scale_factors.argtypes = [vp, vp, vp, vp, vp]
    L.orbb_get_features_per_level.argtypes = [vp, vp]
    L.orbb_max_keypoints_per_frame.argtypes = [vp]
    L.orbb_get_launch_count.argtypes = [vp]
    L.orbb_get_level.argtypes = [vp, i32, i32, C.POINTER(Level)]
    L.orbb_extract_batch_device.argtypes = [vp, vp, sz, sz, i32, vp, vp, vp, i32, vp]
    L.orbb_extract_batch_host.argtypes = [vp, vp, sz, sz, i32, vp, vp, vp, i32, vp]
    L.orbb_extract_batch_host_async.argtypes = [vp, vp, sz, sz, i32, vp, vp, vp, i32, vp]
    L.orbb_wait.argtypes = [vp, i32]
    L.orbb_stage_upload.argtypes = [vp, vp, sz, sz, i32, vp]
    L.orbb_pyramid_create_levels.argtypes = [vp, vp]
    L.orbb_detect.argtypes = [vp, vp]
    L.orbb_detect_fast.argtypes = [vp, vp]
    L.orbb_detect_distribute.argtypes = [vp, vp]
    L.orbb_gaussian_blur.argtypes = [vp, vp]
    L.orbb_compute_angle_and_orb.argtypes = [vp, vp, vp, vp, i32, vp]
    L.orbb_match_knn.argtypes = [vp, vp, i32, vp, i32, i32, f32, vp, vp, vp, vp, vp]
    L.orbb_match_knn_batch.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, f32, vp, vp, vp, vp, vp]
    L.orbb_compute_fast_angle.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.orbb_calc_orb.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.orbb_detect_export.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.orbb_match_knn_segmented.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp]
    L.orbb_match_windowed.argtypes = [vp, vp, vp, i32, i32, vp, vp, i32, i32, f32, i32, vp, vp, vp, vp]
    L.orbb_debug_popc_rate.argtypes = [vp, C.POINTER(C.c_double)]
    L.orbb_debug_imma_rate.argtypes = [vp, C.POINTER(C.c_double)]
    L.orbb_debug_matcher_kind.argtypes = []
    L.orbb_debug_poison.argtypes = [vp, i32]
    L.orbb_debug_get_padded.argtypes = [vp, i32, i32, vp]
    L.orbb_debug_get_blurred.argtypes = [vp, i32, i32, vp]
    L.orbb_debug_get_scores.argtypes = [vp, i32, i32, vp]
    L.orbb_debug_get_candidates.argtypes = [vp, i32, i32, vp, i32]
    L.orbb_debug_get_selected.argtypes = [vp, i32, i32, vp, i32]
    L.orbb_debug_distribute.argtypes = [vp, i32, vp, i32, i32, vp, i32]
    L.orbb_align_depth_to_other.argtypes = [vp, vp, i32, f32, C.POINTER(Intrinsics), C.POINTER(Intrinsics),
                                            C.POINTER(Extrinsics), vp, vp]
    L.orbb_keypoint_pixel_to_point.argtypes = [vp, vp, C.POINTER(Intrinsics), i32, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    L.orbb_reproject_points.argtypes = [vp, vp, vp, i32, i32, vp, C.POINTER(Intrinsics), vp, vp]
    L.orbb_match_windowed_batch.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32, i32, f32, i32, vp, vp, vp, vp, vp, vp,
                                            vp, vp, vp]
    L.orbb_match_projection_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, i32, i32, vp, vp, vp, vp]
    L.orbb_compute_stereo_matches.argtypes = [vp, vp, vp, vp, i32, i32, f32, f32, vp, vp, vp, vp]
    L.orbb_rgb_to_grayscale.argtypes = [vp, vp, sz, sz, i32, i32, i32, vp, sz, sz, vp]
    L.orbb_slam_frame_bson_size.argtypes = [i32, sz]
    L.orbb_slam_frame_to_bson.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, sz, vp, sz]
    L.orbb_rgbd_stage_create.argtypes = [C.POINTER(vp), C.POINTER(RgbdConfig), i32]
    L.orbb_rgbd_stage_destroy.argtypes = [vp]
    L.orbb_rgbd_stage_reset.argtypes = [vp]
    L.orbb_rgbd_stage_handle.argtypes = [vp]
    L.orbb_rgbd_stage_submit.argtypes = [vp, vp, vp, i32, vp]
    L.orbb_rgbd_stage_wait.argtypes = [vp, i32, C.POINTER(SlamFrames)]
    L.orbb_preview_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    L.orbb_preview_destroy.argtypes = [vp]
    L.orbb_preview_last_error.argtypes = [vp]
    L.orbb_preview_last_error.restype = C.c_char_p
    L.orbb_preview_encode_host.argtypes = [vp, vp, sz, vp, i32, i32, vp, vp, sz, C.POINTER(C.c_size_t), vp]
    L.orbb_preview_debug_planes.argtypes = [vp, vp, sz, vp, i32, i32, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("orbb_strerror", "orbb_last_cuda_error", "orbb_get_launch_count", "orbb_rgbd_stage_handle",
                        "orbb_slam_frame_bson_size", "orbb_slam_frame_to_bson", "orbb_preview_last_error"):
            fn.restype = C.c_int
    L.orbb_slam_frame_bson_size.restype = C.c_size_t
    L.orbb_slam_frame_to_bson.restype = C.c_longlong
    L.orbb_rgbd_stage_handle.restype = vp
    L.orbb_get_launch_count.restype = C.c_longlong
    _lib = L
    return L


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _dev_ptr(t):
    """torch CUDA tensor or raw int address -> void*"""
    if hasattr(t, "data_ptr"):
        if not t.is_cuda or not t.is_contiguous():
            raise OrbbError("device entry points need contiguous CUDA tensors")
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(int(t))


def _stream_ptr(stream):
    if stream is None:
        return C.c_void_p(0)
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


class ORBextractor:
    """B200 ORB extractor for frames of one fixed size; batch capacity ``max_batch``."""

    def __init__(self, nfeatures: int = 1000, scaleFactor: float = 1.2, nlevels: int = 8, iniThFAST: int = 20,
                 minThFAST: int = 7, *, width: int, height: int, max_batch: int = 1, device: int = -1):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.params = Params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self.width, self.height, self.max_batch = width, height, max_batch
        rc = self._lib.orbb_create(C.byref(self._h), C.byref(self.params), width, height, max_batch, device)
        if rc != 0:
            self._h = C.c_void_p()
            raise OrbbError(f"orbb_create: {self._lib.orbb_strerror(rc).decode()} ({rc})")
        self.nlevels = self._lib.orbb_get_levels(self._h)
        self.max_kp = self._lib.orbb_max_keypoints_per_frame(self._h)
        n = self.nlevels
        self._scale = np.zeros(n, np.float32)
        self._inv_scale = np.zeros(n, np.float32)
        self._sigma2 = np.zeros(n, np.float32)
        self._inv_sigma2 = np.zeros(n, np.float32)
        self._check(self._lib.orbb_get_scale_factors(self._h, _np_ptr(self._scale), _np_ptr(self._inv_scale),
                                                     _np_ptr(self._sigma2), _np_ptr(self._inv_sigma2)))
        self.features_per_level = np.zeros(n, np.int32)
        self._check(self._lib.orbb_get_features_per_level(self._h, _np_ptr(self.features_per_level)))

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.orbb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            msg = self._lib.orbb_strerror(rc).decode()
            if rc == -3:
                msg += ": " + self._lib.orbb_last_cuda_error(self._h).decode()
            raise OrbbError(msg)
        return rc

    # -- upstream getters -------------------------------------------------------------------
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(self.params.scale_factor)

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv_scale.copy()

    def GetScaleSigmaSquares(self):
        return self._sigma2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._inv_sigma2.copy()

    def launch_count(self) -> int:
        return int(self._lib.orbb_get_launch_count(self._h))

    def level_info(self, level: int, frame: int = 0) -> Level:
        out = Level()
        self._check(self._lib.orbb_get_level(self._h, frame, level, C.byref(out)))
        return out

    # -- operator() -------------------------------------------------------------------------
    def __call__(self, image: np.ndarray, mask=None):
        """ORBextractor::operator()(image, mask[ignored]) -> (keypoints[KEYPOINT_DTYPE], descriptors[n,32])."""
        kp, desc, counts = self.extract_batch(np.asarray(image)[None])
        n = int(counts[0])
        return kp[0, :n].copy(), desc[0, :n].copy()

    def extract_batch(self, frames: np.ndarray, stream=None):
        """HOST frames [n,h,w] u8 -> (kp [n,max_kp], desc [n,max_kp,32], counts [n]); H2D/D2H inside."""
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim != 3 or frames.shape[1:] != (self.height, self.width):
            raise OrbbError(f"expected [n,{self.height},{self.width}] u8 frames, got {frames.shape}")
        n = frames.shape[0]
        kp = np.zeros((n, self.max_kp), KEYPOINT_DTYPE)
        desc = np.zeros((n, self.max_kp, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        self._check(self._lib.orbb_extract_batch_host(self._h, _np_ptr(frames), frames.strides[1], frames.strides[0], n,
                                                      _np_ptr(kp), _np_ptr(desc), _np_ptr(counts), self.max_kp,
                                                      _stream_ptr(stream)))
        return kp, desc, counts

    def extract_batch_host_into(self, frames_ptr, pitch, stride, n, kp_ptr, desc_ptr, counts_ptr, stream=None):
        """Raw-pointer form of extract_batch (pinned host buffers owned by the caller, e.g. bench.py)."""
        self._check(self._lib.orbb_extract_batch_host(self._h, C.c_void_p(frames_ptr), pitch, stride, n,
                                                      C.c_void_p(kp_ptr), C.c_void_p(desc_ptr),
                                                      C.c_void_p(counts_ptr), self.max_kp, _stream_ptr(stream)))

    def extract_batch_host_async(self, frames_ptr, pitch, stride, n, kp_ptr, desc_ptr, counts_ptr, stream=None) -> int:
        """Submit a batch (pinned host buffers) without waiting; returns a ticket for ``wait``."""
        return self._check(self._lib.orbb_extract_batch_host_async(
            self._h, C.c_void_p(frames_ptr), pitch, stride, n, C.c_void_p(kp_ptr), C.c_void_p(desc_ptr),
            C.c_void_p(counts_ptr), self.max_kp, _stream_ptr(stream)))

    def wait(self, ticket: int):
        self._check(self._lib.orbb_wait(self._h, ticket))

    def extract_batch_device(self, d_frames, n: int, d_kp, d_desc, d_counts, pitch=None, stride=None, stream=None):
        """DEVICE-resident frames (torch CUDA uint8 tensor or address); async on ``stream``."""
        pitch = self.width if pitch is None else pitch
        stride = self.width * self.height if stride is None else stride
        self._check(self._lib.orbb_extract_batch_device(self._h, _dev_ptr(d_frames), pitch, stride, n, _dev_ptr(d_kp),
                                                        _dev_ptr(d_desc), _dev_ptr(d_counts), self.max_kp,
                                                        _stream_ptr(stream)))

    # -- stage interface (reference namespace Jetracer) -------------------------------------
    def stage_upload(self, d_frames, n: int, pitch=None, stride=None, stream=None):
        pitch = self.width if pitch is None else pitch
        stride = self.width * self.height if stride is None else stride
        self._check(self._lib.orbb_stage_upload(self._h, _dev_ptr(d_frames), pitch, stride, n, _stream_ptr(stream)))

    def pyramid_create_levels(self, stream=None):
        self._check(self._lib.orbb_pyramid_create_levels(self._h, _stream_ptr(stream)))

    def detect(self, stream=None):
        self._check(self._lib.orbb_detect(self._h, _stream_ptr(stream)))

    def detect_fast(self, stream=None):
        self._check(self._lib.orbb_detect_fast(self._h, _stream_ptr(stream)))

    def detect_distribute(self, stream=None):
        self._check(self._lib.orbb_detect_distribute(self._h, _stream_ptr(stream)))

    def gaussian_blur(self, stream=None):
        self._check(self._lib.orbb_gaussian_blur(self._h, _stream_ptr(stream)))

    def compute_fast_angle_and_orb(self, d_kp, d_desc, d_counts, stream=None):
        self._check(self._lib.orbb_compute_angle_and_orb(self._h, _dev_ptr(d_kp), _dev_ptr(d_desc), _dev_ptr(d_counts),
                                                         self.max_kp, _stream_ptr(stream)))

    def compute_fast_angle(self, d_angle, d_pos, image_ptr: int, pitch: int, width: int, height: int, n: int, stream=None):
        """Jetracer::compute_fast_angle(d_keypoints_angle, d_keypoints_pos, image, pitch, w, h, n, stream): IC_Angle in
        degrees for n float2 positions of one device image (``image_ptr`` = raw device address, e.g. a level ROI)."""
        self._check(self._lib.orbb_compute_fast_angle(self._h, _dev_ptr(d_angle), _dev_ptr(d_pos), C.c_void_p(int(image_ptr)),
                                                      pitch, width, height, n, _stream_ptr(stream)))

    def calc_orb(self, d_angle, d_pos, d_desc, blurred_ptr: int, pitch: int, width: int, height: int, n: int, stream=None):
        """Jetracer::calc_orb: 256-bit steered BRIEF for n float2 positions on one smoothed device image."""
        self._check(self._lib.orbb_calc_orb(self._h, _dev_ptr(d_angle), _dev_ptr(d_pos), _dev_ptr(d_desc),
                                            C.c_void_p(int(blurred_ptr)), pitch, width, height, n, _stream_ptr(stream)))

    def detect_export(self, d_pos=None, d_score=None, d_level=None, d_level_counts=None, d_counts=None, max_kp=None,
                      stream=None):
        """SoA outputs of Jetracer::detect (pos / score / level per selected keypoint) for the resident batch."""
        opt = lambda t: _dev_ptr(t) if t is not None else C.c_void_p(0)  # noqa: E731
        self._check(self._lib.orbb_detect_export(self._h, opt(d_pos), opt(d_score), opt(d_level), opt(d_level_counts),
                                                 opt(d_counts), self.max_kp if max_kp is None else max_kp,
                                                 _stream_ptr(stream)))

    def match_keypoints(self, d_query, nq: int, d_train, nt: int, d_idx, d_dist, d_accept=None, d_naccept=None,
                        k: int = 2, ratio: float = 0.7, stream=None):
        """Jetracer::match_keypoints slot: brute-force Hamming k-NN + ratio test on device descriptors."""
        self._check(self._lib.orbb_match_knn(self._h, _dev_ptr(d_query), nq, _dev_ptr(d_train), nt, k, ratio,
                                             _dev_ptr(d_idx), _dev_ptr(d_dist),
                                             _dev_ptr(d_accept) if d_accept is not None else C.c_void_p(0),
                                             _dev_ptr(d_naccept) if d_naccept is not None else C.c_void_p(0),
                                             _stream_ptr(stream)))

    def match_keypoints_batch(self, d_query, d_q_counts, n_frames: int, d_train, nt: int, d_idx, d_dist, d_accept=None,
                              d_naccept=None, k: int = 1, ratio: float = 0.7, max_kp=None, stream=None):
        """Every frame of an extraction output ([n][max_kp][32] + device counts) against one train set (cfg 5)."""
        opt = lambda t: _dev_ptr(t) if t is not None else C.c_void_p(0)  # noqa: E731
        self._check(self._lib.orbb_match_knn_batch(self._h, _dev_ptr(d_query), _dev_ptr(d_q_counts), n_frames,
                                                   self.max_kp if max_kp is None else max_kp, _dev_ptr(d_train), nt, k,
                                                   ratio, _dev_ptr(d_idx), _dev_ptr(d_dist), opt(d_accept),
                                                   opt(d_naccept), _stream_ptr(stream)))

    def match_keypoints_segmented(self, d_query, d_q_off, d_train, d_t_off, nseg: int, nq_total: int, max_q_per_seg: int,
                                  max_t_per_seg: int, d_idx, d_dist, d_accept=None, k: int = 2, ratio: float = 0.7,
                                  stream=None):
        """nq_total / max_q_per_seg / max_t_per_seg are HOST-side sizes (the call never reads the offsets back)."""
        self._check(self._lib.orbb_match_knn_segmented(self._h, _dev_ptr(d_query), _dev_ptr(d_q_off), _dev_ptr(d_train),
                                                       _dev_ptr(d_t_off), nseg, nq_total, max_q_per_seg, max_t_per_seg,
                                                       k, ratio,
                                                       _dev_ptr(d_idx), _dev_ptr(d_dist),
                                                       _dev_ptr(d_accept) if d_accept is not None else C.c_void_p(0),
                                                       _stream_ptr(stream)))

    def match_keypoints_windowed(self, d_query, d_query_xy, q_xy_stride: int, nq: int, d_train, d_train_xy,
                                 t_xy_stride: int, nt: int, max_pixel_distance: float, max_hamming_distance: int,
                                 d_idx, d_dist, d_nmatched=None, stream=None):
        """The reference's match_keypoints(current, previous, max_pixel_distance, max_hamming_distance, ...) gate:
        position window first, then best Hamming distance below the cutoff."""
        self._check(self._lib.orbb_match_windowed(
            self._h, _dev_ptr(d_query), _dev_ptr(d_query_xy), q_xy_stride, nq, _dev_ptr(d_train), _dev_ptr(d_train_xy),
            t_xy_stride, nt, max_pixel_distance, max_hamming_distance, _dev_ptr(d_idx), _dev_ptr(d_dist),
            _dev_ptr(d_nmatched) if d_nmatched is not None else C.c_void_p(0), _stream_ptr(stream)))

    # -- RGB-D association (reference src/cuda/cuda-align.cuh, post_processing.cuh) -----------
    def rgb_to_grayscale(self, d_rgb, n: int, d_gray, width=None, height=None, stream=None):
        """Jetracer::rgb_to_grayscale on [n,h,w,3] interleaved RGB8 -> [n,h,w] gray (contiguous device tensors)."""
        w = self.width if width is None else width
        hh = self.height if height is None else height
        self._check(self._lib.orbb_rgb_to_grayscale(self._h, _dev_ptr(d_rgb), 3 * w, 3 * w * hh, w, hh, n, _dev_ptr(d_gray),
                                                    w, w * hh, _stream_ptr(stream)))

    def align_depth_to_other(self, d_depth, n: int, depth_scale: float, depth_intrin: Intrinsics,
                             other_intrin: Intrinsics, depth_to_other: Extrinsics, d_aligned_out, stream=None):
        self._check(self._lib.orbb_align_depth_to_other(self._h, _dev_ptr(d_depth), n, depth_scale, C.byref(depth_intrin),
                                                        C.byref(other_intrin), C.byref(depth_to_other),
                                                        _dev_ptr(d_aligned_out), _stream_ptr(stream)))

    def keypoint_pixel_to_point(self, d_aligned, other_intrin: Intrinsics, n: int, d_kp_in, d_desc_in, d_counts_in,
                                d_kp_out, d_desc_out, d_points, d_valid_counts, stream=None):
        self._check(self._lib.orbb_keypoint_pixel_to_point(
            self._h, _dev_ptr(d_aligned), C.byref(other_intrin), n, _dev_ptr(d_kp_in), _dev_ptr(d_desc_in),
            _dev_ptr(d_counts_in), self.max_kp, _dev_ptr(d_kp_out), _dev_ptr(d_desc_out), _dev_ptr(d_points),
            _dev_ptr(d_valid_counts), _stream_ptr(stream)))

    def reproject_points(self, d_points, d_counts, n: int, d_T, intrin: Intrinsics, d_pos_out, stream=None):
        self._check(self._lib.orbb_reproject_points(
            self._h, _dev_ptr(d_points), _dev_ptr(d_counts), n, self.max_kp,
            _dev_ptr(d_T) if d_T is not None else C.c_void_p(0), C.byref(intrin), _dev_ptr(d_pos_out),
            _stream_ptr(stream)))

    def match_keypoints_windowed_batch(self, d_query, d_query_xy, d_q_counts, d_train, d_train_xy, t_xy_stride: int,
                                       d_t_counts, n: int, max_pixel_distance: float, max_hamming_distance: int,
                                       d_idx, d_dist, d_query_points=None, d_train_points=None, d_prev_matched=None,
                                       d_curr_matched=None, d_xy_u16=None, d_nmatched=None, stream=None):
        opt = lambda t: _dev_ptr(t) if t is not None else C.c_void_p(0)  # noqa: E731
        self._check(self._lib.orbb_match_windowed_batch(
            self._h, _dev_ptr(d_query), _dev_ptr(d_query_xy), _dev_ptr(d_q_counts), _dev_ptr(d_train),
            _dev_ptr(d_train_xy), t_xy_stride, _dev_ptr(d_t_counts), n, self.max_kp, max_pixel_distance,
            max_hamming_distance, _dev_ptr(d_idx), _dev_ptr(d_dist), opt(d_query_points), opt(d_train_points),
            opt(d_prev_matched), opt(d_curr_matched), opt(d_xy_u16), opt(d_nmatched), _stream_ptr(stream)))

    def match_keypoints_projection_batch(self, d_query_desc, d_query_uv, d_query_kp, d_q_counts, d_train_desc,
                                         d_train_kp, d_t_counts, n: int, th: float, th_high: int, check_orientation: bool,
                                         d_idx, d_dist, d_nmatched=None, stream=None):
        """ORB-SLAM2 SearchByProjection gates (radius th * scale^octave, octave band, <= TH_HIGH, rotation bins)."""
        self._check(self._lib.orbb_match_projection_batch(
            self._h, _dev_ptr(d_query_desc), _dev_ptr(d_query_uv), _dev_ptr(d_query_kp), _dev_ptr(d_q_counts),
            _dev_ptr(d_train_desc), _dev_ptr(d_train_kp), _dev_ptr(d_t_counts), n, self.max_kp, th, th_high,
            int(check_orientation), _dev_ptr(d_idx), _dev_ptr(d_dist),
            _dev_ptr(d_nmatched) if d_nmatched is not None else C.c_void_p(0), _stream_ptr(stream)))

    def compute_stereo_matches(self, d_kp, d_desc, d_counts, n_pairs: int, bf: float, fx: float, d_uright, d_depth,
                               d_nstereo=None, stream=None):
        """ORB-SLAM2 Frame::ComputeStereoMatches on the (2p, 2p+1) frames of the batch extracted last."""
        self._check(self._lib.orbb_compute_stereo_matches(
            self._h, _dev_ptr(d_kp), _dev_ptr(d_desc), _dev_ptr(d_counts), self.max_kp, n_pairs, bf, fx, _dev_ptr(d_uright),
            _dev_ptr(d_depth), _dev_ptr(d_nstereo) if d_nstereo is not None else C.c_void_p(0), _stream_ptr(stream)))

    # -- parity / debug access --------------------------------------------------------------
    def debug_poison(self, value: int = 0xCD):
        """fill the handle's stateless scratch with a byte pattern (results must not depend on it)"""
        self._check(self._lib.orbb_debug_poison(self._h, value))

    def debug_popc_rate(self) -> float:
        """measured POPC lanes / clock / SM (register-only microbenchmark; synchronises)"""
        out = C.c_double(0)
        self._check(self._lib.orbb_debug_popc_rate(self._h, C.byref(out)))
        return float(out.value)

    def debug_matcher_kind(self) -> int:
        """0 = XOR / POPC kernel, 1 = warp-level int8 MMA, 2 = tcgen05 int8 MMA (the default)"""
        return int(self._lib.orbb_debug_matcher_kind())

    def debug_imma_rate(self) -> float:
        """measured int8 tensor-core MMAs (m16n8k32) / clock / SM (register-only microbenchmark; synchronises)"""
        out = C.c_double(0)
        self._check(self._lib.orbb_debug_imma_rate(self._h, C.byref(out)))
        return float(out.value)

    def debug_padded(self, level: int, frame: int = 0) -> np.ndarray:
        li = self.level_info(level)
        out = np.zeros((li.height + 38, li.width + 38), np.uint8)
        self._check(self._lib.orbb_debug_get_padded(self._h, frame, level, _np_ptr(out)))
        return out

    def debug_blurred(self, level: int, frame: int = 0) -> np.ndarray:
        li = self.level_info(level)
        out = np.zeros((li.height, li.width), np.uint8)
        self._check(self._lib.orbb_debug_get_blurred(self._h, frame, level, _np_ptr(out)))
        return out

    def debug_scores(self, level: int, frame: int = 0) -> np.ndarray:
        li = self.level_info(level)
        out = np.zeros((li.height, li.width), np.uint8)
        self._check(self._lib.orbb_debug_get_scores(self._h, frame, level, _np_ptr(out)))
        return out

    def debug_candidates(self, level: int, frame: int = 0) -> np.ndarray:
        li = self.level_info(level)
        cap = li.width * li.height // 4 + 64
        out = np.zeros((cap, 3), np.int32)
        n = self._check(self._lib.orbb_debug_get_candidates(self._h, frame, level, _np_ptr(out), cap))
        return out[:n].copy()

    def debug_selected(self, level: int, frame: int = 0) -> np.ndarray:
        out = np.zeros((self.max_kp, 3), np.int32)
        n = self._check(self._lib.orbb_debug_get_selected(self._h, frame, level, _np_ptr(out), self.max_kp))
        return out[:n].copy()

    def debug_distribute(self, level: int, cand_xyr: np.ndarray, quota: int) -> np.ndarray:
        cand = np.ascontiguousarray(cand_xyr, np.int32).reshape(-1, 3)
        out = np.zeros((self.max_kp, 3), np.int32)
        n = self._check(self._lib.orbb_debug_distribute(self._h, level, _np_ptr(cand), cand.shape[0], quota,
                                                        _np_ptr(out), self.max_kp))
        return out[:n].copy()


class RgbdFrameStage:
    """The SlamGpuPipeline slot (reference src/SlamGpuPipeline/buildStream.cpp:345-660) around the B200 handle:
    ``submit(gray, depth[, T])`` enqueues a batch of consecutive RGB-D frames, ``wait(ticket)`` returns the
    per-frame ``slam_frame_t`` fields as numpy views of the stage's pinned buffers."""

    def __init__(self, orb_params: Params, depth_intrin: Intrinsics, image_intrin: Intrinsics,
                 depth_to_image: Extrinsics, depth_scale: float = 0.001, max_pixel_distance: float = 2.0,
                 max_hamming_distance: int = 64, max_batch: int = 1, device: int = -1):
        self._lib = load_library()
        self.cfg = RgbdConfig(orb_params, max_batch, depth_intrin, image_intrin, depth_to_image, depth_scale,
                              max_pixel_distance, max_hamming_distance)
        self._s = C.c_void_p()
        rc = self._lib.orbb_rgbd_stage_create(C.byref(self._s), C.byref(self.cfg), device)
        if rc != 0:
            self._s = C.c_void_p()
            raise OrbbError(f"orbb_rgbd_stage_create: {self._lib.orbb_strerror(rc).decode()} ({rc})")
        self._h = C.c_void_p(self._lib.orbb_rgbd_stage_handle(self._s))
        self.max_kp = self._lib.orbb_max_keypoints_per_frame(self._h)
        self._keep = {}

    def close(self):
        if getattr(self, "_s", None) and self._s.value:
            self._lib.orbb_rgbd_stage_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise OrbbError(f"{self._lib.orbb_strerror(rc).decode()} ({rc})")
        return rc

    def launch_count(self) -> int:
        return int(self._lib.orbb_get_launch_count(self._h))

    def reset(self):
        self._check(self._lib.orbb_rgbd_stage_reset(self._s))

    def submit(self, gray: np.ndarray, depth: np.ndarray, T=None) -> int:
        """gray [n,h,w] u8, depth [n,dh,dw] u16 (numpy, ideally views of pinned memory), T [n,4,4] float64 row-major."""
        gray = np.ascontiguousarray(gray, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        n = gray.shape[0]
        Tc = None
        if T is not None:
            Tc = np.ascontiguousarray(np.asarray(T, np.float64).reshape(n, 4, 4).transpose(0, 2, 1))  # column-major
        t = self._check(self._lib.orbb_rgbd_stage_submit(self._s, _np_ptr(gray), _np_ptr(depth), n,
                                                         _np_ptr(Tc) if Tc is not None else C.c_void_p(0)))
        self._keep[t & 1] = (gray, depth, Tc)  # inputs must outlive the copies
        return t

    def submit_ptr(self, gray_ptr: int, depth_ptr: int, n: int) -> int:
        return self._check(self._lib.orbb_rgbd_stage_submit(self._s, C.c_void_p(gray_ptr), C.c_void_p(depth_ptr), n,
                                                            C.c_void_p(0)))

    def wait(self, ticket: int) -> dict:
        out = SlamFrames()
        self._check(self._lib.orbb_rgbd_stage_wait(self._s, ticket, C.byref(out)))
        n, mk = out.n_frames, out.max_kp

        cache = self.__dict__.setdefault("_views", {})  # the result arrays are stage-owned pinned buffers (two parities)

        def view(ptr, dtype, shape):
            key = (ptr, np.dtype(dtype).str, shape)
            v = cache.get(key)
            if v is None:
                count = int(np.prod(shape))
                buf = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr)
                v = cache[key] = np.frombuffer(buf, dtype, count).reshape(shape)
            return v

        return dict(n_frames=n, max_kp=mk,
                    keypoints_count=view(out.keypoints_count, np.int32, (n,)),
                    valid_keypoints_num=view(out.valid_keypoints_num, np.int32, (n,)),
                    matched_keypoints_num=view(out.matched_keypoints_num, np.int32, (n,)),
                    keypoints=view(out.keypoints, KEYPOINT_DTYPE, (n, mk)),
                    descriptors=view(out.descriptors, np.uint8, (n, mk, 32)),
                    points=view(out.points, np.float64, (n, mk, 3)),
                    previous_matched_points=view(out.previous_matched_points, np.float64, (n, mk, 3)),
                    current_matched_points=view(out.current_matched_points, np.float64, (n, mk, 3)),
                    matched_xy=view(out.matched_xy, np.uint16, (n, 2, mk)))


class Preview:
    """The JPEG the reference sends with every frame (buildStream.cpp:491-521, 613-621): gray frame, keypoints painted
    into the G plane, nvJPEG quality 90 / 4:2:0."""

    def __init__(self, width: int, height: int, quality: int = 90, device: int = -1):
        self._lib = load_library()
        self._p = C.c_void_p()
        self.width, self.height = width, height
        rc = self._lib.orbb_preview_create(C.byref(self._p), width, height, quality, device)
        if rc != 0:
            self._p = C.c_void_p()
            raise OrbbError(f"orbb_preview_create: {self._lib.orbb_strerror(rc).decode()} ({rc})")
        self._buf = np.empty(width * height * 3 + 4096, np.uint8)  # a JPEG never needs more than the raw planes

    def close(self):
        if getattr(self, "_p", None) and self._p.value:
            self._lib.orbb_preview_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise OrbbError(f"{self._lib.orbb_strerror(rc).decode()}: {self._lib.orbb_preview_last_error(self._p).decode()}")

    def encode(self, d_gray, d_xy=None, xy_stride: int = 28, n_kp: int = 0, d_n_kp=None, gray_pitch=None, stream=None) -> bytes:
        n = C.c_size_t(0)
        self._check(self._lib.orbb_preview_encode_host(
            self._p, _dev_ptr(d_gray), self.width if gray_pitch is None else gray_pitch,
            _dev_ptr(d_xy) if d_xy is not None else C.c_void_p(0), xy_stride, n_kp,
            _dev_ptr(d_n_kp) if d_n_kp is not None else C.c_void_p(0), _np_ptr(self._buf), self._buf.size, C.byref(n),
            _stream_ptr(stream)))
        return self._buf[: n.value].tobytes()

    def debug_planes(self, d_gray, d_xy=None, xy_stride: int = 28, n_kp: int = 0, d_n_kp=None, gray_pitch=None) -> np.ndarray:
        out = np.zeros((3, self.height, self.width), np.uint8)
        self._check(self._lib.orbb_preview_debug_planes(
            self._p, _dev_ptr(d_gray), self.width if gray_pitch is None else gray_pitch,
            _dev_ptr(d_xy) if d_xy is not None else C.c_void_p(0), xy_stride, n_kp,
            _dev_ptr(d_n_kp) if d_n_kp is not None else C.c_void_p(0), _np_ptr(out)))
        return out


def slam_frame_to_bson(ax: int, ay: int, az: int, width: int, height: int, keypoints_x, keypoints_y, image=b"",
                       channels: int = 1) -> bytes:
    """The reference's per-frame WebSocket message (WebSocketCom.cpp:164-184) as bytes."""
    L = load_library()
    kx = np.ascontiguousarray(keypoints_x, np.uint16)
    ky = np.ascontiguousarray(keypoints_y, np.uint16)
    if kx.shape != ky.shape or kx.ndim != 1:
        raise OrbbError("keypoints_x / keypoints_y must be 1-D arrays of the same length")
    img = np.frombuffer(bytes(image), np.uint8)
    n = int(L.orbb_slam_frame_bson_size(kx.shape[0], img.shape[0]))
    out = np.empty(n, np.uint8)
    rc = L.orbb_slam_frame_to_bson(ax, ay, az, width, height, channels, _np_ptr(kx), _np_ptr(ky), kx.shape[0],
                                   _np_ptr(img) if img.shape[0] else C.c_void_p(0), img.shape[0], _np_ptr(out), n)
    if rc != n:
        raise OrbbError(f"orbb_slam_frame_to_bson: {rc}")
    return out.tobytes()


def match_knn_host(ex: ORBextractor, query: np.ndarray, train: np.ndarray, k: int = 2, ratio: float = 0.7):
    """Convenience for tests: host descriptors -> device matcher -> host (idx[nq,2], dist[nq,2], accept[nq])."""
    import torch
    q = torch.from_numpy(np.ascontiguousarray(query, np.uint8).reshape(-1, 32)).cuda()
    t = torch.from_numpy(np.ascontiguousarray(train, np.uint8).reshape(-1, 32)).cuda()
    nq, nt = q.shape[0], t.shape[0]
    idx = torch.full((max(nq, 1), 2), -7, dtype=torch.int32, device="cuda")
    dist = torch.full((max(nq, 1), 2), -7, dtype=torch.int32, device="cuda")
    acc = torch.zeros(max(nq, 1), dtype=torch.uint8, device="cuda")
    nacc = torch.zeros(1, dtype=torch.int32, device="cuda")
    ex.match_keypoints(q, nq, t, nt, idx, dist, acc, nacc, k=k, ratio=ratio,
                       stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    return idx[:nq].cpu().numpy(), dist[:nq].cpu().numpy(), acc[:nq].cpu().numpy().astype(bool), int(nacc.item())

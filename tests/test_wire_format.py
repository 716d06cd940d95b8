"""CPU tier: the per-frame BSON message against (1) the REFERENCE'S OWN Bson class, compiled from
/root/reference/src/WebSocket/bson.cpp into oracle/_ref/libref_bson.so (oracle/ref_bson_shim.cpp repeats the add()
sequence of WebSocketCom.cpp:164-184), and (2) an independent struct-level restatement of the BSON layout.  This is the
one stage where the reference itself pins the oracle."""
import ctypes as C
import importlib
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_bson.so")


@pytest.fixture(scope="module")
def orbb():
    import __graft_entry__ as g
    g.build()
    return importlib.import_module("jetracer-orbslam2_b200.orbb")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src/WebSocket"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_bson.so not built and /root/reference absent")
    L = C.CDLL(REF_SO)
    L.ref_slam_frame_bson.restype = C.c_size_t
    L.ref_slam_frame_bson.argtypes = [C.c_int32] * 5 + [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    return L


def ref_message(L, ax, ay, az, w, h, kx, ky, image):
    kx = np.ascontiguousarray(kx, np.uint16); ky = np.ascontiguousarray(ky, np.uint16)
    img = np.frombuffer(bytes(image), np.uint8)
    out = np.zeros(64 + 4 * len(kx) + len(img) + 256, np.uint8)
    n = L.ref_slam_frame_bson(ax, ay, az, w, h, kx.ctypes.data, ky.ctypes.data, len(kx),
                              img.ctypes.data if len(img) else None, len(img), out.ctypes.data, out.size)
    assert n <= out.size
    return out[:n].tobytes()


def py_message(ax, ay, az, w, h, kx, ky, image, channels=1):
    body = b""
    for key, v in (("ax", ax), ("ay", ay), ("az", az), ("width", w), ("height", h), ("channels", channels)):
        body += b"\x10" + key.encode() + b"\x00" + struct.pack("<i", v)
    for key, data in (("keypoints_x", np.asarray(kx, "<u2").tobytes()), ("keypoints_y", np.asarray(ky, "<u2").tobytes()),
                      ("image", bytes(image))):
        body += b"\x05" + key.encode() + b"\x00" + struct.pack("<I", len(data)) + b"\x80" + data
    return struct.pack("<I", 4 + len(body) + 1) + body + b"\x00"


CASES = [
    (12, -7, 300, 848, 480, 0, 0),
    (0, 0, 0, 640, 480, 1, 10),
    (-180, 179, -90, 848, 480, 405, 31_337),
    (1, 2, 3, 1280, 720, 2000, 250_000),
]


@pytest.mark.parametrize("ax,ay,az,w,h,n,img_len", CASES)
def test_bson_matches_reference_class_and_layout(orbb, ref, ax, ay, az, w, h, n, img_len):
    rng = np.random.default_rng(n + img_len)
    kx = rng.integers(0, w, n).astype(np.uint16); ky = rng.integers(0, h, n).astype(np.uint16)
    image = rng.integers(0, 256, img_len).astype(np.uint8).tobytes()
    got = orbb.slam_frame_to_bson(ax, ay, az, w, h, kx, ky, image)
    assert got == ref_message(ref, ax, ay, az, w, h, kx, ky, image), "differs from the reference's Bson class"
    assert got == py_message(ax, ay, az, w, h, kx, ky, image)
    assert struct.unpack("<I", got[:4])[0] == len(got) and got[-1] == 0


def test_bson_from_stage_rows_and_errors(orbb):
    lib = orbb.load_library()
    xy = np.arange(2 * 7, dtype=np.uint16).reshape(2, 7)  # matched_xy rows of one frame: x row, y row
    msg = orbb.slam_frame_to_bson(1, 2, 3, 848, 480, xy[0, :5], xy[1, :5])
    assert msg == py_message(1, 2, 3, 848, 480, xy[0, :5], xy[1, :5], b"")
    small = np.zeros(8, np.uint8)
    rc = lib.orbb_slam_frame_to_bson(0, 0, 0, 1, 1, 1, None, None, 0, None, 0, small.ctypes.data, small.size)
    assert rc == -5  # ORBB_ERR_CAPACITY: never writes past the caller's buffer
    assert lib.orbb_slam_frame_to_bson(0, 0, 0, 1, 1, 1, None, None, 3, None, 0, small.ctypes.data, small.size) == -1

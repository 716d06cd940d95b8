"""The C-ABI library builds, loads and exports every symbol include/orbb200.h declares.
No compute call is made (there is no GPU in the CPU test tier)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return importlib.import_module("jetracer-orbslam2_b200.orbb")


def _declared():
    txt = open(os.path.join(ROOT, "include", "orbb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orbb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in orbb200.h but not exported"
    assert sorted(built.EXPORTS) == names


def test_struct_layouts(built):
    assert built.KEYPOINT_DTYPE.itemsize == 28  # == cv::KeyPoint
    assert ctypes.sizeof(built.Params) == 20


def test_strerror_and_no_device_is_loud(built):
    lib = built.load_library()
    assert lib.orbb_strerror(0) == b"ok"
    assert b"no CPU path" in lib.orbb_strerror(-2)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(built.OrbbError):
            built.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480)  # must not fall back to a CPU path


def test_cpp_host_mirror_compiles_and_links(built, tmp_path):
    """include/orbb200.hpp (what a reference maintainer includes) compiles as C++17 and links against the library;
    also pins the layouts the stage structs share with librealsense (rs2_intrinsics 48 B, rs2_extrinsics 48 B)."""
    import subprocess
    src = tmp_path / "mirror.cpp"
    src.write_text('#include "orbb200.hpp"\n'
                   'static_assert(sizeof(orbb_intrinsics) == 48 && sizeof(orbb_extrinsics) == 48, "rs2 layout");\n'
                   'static_assert(sizeof(orbb_keypoint) == 28, "cv::KeyPoint layout");\n'
                   'int main(int argc, char **) {\n'
                   '  if (argc > 100) { orbb_rgbd_config c{}; orbb200::RgbdFrameStage s(c); s.reset();\n'
                   '    orbb200::ORBextractor e(1000, 1.2f, 8, 20, 7, 640, 480); (void)e.GetLevels(); }\n'
                   '  return 0; }\n')
    libdir = os.path.dirname(built.LIB_PATH)
    exe = tmp_path / "mirror"
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lorbb200", "-Wl,-rpath," + libdir], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
    assert ctypes.sizeof(built.Intrinsics) == 48 and ctypes.sizeof(built.Extrinsics) == 48

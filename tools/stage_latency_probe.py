#!/usr/bin/env python3
"""Latency of one RGB-D frame through orbb_rgbd_stage_submit + wait (max_batch = 1, the reference's per-wake-up
use), 848x480 gray + depth from pinned memory, median over 200 frames: through the Python wrapper, and from a C++ host
(tools/stage_latency.cpp, compiled here with g++) the way the reference's pipeline thread would call the C ABI.  Consecutive
frames are the same texture shifted by one pixel, so the windowed matcher finds its pairs.
usage (GPU box): python tools/stage_latency_probe.py"""
import importlib, os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
synth = importlib.import_module("jetracer-orbslam2_b200.synth")
for (w, h, nf) in ((848, 480, 1200), (640, 480, 1000)):
    base = synth.textured_frame(w, h, 50)
    gray = torch.from_numpy(np.stack([np.roll(base, i if i < 3 else 1, axis=1) for i in range(4)])).pin_memory()
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    depth = np.clip(1500 + 900 * np.sin(xx / 97.0) * np.cos(yy / 61.0), 1, 65535).astype(np.uint16)
    pd = torch.from_numpy(np.stack([depth] * 4).view(np.int16)).pin_memory()
    di = orbb.make_intrinsics(w, h, w * 0.5 + 3.7, h * 0.5 - 2.2, 0.502 * w, 0.502 * w, 4)
    oi = orbb.make_intrinsics(w, h, w * 0.5 - 5.1, h * 0.5 + 4.3, 0.72 * w, 0.725 * w, 2)
    ex = orbb.make_extrinsics((1, 0, 0, 0, 1, 0, 0, 0, 1), (0.0148, 0.0002, 0.0003))
    stage = orbb.RgbdFrameStage(orbb.Params(nf, 1.2, 8, 20, 7), di, oi, ex, 0.001, 2.0, 64, max_batch=1)
    fb, db = w * h, 2 * w * h
    for i in range(10):
        stage.wait(stage.submit_ptr(gray.data_ptr() + (i % 4) * fb, pd.data_ptr() + (i % 4) * db, 1))
    ts, tsub = [], []
    for i in range(200):
        t = time.perf_counter()
        tk = stage.submit_ptr(gray.data_ptr() + (i % 4) * fb, pd.data_ptr() + (i % 4) * db, 1)
        t1 = time.perf_counter()
        r = stage.wait(tk)
        ts.append(time.perf_counter() - t); tsub.append(t1 - t)
    # ORBB_STAGE_PROF=1 additionally prints the device-side timeline of the phases (events on the stage's streams)
    print(f"{w}x{h} {nf} kp: one RGB-D frame submit+wait median {1e6 * float(np.median(ts)):.1f} us "
          f"(host issue inside submit {1e6 * float(np.median(tsub)):.1f} us), "
          f"valid {int(r['valid_keypoints_num'][0])}, matched {int(r['matched_keypoints_num'][0])}", flush=True)
    stage.close()
    # the same frames from a C++ host
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "_build", "stage_latency")
    lib = os.path.join(root, "jetracer-orbslam2_b200")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"), "-I/usr/local/cuda/include",
                    os.path.join(root, "tools", "stage_latency.cpp"), "-o", exe, "-L" + lib, "-lorbb200",
                    "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + lib], check=True)
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        f.write(gray.numpy().tobytes()); f.write(pd.numpy().tobytes()); f.flush()
        subprocess.run([exe, f.name, str(w), str(h), str(nf), "1000"], check=True)

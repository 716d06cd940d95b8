#!/usr/bin/env python3
"""Throughput / latency sweep over the BASELINE.json configs that bench.py does not time (cfg 2, 3, 4): device-resident
frames/s through orbb_extract_batch_device for several batch sizes, one JSON line per (config, batch).
usage (GPU box): python tools/sweep_configs.py > gpurun_out/sweep.jsonl"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

g.build()
orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
synth = importlib.import_module("jetracer-orbslam2_b200.synth")
CONFIGS = [("cfg1/5 640x480/1000", 640, 480, 1000, [1, 4, 16, 64, 256]),
           ("cfg2 848x480/1200", 848, 480, 1200, [1, 16, 128]),
           ("cfg3 848x800/1000", 848, 800, 1000, [1, 16, 128]),
           ("cfg4 1280x720/2000", 1280, 720, 2000, [1, 4, 16, 64, 256])]
st = torch.cuda.current_stream()
for name, w, h, nf, batches in CONFIGS:
    base = [synth.textured_frame(w, h, 7000 + i) for i in range(8)]
    for B in batches:
        frames = np.stack([np.roll(base[i % 8], (5 * (i // 8), 3 * (i // 8)), axis=(0, 1)) for i in range(B)])
        ex = orbb.ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=B)
        d_in = [torch.from_numpy(frames).cuda(), torch.from_numpy(frames[::-1].copy()).cuda()]
        d_kp = torch.zeros(B * ex.max_kp * 28, dtype=torch.uint8, device="cuda")
        d_desc = torch.zeros(B * ex.max_kp * 32, dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        iters = max(5, min(200, 2048 // B))
        for i in range(3):
            ex.extract_batch_device(d_in[i % 2], B, d_kp, d_desc, d_cnt, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(iters):
            ex.extract_batch_device(d_in[i % 2], B, d_kp, d_desc, d_cnt, stream=st)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(json.dumps({"config": name, "batch": B, "ms_per_batch": round(ms, 4), "frames_per_s": round(B / ms * 1e3, 1),
                          "keypoints_per_frame": float(d_cnt.float().mean().item())}), flush=True)
        ex.close()
        del d_in, d_kp, d_desc
        torch.cuda.empty_cache()

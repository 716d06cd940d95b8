#!/usr/bin/env python3
"""Matcher throughput: nq random 256-bit queries against an nt-row map, k = 1 and k = 2 (Gpairs/s, CUDA events).
usage (GPU box): python tools/matcher_probe.py"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=640, height=480, max_batch=1)
st = torch.cuda.current_stream()
g = torch.Generator(device="cuda").manual_seed(1)
for nq, nt in ((257020, 50000), (100000, 200000), (2000, 2000)):
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
    t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda", generator=g)
    idx = torch.zeros((nq, 2), dtype=torch.int32, device="cuda"); dist = torch.zeros_like(idx)
    for k in (1, 2):
        for _ in range(2):
            ex.match_keypoints(q, nq, t, nt, idx, dist, k=k, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 5 if nq * nt > 1e9 else 200
        e0.record(st)
        for _ in range(it):
            ex.match_keypoints(q, nq, t, nt, idx, dist, k=k, stream=st)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print(f"nq={nq} nt={nt} k={k}: {ms:.4f} ms, {nq * nt / ms / 1e6:.1f} Gpairs/s", flush=True)

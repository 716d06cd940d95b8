"""Lone-frame call latency against the number of pyramid levels (848x480, 1200 keypoints): the slope is what one more
level of the resize -> FAST -> quadtree chain costs a single frame.  usage: python tools/lone_levels_probe.py [iters]"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
synth = importlib.import_module("jetracer-orbslam2_b200.synth")
st = torch.cuda.current_stream()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w, h = 848, 480
for nl in range(1, 9):
    ex = orbb.ORBextractor(1200, 1.2, nl, 20, 7, width=w, height=h, max_batch=1)
    d_in = torch.from_numpy(synth.textured_frame(w, h, 2000)[None]).cuda()
    d_kp = torch.zeros(ex.max_kp * 28, dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros(ex.max_kp * 32, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(5):
        ex.extract_batch_device(d_in, 1, d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lat = []
    for _ in range(iters):
        e0.record(st); ex.extract_batch_device(d_in, 1, d_kp, d_desc, d_cnt, stream=st); e1.record(st)
        torch.cuda.synchronize(); lat.append(e0.elapsed_time(e1))
    print(f"levels={nl}: median {1e3*np.median(lat):.1f} us, best {1e3*min(lat):.1f} us, keypoints {int(d_cnt.item())}", flush=True)
    ex.close()

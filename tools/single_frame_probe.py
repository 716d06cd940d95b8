"""Single-frame call (the reference's operating mode: one frame in flight): a few isolated calls at 848x480 for the
1-level / 405-kp shape and the 8-level / 1200-kp shape.  Run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel durations, or alone for the call latency (CUDA events + synchronise)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
synth = importlib.import_module("jetracer-orbslam2_b200.synth")
st = torch.cuda.current_stream()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w, h = 848, 480
for nf, nl in ((405, 1), (1200, 8)):
    ex = orbb.ORBextractor(nf, 1.2, nl, 20, 7, width=w, height=h, max_batch=1)
    d_in = torch.from_numpy(synth.textured_frame(w, h, 2000)[None]).cuda()
    d_kp = torch.zeros(ex.max_kp * 28, dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros(ex.max_kp * 32, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ex.extract_batch_device(d_in, 1, d_kp, d_desc, d_cnt, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lat = []
    for _ in range(iters):
        e0.record(st); ex.extract_batch_device(d_in, 1, d_kp, d_desc, d_cnt, stream=st); e1.record(st)
        torch.cuda.synchronize(); lat.append(e0.elapsed_time(e1))
    import time
    enq = []
    for _ in range(iters):  # host time spent inside the call (enqueue only), GPU idle at entry
        t0 = time.perf_counter(); ex.extract_batch_device(d_in, 1, d_kp, d_desc, d_cnt, stream=st); enq.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    print(f"{w}x{h} levels={nl} nfeatures={nf}: call latency median {1e3*np.median(lat):.1f} us, best {1e3*min(lat):.1f} us; "
          f"host enqueue time median {1e6*np.median(enq):.1f} us", flush=True)
    ex.close()

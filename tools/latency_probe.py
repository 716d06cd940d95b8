import importlib, os, sys, time, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
synth = importlib.import_module("jetracer-orbslam2_b200.synth")
st = torch.cuda.current_stream()
for (w,h,nf) in ((640,480,1000),(848,480,1200),(1280,720,2000)):
    for B in (1,4):
        frames = np.stack([synth.textured_frame(w,h,7000+i) for i in range(B)])
        ex = orbb.ORBextractor(nf,1.2,8,20,7,width=w,height=h,max_batch=B)
        d_in = torch.from_numpy(frames).cuda()
        d_kp = torch.zeros(B*ex.max_kp*28,dtype=torch.uint8,device="cuda"); d_desc = torch.zeros(B*ex.max_kp*32,dtype=torch.uint8,device="cuda"); d_cnt = torch.zeros(B,dtype=torch.int32,device="cuda")
        for i in range(5):
            ex.extract_batch_device(d_in,B,d_kp,d_desc,d_cnt,stream=st)
        torch.cuda.synchronize()
        ts=[]
        for i in range(200):
            t=time.perf_counter(); ex.extract_batch_device(d_in,B,d_kp,d_desc,d_cnt,stream=st); torch.cuda.synchronize(); ts.append(time.perf_counter()-t)
        print(os.environ.get("ORBB_GRAPH","1"), w,h,B, "median latency us", round(1e6*float(np.median(ts)),1), flush=True)
        ex.close()

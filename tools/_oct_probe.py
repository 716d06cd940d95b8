import importlib, os, sys, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
synth = importlib.import_module("jetracer-orbslam2_b200.synth")
w,h=640,480
ex = orbb.ORBextractor(1000,1.2,8,20,7,width=w,height=h,max_batch=1)
img = synth.textured_frame(w,h,1000)
for i in range(2):
    ex(img)
torch.cuda.synchronize()

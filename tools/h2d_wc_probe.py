#!/usr/bin/env python3
"""Concurrent pinned-host -> device copy bandwidth per rank with ordinary pinned buffers vs write-combined ones
(cudaHostAllocWriteCombined: the CPU only writes such a buffer, the GPU's PCIe reads skip the cache snoops).
One process per GPU under torchrun; copies are barrier-aligned; the slowest rank and the sum are printed.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_wc_probe.py"""
import ctypes, os, sys, glob
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import nvidia.cuda_runtime, os.path as P
rt = ctypes.CDLL(glob.glob(P.join(P.dirname(nvidia.cuda_runtime.__file__), "lib", "libcudart.so*"))[0])
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
NB, ND = 256 * 640 * 480, 16344064


def host(nbytes, flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), nbytes, flags) == 0
    return p


def bar():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()


dev = torch.empty(NB, dtype=torch.uint8, device="cuda"); dout = torch.empty(ND, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, flags in (("default", 0), ("write_combined", 4), ("portable", 1)):
    hin = host(NB, flags); hout = host(ND, 0)
    ctypes.memset(hin, 7, NB)
    for duplex in (False, True):
        best = []
        for rep in range(6):
            bar()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s1)
            for k in range(4):
                rt.cudaMemcpyAsync(dev.data_ptr(), hin, NB, 1, ctypes.c_void_p(s1.cuda_stream))
                if duplex:
                    rt.cudaMemcpyAsync(hout, dout.data_ptr(), ND, 2, ctypes.c_void_p(s2.cuda_stream))
            e1.record(s1)
            torch.cuda.synchronize()
            if rep:
                best.append(4 * NB / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        res[(name, duplex)] = float(np.median(best))
keys = sorted(res)
t = torch.tensor([res[k] for k in keys], device="cuda")
if world > 1:
    mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN); sm = t.clone(); dist.all_reduce(sm)
else:
    mn = sm = t
if rank == 0:
    for k, a, b in zip(keys, mn.tolist(), sm.tolist()):
        print(f"{k[0]:15s} duplex={k[1]!s:5s} H2D GB/s slowest rank {a:6.1f}  aggregate {b:7.1f}", flush=True)
if world > 1:
    dist.destroy_process_group()

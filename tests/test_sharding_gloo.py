"""world_size-2 gloo test of the multi-GPU host logic (frame sharding, count gather, ragged match gather).
Runs on CPU; the NCCL path uses the same code with CUDA tensors."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sh = importlib.import_module("jetracer-orbslam2_b200.sharding")
    lo, hi = sh.shard_range(n_frames, rank, world)
    local_counts = torch.arange(lo, hi, dtype=torch.int32) * 3 + 1  # "keypoints of frame f" = 3f+1
    allc = sh.gather_counts(local_counts, n_frames)
    m = torch.full((1600, 32), 7, dtype=torch.uint8) if rank == 0 else torch.zeros((1600, 32), dtype=torch.uint8)
    sh.broadcast_map(m)
    rows = torch.stack([torch.arange(lo, hi), torch.arange(lo, hi) * 2, torch.full((hi - lo,), rank)], 1).to(torch.int32)
    rows = rows[: (hi - lo) - rank]  # ragged: rank 1 contributes one row less
    gathered, per_rank = sh.gather_ragged_to_rank0(rows)
    # fixed-stride form (bench.py cfg 5): counts and [frames][max_kp] result rows in ONE all_gather_into_tensor
    max_kp, n_pad = 5, (n_frames + world - 1) // world
    lay = sh.GatherLayout(n_pad, max_kp)
    buf = torch.full((lay.total,), -1, dtype=torch.int32)
    cnt, idx, dst = lay.views(buf)
    cnt.zero_()
    for f in range(lo, hi):
        c = f % (max_kp + 1)                       # "keypoints of frame f"
        cnt[f - lo] = c
        for s_ in range(c):
            idx[(f - lo) * max_kp + s_, 0] = 100 * f + s_   # "train index"
            dst[(f - lo) * max_kp + s_, 0] = f + s_         # "distance"
    g = sh.gather_fixed(buf)
    rec = lay.records(g.numpy(), n_frames)
    q.put((rank, allc.tolist(), int(m.sum()), None if gathered is None else gathered.tolist(), per_rank, rec.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 8, 1])  # 7 and 1: uneven shards, padded gather rows
def test_shard_and_gather_world2(n_frames):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_counts = [3 * f + 1 for f in range(n_frames)]
    exp_rec = [[f, s_, 100 * f + s_, f + s_] for f in range(n_frames) for s_ in range(f % 6)]
    for rank, allc, msum, gathered, per_rank, rec in res:
        assert allc == expect_counts
        assert msum == 1600 * 32 * 7  # map broadcast from rank 0
        assert rec == exp_rec         # every rank holds the same gathered records, in global frame order
    sh = importlib.import_module("jetracer-orbslam2_b200.sharding")
    sizes = [sh.shard_range(n_frames, r, world) for r in range(world)]
    exp_rows = []
    for r, (lo, hi) in enumerate(sizes):
        exp_rows += [[f, 2 * f, r] for f in range(lo, hi)][: max((hi - lo) - r, 0)]
    assert res[0][3] == exp_rows and res[1][3] is None
    assert res[0][4] == [max(hi - lo - r, 0) for r, (lo, hi) in enumerate(sizes)]


def test_shard_range_partition():
    sh = importlib.import_module("jetracer-orbslam2_b200.sharding")
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 4, 8):
            spans = [sh.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1

"""Multi-GPU plumbing for the ORB front-end: frames are independent units, so a batch is sharded into
contiguous blocks, one block per rank (one process per GPU), with NO collective on the data path.
torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests) is used only to
  * broadcast the descriptor map once (cfg 5: 50k x 32 B = 1.6 MB),
  * gather per-frame keypoint counts, and
  * gather the ragged per-rank match records {query, train, distance} to rank 0
(SURVEY.md 8e).  The reference has no multi-GPU path to mirror (no NCCL/MPI symbols anywhere)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of frames owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_map(d_map: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(d_map, src=src, group=group)
    return d_map


def gather_counts(local_counts: torch.Tensor, n_frames: int, group=None) -> torch.Tensor:
    """All ranks get the per-frame keypoint counts of the whole batch, in frame order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_counts.clone()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_frames, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(pad, dtype=local_counts.dtype, device=local_counts.device)
    buf[: local_counts.numel()] = local_counts
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


def gather_ragged_to_rank0(local_rows: torch.Tensor, group=None):
    """Ragged gather of [n_i, k] rows (match records) to rank 0: all_gather of the row counts, then one
    padded all_gather.  Returns (rows, per_rank_counts) on rank 0 and (None, per_rank_counts) elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_rows, [int(local_rows.shape[0])]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    ns = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    counts = [int(v.item()) for v in ns]
    pad = max(max(counts), 1)
    buf = torch.zeros((pad,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    buf[: local_rows.shape[0]] = local_rows
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    if rank != 0:
        return None, counts
    return torch.cat([o[:c] for o, c in zip(out, counts)]), counts

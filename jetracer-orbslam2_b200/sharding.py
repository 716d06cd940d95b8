"""Multi-GPU plumbing for the ORB front-end: frames are independent units, so a batch is sharded into
contiguous blocks, one block per rank (one process per GPU), with NO collective on the data path.
torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests) is used only to
  * broadcast the descriptor map once (cfg 5: 50k x 32 B = 1.6 MB),
  * gather per-frame keypoint counts and match results: ONE fixed-stride all_gather_into_tensor of a buffer in which
    the counts ride in front of the [frames][max_kp] result rows (gather_fixed) -- no per-rank .item(), no ragged
    second round, no concatenation,
  * (older, ragged form kept for callers with compacted records: gather_counts / gather_ragged_to_rank0)
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


class GatherLayout:
    """One rank's int32 gather buffer for `n` local frames (every rank uses the same n; pad a short shard with
    zero-count frames): [counts: n | idx: n x max_kp x 2 | dist: n x max_kp x 2].  The extraction writes its
    per-frame counts and the batch matcher its idx/dist rows straight into views of this buffer, so nothing is packed
    before the collective and the counts ride in the same message."""

    def __init__(self, n_frames_local: int, max_kp: int):
        self.n, self.max_kp = int(n_frames_local), int(max_kp)
        self.o_idx = self.n
        self.o_dist = self.o_idx + 2 * self.n * self.max_kp
        self.total = self.o_dist + 2 * self.n * self.max_kp

    def views(self, buf: torch.Tensor):
        """(counts [n], idx [n*max_kp, 2], dist [n*max_kp, 2]) views of a flat int32 buffer of `total` elements"""
        return (buf[: self.n], buf[self.o_idx: self.o_dist].view(-1, 2), buf[self.o_dist: self.total].view(-1, 2))

    def records(self, gathered, n_frames_total: int):
        """HOST side (numpy): [world, total] gathered buffers -> int32 records (global frame, keypoint slot, train
        index, distance) of every matched-against keypoint, in global frame order -- independent of the sharding."""
        import numpy as np
        g = np.asarray(gathered).reshape(-1, self.total)
        out = []
        for r in range(g.shape[0]):
            frame0, hi = shard_range(n_frames_total, r, g.shape[0])
            cnt = g[r, : self.n]
            idx = g[r, self.o_idx: self.o_dist].reshape(self.n, self.max_kp, 2)
            dst = g[r, self.o_dist: self.total].reshape(self.n, self.max_kp, 2)
            for f in range(hi - frame0):
                c = int(min(cnt[f], self.max_kp))
                rec = np.empty((c, 4), np.int32)
                rec[:, 0] = frame0 + f
                rec[:, 1] = np.arange(c)
                rec[:, 2] = idx[f, :c, 0]
                rec[:, 3] = dst[f, :c, 0]
                out.append(rec)
        return np.concatenate(out) if out else np.zeros((0, 4), np.int32)


def gather_fixed(local_buf: torch.Tensor, out: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """One all_gather_into_tensor of equally sized per-rank buffers -> [world, len(local_buf)] on every rank (rank 0
    is the consumer).  `out` may be preallocated (world * numel) so a timed loop allocates nothing."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if out is None:
            return local_buf.reshape(1, -1)
        out.reshape(1, -1).copy_(local_buf.reshape(1, -1))
        return out.reshape(1, -1)
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty(world * local_buf.numel(), dtype=local_buf.dtype, device=local_buf.device)
    dist.all_gather_into_tensor(out, local_buf.reshape(-1), group=group)
    return out.reshape(world, -1)

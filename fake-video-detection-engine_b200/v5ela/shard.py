"""Multi-GPU plumbing: one process per GPU (torchrun), frames sharded by contiguous ranges, records gathered to rank 0.

Frames are independent (halos are intra-frame), so the data path has NO collective; the only exchange is the small
fixed-size record per frame (3144 B) or per video, gathered once per batch with NCCL over NVLink.
Works with the gloo backend on CPU tensors too, which is how the host logic is tested without GPUs.
"""
from __future__ import annotations

from .records import RECORD_BYTES


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of `n` items owned by `rank`; sizes differ by at most one, earlier ranks larger."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_videos(n_videos: int, frames_per_video: int, rank: int, world: int) -> tuple[int, int]:
    """Frame range owned by `rank` when whole videos are kept on one rank (per-video reductions stay local)."""
    v0, v1 = shard_range(n_videos, rank, world)
    return v0 * frames_per_video, v1 * frames_per_video


_recv_cache: dict = {}


def gather_records(local, total: int, dst: int = 0, group=None):
    """Gather per-rank record tensors (n_local, 3144) uint8 to rank `dst` in rank order -> (total, 3144) or None.

    Ranks may hold different counts (shard_range); shorter shards are padded to the longest for the collective. The receive
    block on `dst` is allocated once per (world, longest, device) and reused: with equal shards (every BASELINE.json configuration
    on 1/2/4/8 GPUs) the result is a view of it — no allocation and no concatenation per step; it is overwritten by the next call.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_range(total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in counts)
    send = local
    if local.shape[0] != longest:
        send = torch.zeros((longest, RECORD_BYTES), dtype=torch.uint8, device=local.device)
        send[: local.shape[0]] = local
    bufs = None
    if rank == dst:
        key = (world, longest, str(local.device), id(group))
        block = _recv_cache.get(key)
        if block is None:
            block = _recv_cache[key] = torch.empty((world, longest, RECORD_BYTES), dtype=torch.uint8, device=local.device)
        bufs = list(block.unbind(0))
    dist.gather(send.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    if all(hi - lo == longest for lo, hi in counts):
        return block.view(world * longest, RECORD_BYTES)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, counts)], dim=0)


def analyze_sharded(frames_local, total: int, quality: int = 90, group_size: int = 0, dst: int = 0, pg=None):
    """Analyse this rank's shard and gather to rank `dst`.

    frames_local : this rank's frames (n_local, H, W, 3) uint8 on its GPU — rank r holds shard_range(total, r, world)
                   (or shard_videos(...) when group_size > 0, `total` then counts VIDEOS).
    group_size   : 0 = gather per-frame records; k > 0 = reduce every k consecutive frames (one video) on the device
                   first and gather only the per-video records.
    Returns (records on rank dst | None elsewhere).
    """
    from .batch import analyze_batch, reduce_records

    out = analyze_batch(frames_local, quality=quality)
    recs = out["records"]
    if group_size > 0:
        recs = reduce_records(recs, group_size)
    return gather_records(recs, total, dst=dst, group=pg)

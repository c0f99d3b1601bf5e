#!/usr/bin/env python
"""Multi-GPU parity check for BASELINE.json config 4 (videos x keyframes, per-video feature gather via NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multi_gpu_check.py [--videos 16] [--frames 8] [--height 270] [--width 480]

Video v, frame k = gen_frame(frames*v + k, H, W, seed=v). Whole videos stay on one rank
(v5ela.shard.shard_videos); every rank reduces its videos on the device and rank 0 gathers the per-video records.
Rank 0 recomputes every video with the C oracle (a pool of host processes forked before CUDA is touched, so that the full-size
case — 64 x 32 x 1080p — costs seconds of box time) and requires byte-identical records. Prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fake-video-detection-engine_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _oracle_video(job):
    """One video through the C oracle on the host: its per-video record (bytes) and the records of its first frames."""
    import v5ela
    from oracle import c_oracle

    v, frames, h, w = job
    recs, _ = c_oracle.analyze(v5ela.gen_batch(frames * v, frames, h, w, seed=v), 90)
    return v, v5ela.combine(recs).tobytes(), recs.tobytes() if v == 0 else b""


def main():
    import multiprocessing as mp

    ap0 = argparse.ArgumentParser(add_help=False)
    for name, dflt in (("--videos", 16), ("--frames", 8), ("--height", 270), ("--width", 480)):
        ap0.add_argument(name, type=int, default=dflt)
    a0, _ = ap0.parse_known_args()
    pool = jobs = None
    if int(os.environ.get("RANK", 0)) == 0:                     # before CUDA / NCCL exist in this process: fork is safe
        pool = mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32))
        jobs = pool.map_async(_oracle_video, [(v, a0.frames, a0.height, a0.width) for v in range(a0.videos)], chunksize=1)

    import numpy as np
    import torch
    import torch.distributed as dist

    import v5ela
    from v5ela.shard import analyze_sharded, shard_range

    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=16)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--height", type=int, default=270)
    ap.add_argument("--width", type=int, default=480)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    v0, v1 = shard_range(a.videos, rank, world)
    chunks = [v5ela.gen_batch_torch(a.frames * v, a.frames, a.height, a.width, seed=v, device=dev) for v in range(v0, v1)]
    frames = torch.cat(chunks, 0) if chunks else torch.empty((0, a.height, a.width, 3), dtype=torch.uint8, device=dev)
    per_video = analyze_sharded(frames, a.videos, quality=90, group_size=a.frames)
    per_frame = analyze_sharded(frames, a.videos * a.frames, quality=90, group_size=0) if a.videos % world == 0 else None
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        from v5ela.batch import get_handle

        got = v5ela.as_records(per_video)
        oracle = {v: (agg, first) for v, agg, first in jobs.get(timeout=1800)}
        pool.close()
        for v in range(a.videos):
            if got[v].tobytes() != oracle[v][0]:
                ok = False
                print(f"video {v}: per-video record differs from the oracle", file=sys.stderr)
        if per_frame is not None:
            gf = v5ela.as_records(per_frame)
            ok = ok and gf[: a.frames].tobytes() == oracle[0][1] and gf.shape[0] == a.videos * a.frames
        hd = get_handle(local)
        print(json.dumps({"check": "config4_per_video_gather", "world": world, "videos": a.videos, "frames_per_video": a.frames,
                          "height": a.height, "width": a.width, "bit_exact_vs_oracle": ok, "videos_checked": a.videos,
                          "instantiation": hd.last_instantiation, "block_stage": hd.block_stage}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

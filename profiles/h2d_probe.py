"""Host-to-device copy ceiling of the `e2e` path, alone and with every rank copying at once (VERDICT r01 task 4).

    python profiles/h2d_probe.py                                                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        profiles/h2d_probe.py                                                          # N ranks, one pinned cudaMemcpyAsync stream each

Every rank owns one pinned host buffer of 256 x 1080p frames (1.59 GB, first-touched after the CPU-affinity step) and copies it to its
GPU with plain `tensor.copy_(non_blocking=True)` = one cudaMemcpyAsync per iteration (nothing batched). Two measurements: every rank
alone while the others idle, then all ranks between two barriers. Prints one JSON line on rank 0: per-rank GB/s in both modes and
the frames/s ceiling they imply for `bench.py`'s `e2e` (which reports the same concurrent figure itself, `e2e.ceiling_frames_s`)."""
import json
import os
import sys

import torch
import torch.distributed as dist

FRAMES, H, W = 256, 1080, 1920


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = None
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        affinity = sorted(os.sched_getaffinity(0))
    except Exception as e:
        affinity = repr(e)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
    nbytes = FRAMES * H * W * 3
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(rank + 1)                                        # first touch on this rank's cores
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(reps=4):
        devbuf.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            devbuf.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    alone = 0.0
    for r in range(world):                                      # one rank at a time
        barrier()
        if r == rank:
            alone = timed()
    barrier()
    together = timed()                                          # every rank between the same two barriers
    barrier()
    t = torch.tensor([alone, together], dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
    else:
        allr = [t]
    if rank == 0:
        al = [round(float(x[0]), 1) for x in allr]
        tg = [round(float(x[1]), 1) for x in allr]
        frame = H * W * 3
        print(json.dumps({"probe": "pinned H2D, one cudaMemcpyAsync of 256 x 1080p frames per iteration", "world": world,
                          "alone_gbs": al, "concurrent_gbs": tg, "aggregate_concurrent_gbs": round(sum(tg), 1),
                          "e2e_ceiling_frames_s_alone": round(sum(a * 1e9 / frame for a in al)),
                          "e2e_ceiling_frames_s_concurrent": round(world * min(tg) * 1e9 / frame),
                          "cpu_affinity_rank0": affinity if not isinstance(affinity, list) else f"{len(affinity)} cpus: {affinity[0]}..{affinity[-1]}"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())

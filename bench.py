#!/usr/bin/env python
"""bench.py — 1080p keyframes/s of the V5 ELA+texture hot path on N B200s (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # N=1: plain python; N>1: launched under torchrun
    python bench.py --impl reference [...]                          # the reference's CPU path on the host cores

A step = one pass of the hot path over one batch of 256 synthetic 1920x1080 RGB keyframes per GPU at JPEG q=90
(BASELINE.json configs[1]; weak scaling: every rank owns its own 256 frames), records-only mode, followed for N>1 by
the NCCL gather of the per-frame records to rank 0.
  value     : whole-job frames/s with inputs resident in HBM (CUDA events, max over ranks).
  e2e       : same metric through the reference-facing C-ABI call with HOST (pinned) buffers — the H2D copy of every
              frame and the D2H copy of the records are inside the timed region.
  roofline  : fused kernel only — algorithmic bytes (3*H*W + 3144 per frame) / its CUDA-event duration, against the
              measured HBM copy bandwidth in MEASURED_PEAKS.json.
  e2e_from_jpeg_files : (N=1, extra) the same work starting from JPEG files in host memory — GPU decode + analyse.
  cpu_baseline : the oracle port of the reference's ELA core (Pillow/libjpeg-turbo + NumPy/OpenCV statistics) on all
              host cores, on a bounded sample of the same frames. Reported, not the target.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

# stdout must carry exactly one JSON line. Native libraries write there too (NCCL prints its version banner to stdout at the
# VERSION and WARN debug levels), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a
# duplicate of the original stdout (emit()).
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "fake-video-detection-engine_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "1080p_keyframes_per_sec_v5_ela_texture"
UNIT = "frames/s"
FRAMES_PER_GPU, H, W, QUALITY = 256, 1080, 1920, 90
RECORD_BYTES = 3144
BYTES_PER_FRAME = 3 * H * W + RECORD_BYTES          # algorithmic bytes, SURVEY.md §8d
HBM_TRAFFIC_NCU = None                              # per-launch dram bytes from profiles/ (filled by load_ncu_traffic)


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": f"{FRAMES_PER_GPU} synthetic {W}x{H} RGB keyframes per GPU (gen_frame, SURVEY App. B), JPEG q={QUALITY}, "
                    "records-only — BASELINE.json configs[1]",
        "frames_per_gpu": FRAMES_PER_GPU, "height": H, "width": W, "quality": QUALITY,
        "global_frames": FRAMES_PER_GPU * n_gpus,
        "l2": "inputs (1.59 GB per GPU) are larger than the 126 MB L2; no flush needed",
        "sharding": "by frame, contiguous per rank; one NCCL gather of 3144-byte records to rank 0 per step" if n_gpus > 1
                    else "single GPU",
    }


# ------------------------------------------------------------------------------------------------- CPU reference arm
_CPU_FRAMES = None


def _cpu_init(frames):
    global _CPU_FRAMES
    _CPU_FRAMES = frames
    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        pass


def _cpu_work(i):
    from oracle import pil_oracle

    rec = pil_oracle.ela_core(_CPU_FRAMES[i % len(_CPU_FRAMES)], QUALITY)
    return int(rec["ela_sum"][0])


def cpu_reference_rate(sample_frames: int, repeats: int = 1, warmup: int = 0):
    """Frames/s of the reference's CPU ELA core (+ §8a statistics) with one process per host core."""
    import multiprocessing as mp

    from oracle import pil_oracle
    from v5ela.synth import gen_frame

    cores = os.cpu_count() or 1
    distinct = [gen_frame(i, H, W, 0) for i in range(min(8, sample_frames))]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(distinct,)) as pool:
        pool.map(_cpu_work, range(cores))                       # spin the workers up
        for _ in range(warmup):
            pool.map(_cpu_work, range(sample_frames), chunksize=1)
        times = []
        for _ in range(max(1, repeats)):
            t0 = time.perf_counter()
            pool.map(_cpu_work, range(sample_frames), chunksize=1)
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return sample_frames * len(times) / total, cores, total / len(times), pil_oracle.versions()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 4 * (os.cpu_count() or 1)                          # per step: a few frames per core (~1-2 s per step)
    fps, cores, step_s, versions = cpu_reference_rate(sample, repeats=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {**workload_config(args.gpus),
                   "note": "reference arm: the reference's own operations (v5_texture_ela.py:66-73: PIL save q=90 -> open -> "
                           "ImageChops.difference -> getextrema) + NumPy/OpenCV record statistics, one process per host core; "
                           "the reference module itself is pure Python over Pillow and is not present on the GPU box",
                   "libraries": versions},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} frames of {W}x{H} per step x {args.steps} steps"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------ clock sampling
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def load_ncu_traffic():
    """Per-launch DRAM bytes of the fused kernel from the committed ncu summary (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "fused_kernel_dram.json")
    try:
        with open(path) as f:
            d = json.load(f)
        if d.get("frames_per_launch") == FRAMES_PER_GPU and d.get("height") == H and d.get("width") == W:
            return float(d["dram_bytes_per_launch"])
        return float(d["dram_bytes_per_frame"]) * FRAMES_PER_GPU
    except Exception:
        return None


def load_issue_stats():
    """ncu-measured issue-slot figures of the fused kernel (the resource that actually binds it), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "fused_kernel_issue.json")) as f:
            return json.load(f)
    except Exception:
        return None


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------- GPU arm
def run_native_arm(args):
    import torch
    import torch.distributed as dist

    import v5ela
    from v5ela import _abi
    from v5ela.batch import analyze_batch, get_handle
    from v5ela.shard import gather_records

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_baseline = None
    if world == 1 and not args.no_cpu:                          # before CUDA is initialised: the pool forks
        sample = 8 * (os.cpu_count() or 1)
        fps, cores, _, versions = cpu_reference_rate(sample_frames=sample, repeats=1)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{sample} frames of {W}x{H} (gen_frame), oracle/pil_oracle.ela_core = the reference's "
                                  "v5_texture_ela.py:66-73 calls + record statistics, one process per core",
                        "libraries": versions}
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU implementation (use --impl reference)")
    _abi.load()
    try:    # run (and first-touch the pinned host buffers) on the CPU cores next to this rank's GPU: matters for `e2e` at N > 1
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    n_gpus = world
    K, Wm = args.steps, args.warmup
    n_local, total = FRAMES_PER_GPU, FRAMES_PER_GPU * world

    frames = v5ela.gen_batch_torch(rank * n_local, n_local, H, W, seed=0, device=dev)
    records = torch.empty((n_local, RECORD_BYTES), dtype=torch.uint8, device=dev)
    handle = get_handle(local_rank)

    def step():
        analyze_batch(frames, quality=QUALITY, records_out=records, handle=handle)
        return gather_records(records, total) if world > 1 else records

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(Wm, 3)):
        step()
    barrier()

    # ---- device-resident throughput (value) + fused-kernel duration (roofline), clocks sampled during the region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    handle.profile_enable(True)
    handle.profile_read(reset=True)
    launches0 = handle.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = handle.launch_count - launches0
    fused_ms, fused_n = handle.profile_read(reset=True)
    handle.profile_enable(False)
    # keep the GPU busy a little longer so that slow nvidia-smi polling still sees the load (local work only: the number
    # of iterations depends on the wall clock, so no collective may be issued here)
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        analyze_batch(frames, quality=QUALITY, records_out=records, handle=handle)
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- end to end through the host-buffer C-ABI entry point (pinned host memory)
    host_frames = torch.empty((n_local, H, W, 3), dtype=torch.uint8, pin_memory=True)
    host_frames.copy_(frames)
    host_records = torch.empty((n_local, RECORD_BYTES), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    Ke = max(2, min(K, 10))
    stream = torch.cuda.current_stream(dev).cuda_stream

    def e2e_step():
        handle.analyze_host(host_frames.data_ptr(), n_local, H, W, host_records.data_ptr(), None, None, stream)
        if world > 1:
            recs = host_records.to(dev, non_blocking=True)      # records already on the host: gather via device
            gather_records(recs, total)

    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(Ke):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    ok_e2e = bool(torch.equal(host_records, records.cpu()))

    # ---- the same work starting from JPEG FILES in host memory, the form V1 hands keyframes/crops over in
    # (cv2.imwrite default quality 95, v1_keyframes_facetrack.py:112,166): only compressed bytes cross PCIe; the GPU
    # decodes, analyses, and the records come back. N=1 only (an extra figure, not the headline).
    files_leg = None
    if world == 1:
        try:
            from v5ela import jpeg
            from v5ela.batch import analyze_jpeg_files

            enc, enc_sizes = jpeg.encode_batch(frames, 95)                  # warm-up (workspace allocation), then timed once
            torch.cuda.synchronize()
            del enc, enc_sizes                                              # let the timed call reuse the 1.6 GB output block
            e0.record()
            enc, enc_sizes = jpeg.encode_batch(frames, 95)
            e1.record()
            torch.cuda.synchronize()
            encode_fps = n_local / (e0.elapsed_time(e1) * 1e-3)
            enc, enc_sizes = enc.cpu().numpy(), enc_sizes.cpu().numpy()
            # the files sit back to back in one page-locked arena (what a loader that reads files for the GPU would use)
            offs = [0]
            for i in range(n_local):
                offs.append(offs[-1] + (int(enc_sizes[i]) + 63) // 64 * 64)
            arena = torch.empty(offs[-1], dtype=torch.uint8, pin_memory=True)
            arena_np = arena.numpy()
            blobs = []
            for i in range(n_local):
                arena_np[offs[i]:offs[i] + enc_sizes[i]] = enc[i, :enc_sizes[i]]
                blobs.append(arena_np[offs[i]:offs[i] + int(enc_sizes[i])])
            del enc
            Kf = max(2, min(K, 5))
            out = None
            for _ in range(2):
                out = analyze_jpeg_files(blobs, quality=QUALITY, device=dev)
                host_records.copy_(out["records"], non_blocking=True)
            torch.cuda.synchronize()
            same = bool(torch.equal(analyze_batch(out["rgb"], quality=QUALITY)["records"].cpu(), host_records))
            t0 = time.perf_counter()
            for _ in range(Kf):
                out = analyze_jpeg_files(blobs, quality=QUALITY, device=dev)
                host_records.copy_(out["records"], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            files_leg = {"value": n_local * Kf / dt, "unit": UNIT, "steps": Kf,
                         "h2d_bytes_per_step": int(sum(len(b) for b in blobs)), "d2h_bytes_per_step": n_local * RECORD_BYTES,
                         "input": f"{n_local} JPEG files (4:2:0, quality 95, mean {sum(len(b) for b in blobs) / n_local / 1e3:.0f} kB) "
                                  "in one pinned host arena; header parsing on the host, wall clock",
                         "decode_status_ok": bool((out["status"] == 0).all().item()), "records_match_decoded_frames": same,
                         "files_written_by": "v5ela_jpeg_encode on the GPU (== cv2.imwrite's bytes), device-resident frames in, "
                                             f"{encode_fps:.0f} files/s (not part of the timed region)",
                         "api": "v5ela_jpeg_decode + v5ela_analyze (C ABI)"}
            del out
        except Exception as e:  # an extra figure must not take the headline down with it
            files_leg = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = load_peak()
        value = total * K / (ms_max * 1e-3)
        kernel_ms = fused_ms / max(fused_n, 1)
        achieved = n_local * BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": max(Wm, 3),
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload_config(n_gpus),
            "e2e": {"value": total * Ke / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n_local * 3 * H * W, "d2h_bytes_per_step": n_local * RECORD_BYTES,
                    "steps": Ke, "records_match_device_path": ok_e2e,
                    "api": "v5ela_analyze_host (C ABI, pinned host buffers, chunked copy/compute overlap)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_ncu_traffic(), "kernel": "v5::ela_fused_kernel", "kernel_ms": kernel_ms,
                         "kernel_launches_timed": int(fused_n), "bytes_per_launch": n_local * BYTES_PER_FRAME,
                         "peak_source": peak_src,
                         "note": "integer-issue bound, not HBM bound: ~%d exact int32 thread-instructions per pixel "
                                 "(DESIGN.md 4.4); int_issue = ncu figures of the committed profile"
                                 % round(32 * ((load_issue_stats() or {}).get("warp_instructions_per_pixel") or 3.83)),
                         "kernel_instantiation": "ela_fused_kernel<FAST=true, TEXHIST=false> (width % 16 == 0, records only)",
                         "int_issue": load_issue_stats()},
        }
        if files_leg is not None:
            line["e2e_from_jpeg_files"] = files_leg
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("native", "reference"), default="native")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_native_arm(args)


if __name__ == "__main__":
    sys.exit(main())

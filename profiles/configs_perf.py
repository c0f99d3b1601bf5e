"""Device-resident throughput of the other BASELINE.json configurations on ONE GPU (parity for them lives in tests/).
Prints one line per configuration; results are committed under profiles/rNN/configs.txt."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200"))
import torch, v5ela
from v5ela.batch import get_handle

PEAK = 6464.9

def run(label, frames, quality=90, residual=False, group=0, iters=5):
    n, h, w, _ = frames.shape
    recs = torch.empty((n, 3144), dtype=torch.uint8, device="cuda")
    res = torch.empty_like(frames) if residual else None
    hd = get_handle(0)
    def step():
        out = v5ela.analyze_batch(frames, quality=quality, want_residual=residual, records_out=recs, residual_out=res)
        if group:
            v5ela.reduce_records(out["records"], group)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fps = n / ms * 1e3
    gbs = fps * (3 * h * w * (2 if residual else 1) + 3144) / 1e9
    print(f"{label:58s} {ms:8.3f} ms/batch {fps:10,.0f} frames/s {gbs:7.1f} GB/s algorithmic = {gbs / PEAK * 100:5.2f}% of measured HBM peak")

g = v5ela.gen_batch_torch
run("config 1 shape: 16 x 1280x720 q90", g(0, 16, 720, 1280, 0))
run("config 2: 256 x 1920x1080 q90 (bench workload)", g(0, 256, 1080, 1920, 0))
run("config 2 + residual map output", g(0, 256, 1080, 1920, 0), residual=True)
run("config 3 per-GPU share: 128 x 3840x2160 q90", g(0, 128, 2160, 3840, 0))
t = torch.cat([g(32 * v, 32, 1080, 1920, seed=v) for v in range(8)], 0)
run("config 4 per-GPU share: 8 videos x 32 x 1080p + per-video reduce", t, group=32)
del t
f = g(0, 256, 1080, 1920, 0)
for q in (75, 85, 90, 95):
    run(f"config 5: 256 x 1080p q{q}", f, quality=q)
del f
run("uniform-noise content: 128 x 1080p q90 (residuals up to 255)", torch.randint(0, 256, (128, 1080, 1920, 3), dtype=torch.uint8, device="cuda"))
run("single frame latency shape: 1 x 1080p", g(0, 1, 1080, 1920, 0), iters=20)
run("3 face crops 257x301 (node-sized call)", g(0, 3, 257, 301, 0), iters=20)

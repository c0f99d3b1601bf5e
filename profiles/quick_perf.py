"""Quick device-resident throughput probe (development aid; bench.py is the judged harness)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200"))
import torch
import v5ela

def run(n, h, w, residual=False, iters=5, noise=False):
    t = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda") if noise else v5ela.gen_batch_torch(0, n, h, w, 0)
    recs = torch.empty((n, 3144), dtype=torch.uint8, device="cuda")
    res = torch.empty_like(t) if residual else None
    for _ in range(2):
        v5ela.analyze_batch(t, want_residual=residual, records_out=recs, residual_out=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        v5ela.analyze_batch(t, want_residual=residual, records_out=recs, residual_out=res)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fps = n / ms * 1e3
    gbs = fps * (3 * h * w + 3144) / 1e9
    print(f"n={n} {w}x{h} residual={residual} noise={noise}: {ms:.3f} ms/batch  {fps:,.0f} frames/s  {gbs:.1f} GB/s algorithmic ({gbs/6464.9*100:.2f}% of measured HBM)")

if __name__ == "__main__":
    run(64, 1080, 1920)
    run(256, 1080, 1920)
    run(256, 1080, 1920, residual=True)
    run(64, 1080, 1920, noise=True)
    run(64, 2160, 3840)
    run(256, 720, 1280)

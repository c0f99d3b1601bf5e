"""Development aid: throughput of the spectrum path and its exact-match rate vs NumPy."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fake-video-detection-engine_b200"))
import numpy as np, torch, v5ela
from oracle import pil_oracle
for (n, h, w) in [(64, 257, 301), (16, 720, 1280), (8, 1080, 1920)]:
    g = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda")
    for _ in range(2): v5ela.spectrum_batch(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = v5ela.spectrum_batch(g)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    t0 = time.perf_counter(); ref = pil_oracle.fft_spectrum(g[0].cpu().numpy()); cpu = time.perf_counter() - t0
    d = np.abs(out[0].cpu().numpy().astype(np.int16) - ref.astype(np.int16))
    wh = w // 2 + 1
    gbs = n * (h * w * 2 + h * wh * (16 * 2 + 8 * 2)) / ms / 1e6       # gray in, image out, G written+read, ms written+read
    print(f"{n} x {w}x{h}: {ms:8.3f} ms/batch = {n / ms * 1e3:9,.0f} images/s ({gbs:6.1f} GB/s of intermediate+io traffic)  numpy 1 thread {1 / cpu:6.1f} images/s  max diff {d.max()}  exact {float((d == 0).mean()) * 100:.4f}%")

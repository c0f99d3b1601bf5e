"""Profiling target: the spectrum path on 8 x 1080p and 16 x 720p luma planes (three launches each)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200"))
import torch
import v5ela

for n, h, w in ((8, 1080, 1920), (16, 720, 1280)):
    g = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        out = v5ela.spectrum_batch(g)
    torch.cuda.synchronize()
print("ok", int(out.sum().item()))

"""Development aid: sweep of the work-decomposition knobs (V5ELA_DECOMP = items per CTA, tail segment rows, tail size in eighths of
a strip column per CTA) on the benchmark batch and two others; CUDA-event kernel time."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200"))
import torch, v5ela
from v5ela import _abi
sets = [("256x1080p", v5ela.gen_batch_torch(0, 256, 1080, 1920, 0)), ("64x4K", v5ela.gen_batch_torch(0, 64, 2160, 3840, 0)),
        ("256x720p", v5ela.gen_batch_torch(0, 256, 720, 1280, 0))]
recs = torch.empty((256, 3144), dtype=torch.uint8, device="cuda")
for knob in ("3,8,4", "3,8,0", "2,8,4", "4,8,4", "6,8,4", "3,4,4", "3,6,4", "3,12,4", "3,8,2", "3,8,6", "3,8,8", "3,8,12", "2,8,8", "2,6,6"):
    os.environ["V5ELA_DECOMP"] = knob
    hd = _abi.Handle(0)
    out = []
    for name, frames in sets:
        n = frames.shape[0]
        for _ in range(3):
            v5ela.analyze_batch(frames, records_out=recs[:n], handle=hd)
        torch.cuda.synchronize()
        hd.profile_enable(True); hd.profile_read(True)
        for _ in range(8):
            v5ela.analyze_batch(frames, records_out=recs[:n], handle=hd)
        ms, cnt = hd.profile_read(True)
        out.append(f"{name} {n / (ms / cnt) * 1e3:9,.0f}")
    print(f"V5ELA_DECOMP={knob:8s} " + "   ".join(out))
    hd.close()

"""Development aid: fused-kernel frame rate (records-only calls, CUDA-event kernel time) of both block-stage builds over several
frame geometries, batches of equal pixel count (a per-geometry choice between the builds was considered and dropped: they are equal)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200"))
import torch, v5ela
from v5ela.batch import get_handle

hd = get_handle(0)
for (h, w) in ((360, 640), (540, 960), (720, 1280), (900, 1600), (1080, 1920), (1440, 2560), (2160, 3840), (2160, 4096)):
    n = max(8, int(256 * 1080 * 1920 / (h * w)))
    t = v5ela.gen_batch_torch(0, n, h, w, 0)
    recs = torch.empty((n, 3144), dtype=torch.uint8, device="cuda")
    row = []
    for stage in ("smem", "mma"):
        hd.block_stage = stage
        for _ in range(3):
            v5ela.analyze_batch(t, records_out=recs)
        torch.cuda.synchronize()
        hd.profile_enable(True); hd.profile_read(True)
        for _ in range(6):
            v5ela.analyze_batch(t, records_out=recs)
        ms, cnt = hd.profile_read(True)
        hd.profile_enable(False)
        row.append((stage, n / (ms / cnt) * 1e3))
    mw = (w + 15) // 16; ns = (mw + 29) // 30; tw = (mw + ns - 1) // ns; blocks = 4 * tw + 2 * (tw + 2)
    print(f"{w}x{h} x{n:4d}  widest strip {tw:2d} MCUs, round packing {blocks / (((blocks + 63) // 64) * 64):.3f}: " +
          "  ".join(f"{s} {fps:10,.0f} fps" for s, fps in row) +
          f"   mma/smem {row[1][1] / row[0][1]:.3f}")
    del t
hd.block_stage = "smem"

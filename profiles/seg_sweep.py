import os, sys
sys.path.insert(0, "fake-video-detection-engine_b200")
import torch, v5ela
from v5ela import _abi
t = v5ela.gen_batch_torch(0, 256, 1080, 1920, 0)
t2 = v5ela.gen_batch_torch(0, 259, 1080, 1920, 0)
recs = torch.empty((259, 3144), dtype=torch.uint8, device="cuda")
for seg in (9, 12, 14, 17, 20, 23, 34, 68):
    os.environ["V5ELA_SEG_ROWS"] = str(seg)
    hd = _abi.Handle(0)
    out = []
    for frames in (t, t2):
        n = frames.shape[0]
        for _ in range(3):
            v5ela.analyze_batch(frames, records_out=recs[:n], handle=hd)
        torch.cuda.synchronize()
        hd.profile_enable(True); hd.profile_read(True)
        for _ in range(8):
            v5ela.analyze_batch(frames, records_out=recs[:n], handle=hd)
        ms, cnt = hd.profile_read(True)
        out.append(n / (ms / cnt) * 1e3)
    items = 256 * 4 * ((68 + seg - 1) // seg)
    print(f"seg_rows {seg:2d}: items per 256 frames {items:5d} ({items / 296:.2f} rounds of 296 CTAs)  256 frames {out[0]:9,.0f} fps   259 frames {out[1]:9,.0f} fps")
    hd.close()

"""Development aid: time one library build (V5ELA_LIB) on the bench workload; prints one line."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200"))
import torch, v5ela
from v5ela.batch import get_handle

def run(n, h, w, iters=6):
    t = v5ela.gen_batch_torch(0, n, h, w, 0)
    recs = torch.empty((n, 3144), dtype=torch.uint8, device="cuda")
    hd = get_handle(0)
    for _ in range(3):
        v5ela.analyze_batch(t, records_out=recs)
    torch.cuda.synchronize()
    hd.profile_enable(True); hd.profile_read(True)
    for _ in range(iters):
        v5ela.analyze_batch(t, records_out=recs)
    ms, cnt = hd.profile_read(True)
    return n / (ms / cnt) * 1e3, recs

fps, recs = run(256, 1080, 1920)
chk = int(recs.to(torch.int64).sum().item())
fps4k, _ = run(64, 2160, 3840)
fps720, _ = run(256, 720, 1280)
print(f"{os.path.basename(os.environ.get('V5ELA_LIB', 'default')):28s} 1080p {fps:9,.0f} fps   4K {fps4k:8,.0f} fps   720p {fps720:9,.0f} fps   checksum {chk}")

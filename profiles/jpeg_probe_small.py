"""ncu target: one encode + one decode of a few 1080p frames through the codec kernels (no timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fake-video-detection-engine_b200")]
import torch, v5ela
from v5ela import jpeg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = v5ela.gen_batch_torch(0, n, 1080, 1920, seed=0, device="cuda")
files, sizes = jpeg.encode_batch(frames, 90)
torch.cuda.synchronize()
sz = sizes.cpu().numpy()
blobs = [files[i, :int(sz[i])].cpu().numpy().tobytes() for i in range(n)]
out = jpeg.decode_batch(blobs)
torch.cuda.synchronize()
print("ok", int((out["status"] == 0).all()), float(sz.mean()))

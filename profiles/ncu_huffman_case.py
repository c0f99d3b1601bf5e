"""One decode of 148 JPEG files (1080p, quality 95, 4:2:0 — the files bench.py's files-in leg uses) = one full wave of
huffman_kernel, for ncu:  ncu --set full --import-source on -k regex:huffman_kernel -s 1 -c 1 python profiles/ncu_huffman_case.py
Prints the CUDA-event time of the decode call (when run without ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fake-video-detection-engine_b200")]
import torch  # noqa: E402

import v5ela  # noqa: E402
from v5ela import jpeg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
frames = v5ela.gen_batch_torch(0, n, 1080, 1920, seed=0, device="cuda")
files, sizes = jpeg.encode_batch(frames, 95)
sz = sizes.cpu().numpy()
blobs = [files[i, :int(sz[i])].cpu().numpy().tobytes() for i in range(n)]
for _ in range(2):
    out = jpeg.decode_batch(blobs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = jpeg.decode_batch(blobs)
e1.record()
torch.cuda.synchronize()
print(f"{n} files, mean {sz.mean() / 1e3:.0f} kB: decode call {e0.elapsed_time(e1):.2f} ms, status ok {bool((out['status'] == 0).all())}")

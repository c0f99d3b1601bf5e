"""Profiling target: three records-only launches of the fused kernel on the benchmark batch (256 x 1080p, q=90).
    python profiles/ncu_fused_case.py && ncu --set full --clock-control none --import-source on -k regex:ela_fused -s 2 -c 1 \
        -o gpurun_out/prof python profiles/ncu_fused_case.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200"))
import torch
import v5ela

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
t = v5ela.gen_batch_torch(0, n, 1080, 1920, 0)
recs = torch.empty((n, 3144), dtype=torch.uint8, device="cuda")
for _ in range(3):
    v5ela.analyze_batch(t, records_out=recs)
torch.cuda.synchronize()
print("ok", int(recs.to(torch.int64).sum().item()))

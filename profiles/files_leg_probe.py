"""Where the time goes in bench.py's e2e_from_jpeg_files leg (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fake-video-detection-engine_b200")]
import numpy as np, torch, v5ela
from v5ela import jpeg
from v5ela.batch import analyze_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
q = int(sys.argv[2]) if len(sys.argv) > 2 else 95
frames = v5ela.gen_batch_torch(0, n, 1080, 1920, seed=0, device="cuda")
enc, sizes = jpeg.encode_batch(frames, q)
torch.cuda.synchronize()
enc, sizes = enc.cpu().numpy(), sizes.cpu().numpy()
offs = [0]
for i in range(n):
    offs.append(offs[-1] + (int(sizes[i]) + 63) // 64 * 64)
arena = torch.empty(offs[-1], dtype=torch.uint8, pin_memory=True)
a = arena.numpy()
blobs = []
for i in range(n):
    a[offs[i]:offs[i] + sizes[i]] = enc[i, :sizes[i]]
    blobs.append(a[offs[i]:offs[i] + int(sizes[i])])
rec_host = torch.empty((n, 3144), dtype=torch.uint8, pin_memory=True)
for _ in range(2):
    out = jpeg.decode_batch(blobs)
    r = analyze_batch(out["rgb"])
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(3):
    t0 = time.perf_counter()
    ev[0].record()
    out = jpeg.decode_batch(blobs)
    t1 = time.perf_counter()
    ev[1].record()
    r = analyze_batch(out["rgb"])
    ev[2].record()
    rec_host.copy_(r["records"], non_blocking=True)
    ev[3].record()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"host: decode_batch call {1e3 * (t1 - t0):6.2f} ms, analyze+copy calls {1e3 * (t2 - t1):6.2f} ms, wait {1e3 * (t3 - t2):6.2f} ms | "
          f"GPU: decode {ev[0].elapsed_time(ev[1]):6.2f} ms (includes the host time of the call), analyze {ev[1].elapsed_time(ev[2]):6.2f} ms, "
          f"records D2H {ev[2].elapsed_time(ev[3]):5.2f} ms | total {1e3 * (t3 - t0):6.2f} ms")
# python-side pieces
t0 = time.perf_counter(); tab = jpeg._file_table(blobs); t1 = time.perf_counter(); d = jpeg.info_batch(blobs, tab); t2 = time.perf_counter()
x = torch.empty((n, 1080, 1920, 3), dtype=torch.uint8, device="cuda"); t3 = time.perf_counter()
print(f"python: file table {1e3 * (t1 - t0):.2f} ms, info_batch {1e3 * (t2 - t1):.2f} ms, torch.empty {1e3 * (t3 - t2):.2f} ms")
if os.environ.get("V5_TORCH_PROFILER"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            out = jpeg.decode_batch(blobs)
            r = analyze_batch(out["rgb"])
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))

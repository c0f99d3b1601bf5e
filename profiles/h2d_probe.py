import torch, time
x = torch.empty(1592524800, dtype=torch.uint8, pin_memory=True)
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"H2D pinned 1.59 GB: {ms:.2f} ms = {x.numel()/ms/1e6:.1f} GB/s -> {256/ms*1e3:.0f} 1080p frames/s ceiling")
s2 = torch.cuda.Stream()
h = x.numel() // 2
e0.record()
for _ in range(5):
    d[:h].copy_(x[:h], non_blocking=True)
    with torch.cuda.stream(s2):
        d[h:].copy_(x[h:], non_blocking=True)
torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(f"two streams: {x.numel()*5/(e0.elapsed_time(e1))/1e6:.1f} GB/s")

"""Codec rows (§8f-2/§8f-3) throughput on one B200: device-resident batches, CUDA-event timing around the library calls.
Usage: python profiles/jpeg_perf.py [n_frames]   (prints a small table; copy it to profiles/rNN/jpeg.txt)"""
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fake-video-detection-engine_b200")]
import torch  # noqa: E402

import v5ela  # noqa: E402
from v5ela import jpeg  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    print(f"device {torch.cuda.get_device_name(0)}; {n} frames per call; best of 5, CUDA events")
    for (h, w, q) in ((1080, 1920, 90), (1080, 1920, 75), (720, 1280, 90), (2160, 3840, 90)):
        cnt = n if h < 2000 else max(1, n // 4)
        frames = v5ela.gen_batch_torch(0, cnt, h, w, seed=0, device="cuda")
        files, sizes = jpeg.encode_batch(frames, q)
        torch.cuda.synchronize()
        t_enc = timed(lambda: jpeg.encode_batch(frames, q))
        sz = sizes.cpu().numpy()
        blobs = [files[i, :int(sz[i])].cpu().numpy().tobytes() for i in range(cnt)]
        t0 = time.perf_counter()
        out = jpeg.decode_batch(blobs)
        torch.cuda.synchronize()
        t_first = (time.perf_counter() - t0) * 1e3
        t_dec = timed(lambda: jpeg.decode_batch(blobs))
        t0 = time.perf_counter()
        for _ in range(3):
            jpeg.decode_batch(blobs)
        torch.cuda.synchronize()
        t_dec_wall = (time.perf_counter() - t0) / 3 * 1e3
        gray = frames[..., 1].contiguous()
        t_enc_g = timed(lambda: jpeg.encode_batch(gray, 95))
        print(f"{h}x{w} q{q}: mean file {sz.mean() / 1e3:8.1f} kB | encode RGB {cnt / t_enc * 1e3:9.0f} img/s ({t_enc:7.2f} ms)"
              f" | encode gray q95 {cnt / t_enc_g * 1e3:9.0f} img/s | decode {cnt / t_dec_wall * 1e3:9.0f} img/s (wall clock per call incl. header"
              f" parsing and upload {t_dec_wall:7.2f} ms back to back; one call alone {t_dec:7.2f} ms; first call {t_first:7.1f} ms) status ok={int((out['status'] == 0).all())}")
    # restart intervals (IMWRITE_JPEG_RST_INTERVAL, one interval per MCU row / per 8 MCUs): files OpenCV writes on request; the
    # decoder takes them through huffman_rst_kernel (one thread per interval)
    import cv2
    for interval in (120, 8):
        frames_np = v5ela.gen_batch(0, min(n, 32), 1080, 1920, seed=0)
        blobs = [cv2.imencode(".jpg", np.ascontiguousarray(f[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, interval])[1].tobytes()
                 for f in frames_np]
        out = jpeg.decode_batch(blobs)
        torch.cuda.synchronize()
        ok = bool((out["status"] == 0).all()) and bool(np.array_equal(out["rgb"][0].cpu().numpy(),
                                                                     cv2.imdecode(np.frombuffer(blobs[0], np.uint8), cv2.IMREAD_COLOR)[..., ::-1]))
        t0 = time.perf_counter()
        for _ in range(3):
            jpeg.decode_batch(blobs)
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 3 * 1e3
        print(f"1080x1920 q90, restart interval {interval} MCUs ({(120 * 68 + interval - 1) // interval} intervals per file), {len(blobs)} files: decode "
              f"{len(blobs) / t * 1e3:9.0f} img/s ({t:7.2f} ms per call, wall clock) pixels == OpenCV: {ok}")
    # the CPU libraries on this box, one thread, same work
    from PIL import Image
    import cv2
    a = v5ela.gen_frame(0, 1080, 1920, 0)
    t0 = time.perf_counter()
    for _ in range(5):
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, "JPEG", quality=90)
    t_pe = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(5):
        np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    t_pd = (time.perf_counter() - t0) / 5
    print(f"Pillow {Image.__version__} on one host thread, 1080p q90: encode {1 / t_pe:6.1f} img/s, decode {1 / t_pd:6.1f} img/s; "
          f"OpenCV {cv2.__version__}")


if __name__ == "__main__":
    main()

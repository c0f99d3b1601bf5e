"""Corrupt-input fuzz of the GPU decoder (development aid, run outside pytest: a device fault would poison the process).
Valid headers, damaged entropy-coded data: the call must return (ok or error), never fault, and a valid file decoded
afterwards must still equal Pillow's pixels."""
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fake-video-detection-engine_b200")]
from PIL import Image  # noqa: E402

import torch  # noqa: E402
import v5ela  # noqa: E402
from v5ela import _abi, jpeg  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(99)
    good = []
    for it, (h, w, q) in enumerate([(64, 64, 90), (333, 517, 75), (720, 1280, 95), (1080, 1920, 90)]):
        buf = io.BytesIO()
        Image.fromarray(v5ela.gen_frame(it, h, w, 1)).save(buf, "JPEG", quality=q)
        good.append(buf.getvalue())
    refs = [np.asarray(Image.open(io.BytesIO(g)).convert("RGB")) for g in good]
    ok = err = 0
    for it in range(cases):
        g = bytearray(good[it % len(good)])
        start = g.index(b"\xff\xda") + 14
        kind = it % 5
        if kind == 0:                                   # random bytes flipped
            for pos in rng.integers(start, len(g) - 2, int(rng.integers(1, 40))):
                g[pos] = int(rng.integers(0, 256))
        elif kind == 1:                                 # a stretch of noise
            a = int(rng.integers(start, len(g) - 2))
            b = min(len(g) - 2, a + int(rng.integers(1, 5000)))
            g[a:b] = rng.integers(0, 256, b - a, dtype=np.uint8).tobytes()
        elif kind == 2:                                 # truncated
            g = g[:int(rng.integers(start, len(g)))]
        elif kind == 3:                                 # all ones / all zeros from some point
            a = int(rng.integers(start, len(g) - 2))
            g[a:len(g) - 2] = bytes([0xFF if it % 2 else 0x00]) * (len(g) - 2 - a)
        else:                                           # bytes deleted
            a = int(rng.integers(start, len(g) - 10))
            del g[a:a + int(rng.integers(1, 64))]
        try:
            jpeg.decode_host([bytes(g)], want_rgb=True, want_gray=True)
            ok += 1
        except _abi.V5ElaError:
            err += 1
        if it % 25 == 24:
            outs = jpeg.decode_host(good)
            assert all(np.array_equal(o["rgb"], r) for o, r in zip(outs, refs)), "a valid file no longer decodes correctly"
            torch.cuda.synchronize()
    print(f"fuzz: {cases} damaged files -> {ok} decoded to something, {err} reported as errors, no fault; valid files still exact")


if __name__ == "__main__":
    main()

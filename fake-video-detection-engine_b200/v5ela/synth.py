"""Deterministic synthetic keyframes (SURVEY.md Appendix B generator).

Pure uint32 wrap-around integer arithmetic so NumPy (host, tests/oracle) and torch (device, bench) produce the
same bytes. ``gen_frame(0, 720, 1280, 0)`` hashes to ``4e73222f1a84a534`` (first 16 hex of SHA-256 over the HWC bytes).

The content is "natural-ish": three triangle-wave gradients plus +-6 hash noise, which after a q=90 JPEG round
trip leaves residuals in bins 0..~16 -- the case that stresses histogram contention (uniform noise would hide it).
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF


def _frame_key(n: int, seed: int) -> int:
    return (n * 2654435761 + seed * 40503) & _M32


def gen_frame(n: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """Frame ``n`` of stream ``seed`` as an (h, w, 3) uint8 RGB array (host / NumPy)."""
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]

    def tri(t, p):
        return np.abs((t % (2 * p)) - p) * 255 // p

    base = np.stack(
        [
            np.broadcast_to(tri(x + 3 * n, 97), (h, w)),
            np.broadcast_to(tri(y + 5 * n, 61), (h, w)),
            tri(x + y + 7 * n, 131),
        ],
        axis=-1,
    )
    xs = (x.astype(np.uint64) * 73856093) & _M32
    ys = (y.astype(np.uint64) * 19349663) & _M32
    cs = (np.arange(3, dtype=np.uint64) * 83492791) & _M32
    hsh = (xs ^ ys)[..., None] ^ cs[None, None, :] ^ np.uint64(_frame_key(n, seed))
    hsh ^= hsh >> np.uint64(13)
    hsh = (hsh * np.uint64(0x5BD1E995)) & np.uint64(_M32)
    hsh ^= hsh >> np.uint64(15)
    noise = (hsh % np.uint64(13)).astype(np.int64) - 6
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def gen_batch(first: int, count: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """Frames ``first .. first+count-1`` stacked as (count, h, w, 3) uint8 (host / NumPy)."""
    return np.stack([gen_frame(first + i, h, w, seed) for i in range(count)], axis=0)


def gen_batch_torch(first: int, count: int, h: int, w: int, seed=0, device="cuda", out=None):
    """Same frames generated directly on ``device`` (int64 arithmetic masked to 32 bits).

    ``seed`` may be an int or a per-frame sequence of ints (config 4: video v, frame k -> seed=v).
    """
    import torch

    dev = torch.device(device)
    if out is None:
        out = torch.empty((count, h, w, 3), dtype=torch.uint8, device=dev)
    x = torch.arange(w, dtype=torch.int64, device=dev)[None, :]
    y = torch.arange(h, dtype=torch.int64, device=dev)[:, None]
    cs = (torch.arange(3, dtype=torch.int64, device=dev) * 83492791) & _M32
    xy = (((x * 73856093) & _M32) ^ ((y * 19349663) & _M32))[..., None] ^ cs[None, None, :]

    def tri(t, p):
        return ((t % (2 * p)) - p).abs() * 255 // p

    for i in range(count):
        n = first + i
        s = int(seed[i]) if hasattr(seed, "__len__") else int(seed)
        base = torch.stack(
            [tri(x + 3 * n, 97).expand(h, w), tri(y + 5 * n, 61).expand(h, w), tri(x + y + 7 * n, 131)], dim=-1
        )
        hsh = xy ^ _frame_key(n, s)
        hsh = hsh ^ (hsh >> 13)
        hsh = (hsh * 0x5BD1E995) & _M32
        hsh = hsh ^ (hsh >> 15)
        out[i] = (base + (hsh % 13) - 6).clamp_(0, 255).to(torch.uint8)
    return out

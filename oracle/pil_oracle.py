"""Live oracle: the reference's own operations on in-memory arrays. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every step below is the call the reference makes, with the file round trip replaced by BytesIO (byte-identical
JPEG stream) — reference lines cited per function are in ``/root/reference/nodes/V_nodes/v5_texture_ela.py``.
The integer statistics that have no reference counterpart (SURVEY.md §8a "V5F v1") are computed with NumPy/OpenCV
over PIL's residual exactly as §8a defines them.
"""
from __future__ import annotations

import io

import numpy as np
from PIL import Image, ImageChops, ImageEnhance

RECORD_DTYPE = np.dtype(
    [
        ("ela_hist", "<u4", (3, 256)),
        ("ela_sum", "<u8", (3,)),
        ("ela_sumsq", "<u8", (3,)),
        ("tex_sumabs", "<u8"),
        ("tex_sumsq", "<u8"),
        ("tex_maxabs", "<u2"),
        ("ela_max", "u1", (3,)),
        ("pad", "u1", (3,)),
    ]
)
assert RECORD_DTYPE.itemsize == 3144


def versions() -> dict:
    import cv2
    from PIL import features

    return {
        "pillow": Image.__version__ if hasattr(Image, "__version__") else __import__("PIL").__version__,
        "libjpeg_turbo": features.version_feature("libjpeg_turbo"),
        "opencv": cv2.__version__,
        "numpy": np.__version__,
    }


def reencode(rgb: np.ndarray, quality: int = 90) -> np.ndarray:
    """``original.save(tmp,'JPEG',quality=q)`` + ``Image.open(tmp)`` (v5_texture_ela.py:66-68) -> decoded RGB."""
    buf = io.BytesIO()
    Image.fromarray(rgb, "RGB").save(buf, "JPEG", quality=quality)
    buf.seek(0)
    return np.asarray(Image.open(buf).convert("RGB"))


def ela_residual(rgb: np.ndarray, quality: int = 90):
    """v5_texture_ela.py:66-76 -> (residual HWC uint8, max_diff with the 0->1 fix, scale)."""
    original = Image.fromarray(rgb, "RGB")
    buf = io.BytesIO()
    original.save(buf, "JPEG", quality=quality)
    buf.seek(0)
    compressed = Image.open(buf)
    diff = ImageChops.difference(original, compressed)
    extrema = diff.getextrema()
    max_diff = max(ex[1] for ex in extrema)
    if max_diff == 0:
        max_diff = 1
    return np.asarray(diff), max_diff, 255.0 / max_diff


def ela_enhanced(rgb: np.ndarray, quality: int = 90) -> np.ndarray:
    """v5_texture_ela.py:66-78 -> the brightness-enhanced residual image (before it is saved as ela_i.jpg)."""
    original = Image.fromarray(rgb, "RGB")
    buf = io.BytesIO()
    original.save(buf, "JPEG", quality=quality)
    buf.seek(0)
    diff = ImageChops.difference(original, Image.open(buf))
    max_diff = max(ex[1] for ex in diff.getextrema()) or 1
    return np.asarray(ImageEnhance.Brightness(diff).enhance(255.0 / max_diff))


def luma(rgb: np.ndarray) -> np.ndarray:
    """libjpeg / PIL ``convert('L')`` luma of the original frame."""
    return np.asarray(Image.fromarray(rgb, "RGB").convert("L"))


def texture_stats(rgb: np.ndarray):
    """§8a texture fields: cv2.Laplacian(Y, CV_16S, ksize=1) -> (sum|L|, sum L^2, max|L|)."""
    import cv2

    lap = cv2.Laplacian(luma(rgb), cv2.CV_16S, ksize=1).astype(np.int64)
    a = np.abs(lap)
    return int(a.sum()), int((lap * lap).sum()), int(a.max())


def texture_hist(rgb: np.ndarray):
    """§8a optional field tex_hist[256]: bincount(min(|cv2.Laplacian(Y, CV_16S, ksize=1)|, 255))."""
    import cv2

    lap = cv2.Laplacian(luma(rgb), cv2.CV_16S, ksize=1).astype(np.int64)
    return np.bincount(np.minimum(np.abs(lap), 255).ravel(), minlength=256).astype(np.uint32)


def record(rgb: np.ndarray, quality: int = 90, with_residual: bool = False):
    """Full V5F v1 record for one frame, as a 0-d structured array (same bytes as the C-ABI record)."""
    d, _, _ = ela_residual(rgb, quality)
    rec = np.zeros((), dtype=RECORD_DTYPE)
    for c in range(3):
        ch = d[..., c].ravel()
        rec["ela_hist"][c] = np.bincount(ch, minlength=256)
        rec["ela_sum"][c] = ch.sum(dtype=np.uint64)
        rec["ela_sumsq"][c] = (ch.astype(np.uint64) ** 2).sum(dtype=np.uint64)
        rec["ela_max"][c] = ch.max()
    rec["tex_sumabs"], rec["tex_sumsq"], rec["tex_maxabs"] = texture_stats(rgb)
    return (rec, d) if with_residual else rec


def ela_core(rgb: np.ndarray, quality: int = 90):
    """The bench's CPU-baseline unit of work: the reference's ELA core (v5…:66-73) + the §8a statistics."""
    return record(rgb, quality)


def fft_spectrum(gray: np.ndarray) -> np.ndarray:
    """v5_texture_ela.py:84-88: fft2 -> fftshift -> 20*ln(|F|+1) -> cv2.normalize MINMAX to uint8."""
    import cv2

    f = np.fft.fft2(gray)
    fshift = np.fft.fftshift(f)
    ms = 20 * np.log(np.abs(fshift) + 1)
    return cv2.normalize(ms, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U)

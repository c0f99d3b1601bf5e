"""ctypes wrapper over oracle/libv5ela_oracle.so (v5ela_oracle.c). TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .pil_oracle import RECORD_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libv5ela_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in ("v5ela_oracle.c", "v5jpeg_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libv5ela_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.v5o_record_bytes.restype = ctypes.c_size_t
        L.v5o_analyze_frame.restype = ctypes.c_int
        L.v5o_analyze_frame.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                        ctypes.c_void_p, u8p, u8p, u8p, u8p, u8p]
        L.v5o_analyze.restype = ctypes.c_int
        L.v5o_analyze.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                  ctypes.c_int, ctypes.c_void_p, u8p]
        L.v5o_quant_tables.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_uint16)]
        L.v5o_enhance_lut.argtypes = [ctypes.c_int, u8p]
        i16p = ctypes.POINTER(ctypes.c_int16)
        L.v5jo_encode.restype = ctypes.c_int64
        L.v5jo_encode.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, u8p,
                                  ctypes.c_size_t, i16p]
        L.v5jo_info.restype = ctypes.c_int
        L.v5jo_info.argtypes = [u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                ctypes.POINTER(ctypes.c_int)]
        L.v5jo_layout.restype = ctypes.c_int
        L.v5jo_layout.argtypes = [u8p, ctypes.c_size_t] + [ctypes.POINTER(ctypes.c_int)] * 3
        L.v5jo_decode.restype = ctypes.c_int
        L.v5jo_decode.argtypes = [u8p, ctypes.c_size_t, u8p, u8p, i16p]
        assert L.v5o_record_bytes() == RECORD_DTYPE.itemsize
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)) if a is not None else None


def quant_tables(quality: int):
    lu = np.zeros(64, np.uint16)
    ch = np.zeros(64, np.uint16)
    lib().v5o_quant_tables(quality, lu.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                           ch.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    return lu.reshape(8, 8), ch.reshape(8, 8)


def enhance_lut(max_diff: int) -> np.ndarray:
    lut = np.zeros(256, np.uint8)
    lib().v5o_enhance_lut(int(max_diff), _u8(lut))
    return lut


def analyze_frame(rgb: np.ndarray, quality: int = 90, planes: bool = False):
    """-> dict(record, residual, recon[, y, cb, cr]) for one (H, W, 3) uint8 frame."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    rec = np.zeros((), dtype=RECORD_DTYPE)
    residual = np.empty((h, w, 3), np.uint8)
    recon = np.empty((h, w, 3), np.uint8)
    hm, wm = 16 * ((h + 15) // 16), 16 * ((w + 15) // 16)
    y = np.empty((hm, wm), np.uint8) if planes else None
    cb = np.empty((hm // 2, wm // 2), np.uint8) if planes else None
    cr = np.empty((hm // 2, wm // 2), np.uint8) if planes else None
    rc = lib().v5o_analyze_frame(_u8(rgb), h, w, 3 * w, quality, rec.ctypes.data_as(ctypes.c_void_p),
                                 _u8(residual), _u8(recon), _u8(y), _u8(cb), _u8(cr))
    if rc != 0:
        raise RuntimeError(f"v5o_analyze_frame failed: {rc}")
    out = {"record": rec, "residual": residual, "recon": recon}
    if planes:
        out.update(y=y, cb=cb, cr=cr)
    return out


def analyze(frames: np.ndarray, quality: int = 90, want_residual: bool = False):
    """(N, H, W, 3) uint8 -> (records[N], residual[N,H,W,3] | None)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, h, w, _ = frames.shape
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    residual = np.empty((n, h, w, 3), np.uint8) if want_residual else None
    rc = lib().v5o_analyze(_u8(frames), n, h, w, h * w * 3, w * 3, quality,
                           recs.ctypes.data_as(ctypes.c_void_p), _u8(residual))
    if rc != 0:
        raise RuntimeError(f"v5o_analyze failed: {rc}")
    return recs, residual


# ------------------------------------------------------------------------------------ codec restatement (v5jpeg_oracle.c)
def jpeg_blocks(h: int, w: int, channels: int) -> int:
    """Blocks in the scan (dummy blocks included): 6 per 16x16 MCU for 4:2:0, 1 per 8x8 for one component."""
    return ((h + 15) // 16) * ((w + 15) // 16) * 6 if channels == 3 else ((h + 7) // 8) * ((w + 7) // 8)


def jpeg_encode(img: np.ndarray, quality: int, want_coef: bool = False):
    """(H, W) gray or (H, W, 3) RGB uint8 -> the bytes libjpeg writes (PIL Image.save / cv2.imwrite); optionally also the
    quantised coefficients [blocks, 64] in scan order, zigzag order inside a block."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    ch = 1 if img.ndim == 2 else 3
    h, w = img.shape[:2]
    coef = np.zeros((jpeg_blocks(h, w, ch), 64), np.int16) if want_coef else None
    cap = 4 * h * w * ch + 4096
    out = np.empty(cap, np.uint8)
    n = lib().v5jo_encode(_u8(img), h, w, ch, w * ch, quality, _u8(out), cap,
                          coef.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)) if want_coef else None)
    if n < 0 or n > cap:
        raise RuntimeError(f"v5jo_encode failed: {n}")
    data = out[:n].tobytes()
    return (data, coef) if want_coef else data


def jpeg_info(data: bytes):
    buf = np.frombuffer(data, np.uint8)
    h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().v5jo_info(_u8(buf), len(data), ctypes.byref(h), ctypes.byref(w), ctypes.byref(c))
    if rc != 0:
        raise ValueError(f"v5jo_info: {rc}")
    return h.value, w.value, c.value


def jpeg_layout(data: bytes):
    """-> (hs, vs, restart_interval): luma blocks per MCU across / down, MCUs per restart interval (0 = none)."""
    buf = np.frombuffer(data, np.uint8)
    hs, vs, ri = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().v5jo_layout(_u8(buf), len(data), ctypes.byref(hs), ctypes.byref(vs), ctypes.byref(ri))
    if rc != 0:
        raise ValueError(f"v5jo_layout: {rc}")
    return hs.value, vs.value, ri.value


def jpeg_decode(data: bytes, want_coef: bool = False):
    """-> dict(rgb (H,W,3), gray (H,W)[, coef]) decoded like libjpeg's defaults (ISLOW, fancy upsampling)."""
    h, w, c = jpeg_info(data)
    buf = np.frombuffer(data, np.uint8)
    rgb = np.empty((h, w, 3), np.uint8)
    gray = np.empty((h, w), np.uint8)
    hs, vs, _ = jpeg_layout(data)
    mcus = ((h + 8 * vs - 1) // (8 * vs)) * ((w + 8 * hs - 1) // (8 * hs))
    coef = np.zeros((mcus * (hs * vs + 2 if c == 3 else 1), 64), np.int16) if want_coef else None
    rc = lib().v5jo_decode(_u8(buf), len(data), _u8(rgb), _u8(gray),
                           coef.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)) if want_coef else None)
    if rc != 0:
        raise ValueError(f"v5jo_decode: {rc}")
    out = {"rgb": rgb, "gray": gray, "channels": c}
    if want_coef:
        out["coef"] = coef
    return out

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
    src = os.path.join(_HERE, "v5ela_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
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

"""CPU: the CUDA kernel's device code (csrc/*.cuh), compiled by g++ with threads emulated per barrier phase
(tests/emu/v5ela_emu.cpp), against the oracle. Debug aid for band/halo/edge indexing in a container without a GPU;
the GPU parity tests (-m gpu) are the real gate."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import HERE, golden_frame, load_json, record_matches_golden, sha
from oracle import c_oracle
from oracle.pil_oracle import RECORD_DTYPE
from v5ela.synth import gen_frame

ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module", params=["2cta", "3cta", "2cta_rows_split"])
def emu(request):
    """The default layout (2 CTAs/SM, double-buffered RGB), the -DV5_MIN_CTAS=3 single-buffer layout, and the default layout
    with the two compile-time variants of the width-multiple-of-16 instantiation flipped (one row per residual unit instead
    of two, split barrier on)."""
    variant = request.param
    so = os.path.join(HERE, "emu", f"libv5ela_emu_{variant}.so")
    src = os.path.join(HERE, "emu", "v5ela_emu.cpp")
    csrc = os.path.join(ROOT, "fake-video-detection-engine_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("v5ela_device.cuh", "v5ela_workitem.cuh", "v5ela_host.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        extra = ["-DV5_PAIR_ROWS=0", "-DV5_SPLIT_BARRIER=1"] if variant.endswith("rows_split") else []
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", f"-DV5_MIN_CTAS={variant[0]}", *extra,
                               "-I", os.path.join(ROOT, "include"), "-I", csrc, src, "-o", so])
    lib = ctypes.CDLL(so)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    lib.v5emu_analyze.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                  ctypes.c_int, ctypes.c_void_p, u8p, ctypes.c_int, ctypes.c_void_p]

    def run(frames, q, seg=0, want_residual=True, tex_hist=None):
        frames = np.ascontiguousarray(frames)
        n, h, w, _ = frames.shape
        recs = np.zeros(n, RECORD_DTYPE)
        res = np.zeros((n, h, w, 3), np.uint8) if want_residual else None
        rc = lib.v5emu_analyze(frames.ctypes.data_as(u8p), n, h, w, h * w * 3, w * 3, q,
                               recs.ctypes.data_as(ctypes.c_void_p), res.ctypes.data_as(u8p) if want_residual else None, seg,
                               tex_hist.ctypes.data_as(ctypes.c_void_p) if tex_hist is not None else None)
        assert rc == 0
        return recs, res

    run.lib = lib
    return run


def test_exact_division_constants(emu):
    assert emu.lib.v5emu_quant_selftest() == 0          # every table entry 1..255 x every coefficient -8192..8192


@pytest.mark.parametrize("hw", [(1, 1), (2, 3), (9, 1), (16, 16), (17, 33), (31, 9), (33, 497), (100, 1000), (272, 496)])
@pytest.mark.parametrize("seg", [0, 1, 3])
def test_emulated_kernel_vs_oracle(emu, hw, seg):
    h, w = hw
    rng = np.random.default_rng(h * 7 + w)
    for q in (90, 25):
        for frame in (gen_frame(3, h, w, 1), rng.integers(0, 256, (h, w, 3), dtype=np.uint8)):
            o = c_oracle.analyze_frame(frame, q)
            recs, res = emu(frame[None], q, seg)
            assert np.array_equal(res[0], o["residual"])
            assert recs[0].tobytes() == o["record"].tobytes()


@pytest.mark.parametrize("hw", [(1, 16), (2, 16), (15, 32), (16, 16), (17, 32), (31, 16), (32, 48), (33, 496), (40, 48), (100, 1008),
                                (272, 512), (9, 976)])
@pytest.mark.parametrize("seg", [0, 1, 2])
def test_emulated_fast_instantiation_vs_oracle(emu, hw, seg):
    """Statistics-only calls on widths that are a multiple of 16 take the kernel instantiation without the horizontal edge
    predicates (csrc/v5ela_host.h fast_path_ok); same records as the oracle and as the general instantiation."""
    h, w = hw
    rng = np.random.default_rng(h * 11 + w)
    for q in (90, 40):
        for frame in (gen_frame(5, h, w, 2), rng.integers(0, 256, (h, w, 3), dtype=np.uint8)):
            o = c_oracle.analyze_frame(frame, q)
            recs, _ = emu(frame[None], q, seg, want_residual=False)
            assert recs[0].tobytes() == o["record"].tobytes()
            general, _ = emu(frame[None], q, seg)
            assert general[0].tobytes() == recs[0].tobytes()


@pytest.mark.parametrize("hw", [(1, 1), (1, 7), (9, 1), (16, 16), (17, 33), (40, 48), (33, 497), (100, 1000)])
def test_emulated_texture_histogram(emu, hw):
    """The optional tex_hist[256] output against OpenCV's Laplacian of Pillow's luma; the record and the
    residual map of the same call stay what they are without it."""
    from oracle import pil_oracle

    h, w = hw
    rng = np.random.default_rng(h * 13 + w)
    for frame in (gen_frame(2, h, w, 3), rng.integers(0, 256, (h, w, 3), dtype=np.uint8),
                  np.where(rng.integers(0, 2, (h, w, 1)) > 0, 255, 0).astype(np.uint8).repeat(3, axis=2)):   # saturates the last bin
        for want_residual in (False, True):
            th = np.full((1, 256), 0xA5A5A5A5, np.uint32)
            recs, res = emu(frame[None], 90, 0, want_residual=want_residual, tex_hist=th)
            assert np.array_equal(th[0], pil_oracle.texture_hist(frame))
            assert int(th[0].sum()) == h * w
            o = c_oracle.analyze_frame(frame, 90)
            assert recs[0].tobytes() == o["record"].tobytes()
            if want_residual:
                assert np.array_equal(res[0], o["residual"])


def test_emulated_texture_histogram_goldens(emu):
    """tests/golden/layouts_golden.json: tex_hist recorded from OpenCV's Laplacian of Pillow's luma."""
    for case in load_json("layouts_golden.json")["cases"]:
        th = np.zeros((1, 256), np.uint32)
        emu(golden_frame(case)[None], 90, 0, want_residual=False, tex_hist=th)
        assert sha(th[0]) == case["tex_hist_sha"] and [int(v) for v in th[0][:8]] == case["tex_hist_head"]
        assert int(th[0][255]) == case["tex_hist_last"]


def test_emulated_kernel_vs_reference_goldens(emu):
    cases = [c for c in load_json("frames_golden.json")["cases"] if c["h"] * c["w"] <= 64 * 96]
    assert len(cases) > 20
    for case in cases:
        recs, res = emu(golden_frame(case)[None], case["q"])
        assert sha(res[0]) == case["resid_sha"]
        assert record_matches_golden(recs[0], case) == []

"""GPU: the FFT log-magnitude spectrum image (reference v5_texture_ela.py:84-88) through the C ABI against
NumPy float64 fft2 + cv2.normalize (oracle/pil_oracle.fft_spectrum). Tolerance: 1 grey level (float64 sums are evaluated
in a different order than pocketfft); the fraction of exactly equal pixels is asserted to stay above 99.9 %."""
import os

import cv2
import numpy as np
import pytest

from helpers import GOLDEN, load_json, sha
from oracle import pil_oracle
from v5ela.synth import gen_frame

pytestmark = pytest.mark.gpu


def gpu_spectrum(gray):
    from v5ela import host

    return host.spectrum_host(gray)


def check(gray, min_exact=0.999):
    ref = pil_oracle.fft_spectrum(gray)
    got = gpu_spectrum(gray)
    assert got.shape == ref.shape and got.dtype == np.uint8
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert int(d.max()) <= 1, f"max diff {d.max()}"
    exact = float((d == 0).mean())
    assert exact >= min_exact, exact
    return exact


@pytest.mark.parametrize("hw", [(100, 100), (257, 301), (211, 173), (120, 160), (96, 128), (64, 64), (33, 47), (2, 3), (1, 9), (9, 1), (127, 255)])
def test_crop_sizes_vs_numpy(hw):
    h, w = hw
    gray = pil_oracle.luma(gen_frame(3, h, w, 2))
    check(gray, min_exact=0.99 if h * w < 1000 else 0.999)


def test_reference_fixture_crop_is_exact():
    """The reference's own test crop: decoded Y plane as cv2.imread(IMREAD_GRAYSCALE) gives it (v5_texture_ela.py:83)."""
    gray = cv2.imread(os.path.join(GOLDEN, "node_case0_ref_fixture", "face_000000_0.jpg"), cv2.IMREAD_GRAYSCALE)
    gold = load_json("node_golden.json")["cases"]["case0_ref_fixture"]["ranks"][0]
    assert sha(gray) == gold["gray_sha"]
    got = gpu_spectrum(gray)
    ref = pil_oracle.fft_spectrum(gray)
    assert int(np.abs(got.astype(np.int16) - ref.astype(np.int16)).max()) <= 1
    # the reference's fft_0.jpg decodes to within JPEG loss of this image; the exact artefact comparison is in test_node.py
    ref_file = cv2.imread(os.path.join(GOLDEN, "node_case0_ref_fixture", "fft_0.jpg"), cv2.IMREAD_GRAYSCALE)
    assert ref_file.shape == got.shape


def test_noise_and_flat_inputs():
    rng = np.random.default_rng(5)
    check(rng.integers(0, 256, (180, 320), dtype=np.uint8))
    flat = np.full((40, 56), 77, np.uint8)                      # spectrum = one peak: min == max except DC
    ref = pil_oracle.fft_spectrum(flat)
    got = gpu_spectrum(flat)
    assert int(np.abs(got.astype(np.int16) - ref.astype(np.int16)).max()) <= 1
    zero = np.zeros((16, 24), np.uint8)                          # max == min: cv2.normalize gives all zeros
    assert np.array_equal(gpu_spectrum(zero), pil_oracle.fft_spectrum(zero))


def test_720p_luma_batch():
    import torch
    import v5ela

    frames = np.stack([pil_oracle.luma(gen_frame(i, 720, 1280, 0)) for i in range(2)])
    out = v5ela.spectrum_batch(torch.from_numpy(frames).cuda()).cpu().numpy()
    for i in range(2):
        ref = pil_oracle.fft_spectrum(frames[i])
        d = np.abs(out[i].astype(np.int16) - ref.astype(np.int16))
        assert int(d.max()) <= 1 and float((d == 0).mean()) > 0.999


@pytest.mark.parametrize("hw", [(1080, 1920), (360, 640), (150, 90), (45, 2), (1, 30), (486, 250)])
def test_fft_path_equals_dft_path_within_one_level(hw):
    """Sizes made of 2, 3 and 5 take the shared-memory Stockham FFT; V5ELA_FFT=0 forces the exact-size DFT products."""
    h, w = hw
    rng = np.random.default_rng(h + w)
    gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
    fast = gpu_spectrum(gray)
    os.environ["V5ELA_FFT"] = "0"
    try:
        slow = gpu_spectrum(gray)
    finally:
        del os.environ["V5ELA_FFT"]
    d = np.abs(fast.astype(np.int16) - slow.astype(np.int16))
    assert int(d.max()) <= 1 and float((d == 0).mean()) > 0.999
    if h * w <= 640 * 360:
        check(gray)

"""CPU: pins the codec restatement (oracle/v5jpeg_oracle.c) against Pillow/OpenCV goldens, the reference node's own
artefact files, and the live libraries (SURVEY.md §8f-2 decode, §8f-3 artefact encoders)."""
import glob
import hashlib
import io
import os

import cv2
import numpy as np
import pytest
from PIL import Image

from helpers import check_layout_goldens, GOLDEN, golden_frame, load_json, sha
from oracle import c_oracle

JPEG = load_json("jpeg_golden.json")
SMALL = [c for c in JPEG["cases"] if c["h"] * c["w"] <= 300 * 500]
BIG = [c for c in JPEG["cases"] if c["h"] * c["w"] > 300 * 500]
_id = lambda c: f"{c['spec'][0]}{c['spec'][1]}_{c['h']}x{c['w']}_q{c['q']}"  # noqa: E731


def _file_sha(b):
    return hashlib.sha256(b).hexdigest()[:16]


def pil_file(rgb, q):
    buf = io.BytesIO()
    Image.fromarray(rgb, "RGB").save(buf, "JPEG", quality=q)
    return buf.getvalue()


@pytest.mark.parametrize("case", SMALL + BIG[:1], ids=_id)
def test_encoder_matches_golden_files(case):
    rgb = golden_frame(case)
    assert sha(rgb) == case["in_sha"]
    data = c_oracle.jpeg_encode(rgb, case["q"])
    assert (len(data), _file_sha(data)) == (case["rgb_file_len"], case["rgb_file_sha"])
    gray = np.ascontiguousarray(rgb[..., 1])
    data = c_oracle.jpeg_encode(gray, case["q"])
    assert (len(data), _file_sha(data)) == (case["gray_file_len"], case["gray_file_sha"])


@pytest.mark.parametrize("case", SMALL + BIG[:1], ids=_id)
def test_decoder_matches_golden_pixels(case):
    rgb = golden_frame(case)
    out = c_oracle.jpeg_decode(pil_file(rgb, case["q"]))
    assert sha(out["rgb"]) == case["dec_rgb_sha"]
    assert sha(out["gray"]) == case["dec_y_sha"]
    ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(rgb[..., 1]), [cv2.IMWRITE_JPEG_QUALITY, case["q"]])
    out = c_oracle.jpeg_decode(enc.tobytes())
    assert out["channels"] == 1 and sha(out["gray"]) == case["dec_gray_file_sha"]
    assert np.array_equal(out["rgb"], np.repeat(out["gray"][..., None], 3, axis=2))


def test_reference_node_artefacts_round_trip():
    """The files the unmodified reference node wrote: decode == Pillow/OpenCV, and re-encoding what they were made from
    reproduces them byte for byte (temp_ela = save(crop, q90) v5…:66-67; ela = save(enhanced) v5…:80-81)."""
    d = os.path.join(GOLDEN, "node_case0_ref_fixture")
    files = sorted(glob.glob(os.path.join(d, "*.jpg")))
    assert len(files) == 4
    for f in files:
        data = open(f, "rb").read()
        out = c_oracle.jpeg_decode(data)
        assert np.array_equal(out["rgb"], np.asarray(Image.open(f).convert("RGB"))), f
        assert np.array_equal(out["gray"], cv2.imread(f, cv2.IMREAD_GRAYSCALE)), f
    crop = c_oracle.jpeg_decode(open(os.path.join(d, "face_000000_0.jpg"), "rb").read())["rgb"]
    assert c_oracle.jpeg_encode(crop, 90) == open(os.path.join(d, "temp_ela_0.jpg"), "rb").read()
    from oracle import pil_oracle

    assert c_oracle.jpeg_encode(pil_oracle.ela_enhanced(crop, 90), 75) == open(os.path.join(d, "ela_0.jpg"), "rb").read()
    gray = cv2.imread(os.path.join(d, "face_000000_0.jpg"), cv2.IMREAD_GRAYSCALE)
    assert c_oracle.jpeg_encode(pil_oracle.fft_spectrum(gray), 95) == open(os.path.join(d, "fft_0.jpg"), "rb").read()


def test_live_fuzz_against_pillow_and_opencv():
    rng = np.random.default_rng(5)
    for it in range(60):
        h, w = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        q = int(rng.choice([1, 10, 50, 75, 90, 95, 100]))
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if it % 3 else np.full((h, w, 3), int(rng.integers(256)), np.uint8)
        ref = pil_file(rgb, q)
        assert c_oracle.jpeg_encode(rgb, q) == ref, (h, w, q)
        out = c_oracle.jpeg_decode(ref)
        assert np.array_equal(out["rgb"], np.asarray(Image.open(io.BytesIO(ref)).convert("RGB"))), (h, w, q)
        assert np.array_equal(out["gray"], cv2.imdecode(np.frombuffer(ref, np.uint8), cv2.IMREAD_GRAYSCALE)), (h, w, q)


def test_decoder_takes_custom_tables_and_restart_markers():
    rgb = golden_frame({"spec": ["gen", 1, 2], "h": 123, "w": 211})
    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, "JPEG", quality=80, optimize=True)       # per-image Huffman tables
    assert np.array_equal(c_oracle.jpeg_decode(buf.getvalue())["rgb"], np.asarray(Image.open(buf).convert("RGB")))
    ok, enc = cv2.imencode(".jpg", rgb, [cv2.IMWRITE_JPEG_RST_INTERVAL, 3])
    ref = cv2.cvtColor(cv2.imdecode(enc, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    assert np.array_equal(c_oracle.jpeg_decode(enc.tobytes())["rgb"][..., ::-1], ref[..., ::-1])


def test_decoder_takes_the_other_chroma_layouts_and_restart_intervals():
    """4:4:4 and 4:2:2 files (Pillow subsampling=0 / 1; libjpeg's h2v1 fancy upsampler, jdsample.c), with and without
    restart intervals, decode to the pixels Pillow / OpenCV produce."""
    rng = np.random.default_rng(11)
    for it in range(45):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 110))
        q = int(rng.choice([5, 50, 75, 90, 95, 100]))
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if it % 3 == 0:
            rgb = cv2.GaussianBlur(rgb, (0, 0), 2.5)
        for ss in (0, 1, 2):
            kw = {"restart_marker_blocks": int(rng.integers(1, 7))} if it % 2 else {}
            buf = io.BytesIO()
            Image.fromarray(rgb).save(buf, "JPEG", quality=q, subsampling=ss, **kw)
            data = buf.getvalue()
            hs, vs, ri = c_oracle.jpeg_layout(data)
            assert (hs, vs) == ((1, 1), (2, 1), (2, 2))[ss] and (ri > 0) == bool(kw)
            out = c_oracle.jpeg_decode(data, want_coef=True)
            assert np.array_equal(out["rgb"], np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))), (h, w, q, ss)
            assert np.array_equal(out["gray"], cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)), (h, w, q, ss)


def test_layout_goldens():
    """tests/golden/layouts_golden.json: 100 files in the layouts the decoder row was widened to (4:4:4 / 4:2:2 / 4:2:0 with
    restart intervals, one component with restarts), pixels as recorded from Pillow / OpenCV."""
    check_layout_goldens(lambda blobs: [c_oracle.jpeg_decode(b) for b in blobs])


def test_unsupported_flavours_are_reported():
    rgb = golden_frame({"spec": ["gen", 1, 2], "h": 40, "w": 40})
    for kw in ({"progressive": True},):
        buf = io.BytesIO()
        Image.fromarray(rgb).save(buf, "JPEG", quality=80, **kw)
        with pytest.raises(ValueError):
            c_oracle.jpeg_decode(buf.getvalue())
    with pytest.raises(ValueError):
        c_oracle.jpeg_decode(b"\xff\xd8\xff\xd9")

"""CPU: pins the oracles (oracle/) against the golden vectors produced by executing the reference, and each other."""
import numpy as np
import pytest

from helpers import golden_frame, load_json, record_matches_golden, sha
from oracle import c_oracle, pil_oracle
from v5ela.synth import gen_frame

FRAMES = load_json("frames_golden.json")
SMALL = [c for c in FRAMES["cases"] if c["h"] * c["w"] <= 1280 * 720]
BIG = [c for c in FRAMES["cases"] if c["h"] * c["w"] > 1280 * 720]


def test_library_versions_match_golden():
    # the goldens hash PIL's output; a different libjpeg-turbo build could legitimately differ
    assert pil_oracle.versions()["libjpeg_turbo"].split(".")[0] == FRAMES["versions"]["libjpeg_turbo"].split(".")[0]


def test_generator_known_answer():
    assert sha(gen_frame(0, 720, 1280, 0)) == "4e73222f1a84a534"
    assert sha(gen_frame(1, 211, 173, 0)) == "cb61f10882d9df95"


@pytest.mark.parametrize("case", SMALL, ids=lambda c: f"{c['spec'][0]}{c['spec'][1]}_{c['h']}x{c['w']}_q{c['q']}")
def test_c_oracle_matches_reference_golden(case):
    frame = golden_frame(case)
    assert sha(frame) == case["in_sha"]
    out = c_oracle.analyze_frame(frame, case["q"])
    assert sha(out["residual"]) == case["resid_sha"]
    assert record_matches_golden(out["record"], case) == []


@pytest.mark.parametrize("case", BIG[:3], ids=lambda c: f"{c['h']}x{c['w']}_q{c['q']}")
def test_c_oracle_big_frames(case):
    out = c_oracle.analyze_frame(golden_frame(case), case["q"])
    assert sha(out["residual"]) == case["resid_sha"]
    assert record_matches_golden(out["record"], case) == []


@pytest.mark.parametrize("case", SMALL[:12], ids=lambda c: f"{c['h']}x{c['w']}_q{c['q']}")
def test_pil_oracle_matches_golden(case):
    rec, resid = pil_oracle.record(golden_frame(case), case["q"], with_residual=True)
    assert sha(resid) == case["resid_sha"]
    assert record_matches_golden(rec, case) == []
    assert sha(pil_oracle.ela_enhanced(golden_frame(case), case["q"])) == case["enhanced_sha"]


def test_stored_residual_maps():
    blobs = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "resid_small.npz"))
    n = 0
    for case in FRAMES["cases"]:
        if "resid_key" not in case:
            continue
        out = c_oracle.analyze_frame(golden_frame(case), case["q"])
        assert np.array_equal(out["residual"], blobs[case["resid_key"]]), case["resid_key"]
        n += 1
    assert n >= 30


@pytest.mark.parametrize("q", [1, 25, 50, 75, 85, 90, 95, 100])
def test_quant_tables_match_pillow(q):
    import io

    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(gen_frame(0, 16, 16, 0), "RGB").save(buf, "JPEG", quality=q)
    buf.seek(0)
    qt = Image.open(buf).quantization
    lu, ch = c_oracle.quant_tables(q)
    # Pillow >= 11 reports tables in natural order
    assert list(qt[0]) == lu.ravel().tolist()
    assert list(qt[1]) == ch.ravel().tolist()


def test_enhance_lut_is_float32():
    # Pillow blends in float32 -> 7 * f32(255/7) truncates to 254; float64 arithmetic would give 255
    assert c_oracle.enhance_lut(7)[7] == 254
    from PIL import Image, ImageEnhance

    for m in range(1, 256):
        ramp = np.arange(256, dtype=np.uint8).reshape(16, 16)
        img = Image.fromarray(np.stack([ramp] * 3, -1), "RGB")
        ref = np.asarray(ImageEnhance.Brightness(img).enhance(255.0 / m))[..., 0].ravel()
        assert np.array_equal(c_oracle.enhance_lut(m), ref), m


def test_oracles_agree_on_random_small_frames():
    rng = np.random.default_rng(7)
    for _ in range(40):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        q = int(rng.choice([1, 20, 50, 75, 90, 95, 100]))
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if rng.random() < 0.5 else gen_frame(int(rng.integers(0, 99)), h, w, 3)
        rec, resid = pil_oracle.record(a, q, with_residual=True)
        out = c_oracle.analyze_frame(a, q)
        assert np.array_equal(out["residual"], resid), (h, w, q)
        assert out["record"].tobytes() == rec.tobytes(), (h, w, q)

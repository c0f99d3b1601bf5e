#!/usr/bin/env python
"""Generate the committed golden vectors by EXECUTING THE REFERENCE in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only) — not available on the GPU box

Two families, both written next to this script:

* ``node_golden.json`` (+ ``node_case*/`` small JPEG artefacts): the UNMODIFIED reference node
  (``/root/reference/nodes/V_nodes/v5_texture_ela.py``, loaded by ``oracle/ref_loader.py``) is run with
  ``OPENAI_API_KEY`` unset on crops written with ``cv2.imwrite`` exactly as V1 does (``v1_keyframes_facetrack.py:166``).
  Recorded per selected face: SHA-256 of the three artefact files the node wrote (``temp_ela_i.jpg``, ``ela_i.jpg``,
  ``fft_i.jpg``), of their decoded pixels, and of the residual ``|original - Image.open(temp_ela_i.jpg)|`` recovered
  from the reference's own temp file — i.e. the exact array ``ImageChops.difference`` produced inside the node
  (``v5…:70``) — plus its per-channel maxima (``v5…:72-73``).
* ``frames_golden.json`` (+ ``resid_*.npz`` for the small cases): the reference's call sequence
  (``v5…:66-73``; same Pillow calls, BytesIO instead of a temp file) on raw synthetic frames from
  ``v5ela.synth.gen_frame`` — SURVEY.md Appendix B's known-answer table regenerated, extended with the §8a record
  fields, for q in {75, 85, 90, 95}, odd sizes, tiny sizes and adversarial content.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fake-video-detection-engine_b200"))

import cv2  # noqa: E402
from PIL import Image  # noqa: E402

from oracle import pil_oracle, ref_loader  # noqa: E402
from v5ela.synth import gen_frame  # noqa: E402


def sha(a) -> str:
    if isinstance(a, np.ndarray):
        a = np.ascontiguousarray(a).tobytes()
    return hashlib.sha256(a).hexdigest()[:16]


def file_sha(path) -> str:
    with open(path, "rb") as f:
        return sha(f.read())


def rec_to_json(rec) -> dict:
    return {
        "ela_max": [int(v) for v in rec["ela_max"]],
        "ela_sum": [int(v) for v in rec["ela_sum"]],
        "ela_sumsq": [int(v) for v in rec["ela_sumsq"]],
        "ela_hist_sha": sha(rec["ela_hist"]),
        "ela_hist_r_head": [int(v) for v in rec["ela_hist"][0][:8]],
        "tex_sumabs": int(rec["tex_sumabs"]),
        "tex_sumsq": int(rec["tex_sumsq"]),
        "tex_maxabs": int(rec["tex_maxabs"]),
        "record_sha": sha(rec.tobytes()),
    }


# ------------------------------------------------------------------------------------------------ node goldens
def reference_fixture_bgr() -> np.ndarray:
    """The array the reference's own test builds (tests/test_v5_texture_ela.py:23-24)."""
    img = np.zeros((100, 100, 3), dtype=np.uint8)
    cv2.rectangle(img, (25, 25), (75, 75), (255, 255, 255), -1)
    return img


NODE_CASES = [
    # name, list of (crop source, confidence, bbox w, bbox h)
    ("case0_ref_fixture", [("fixture", 0.99, 100, 100)]),
    ("case1_top3_of_4", [
        (("gen", 1, 211, 173, 0), 0.90, 140, 170),     # score 21420 -> rank 1
        (("gen", 2, 120, 160, 3), 0.99, 130, 100),     # score 12870 -> rank 2
        (("gen", 3, 96, 128, 5), 0.50, 100, 80),       # score 4000  -> dropped (rank 3)
        (("gen", 4, 257, 301, 7), 0.80, 250, 210),     # score 42000 -> rank 0
    ]),
]


def make_crop(src) -> np.ndarray:
    if src == "fixture":
        return reference_fixture_bgr()
    _, n, h, w, seed = src
    return np.ascontiguousarray(gen_frame(n, h, w, seed)[..., ::-1])   # BGR for cv2.imwrite


def run_node_cases():
    v5 = ref_loader.load_reference_v5()
    os.environ.pop("OPENAI_API_KEY", None)
    out = {"versions": pil_oracle.versions(), "cases": {}}
    for name, faces in NODE_CASES:
        tmp = tempfile.mkdtemp()
        try:
            fdir = os.path.join(tmp, "faces")
            os.makedirs(fdir)
            dets, crops = [], []
            for i, (src, conf, bw, bh) in enumerate(faces):
                p = os.path.join(fdir, f"face_{i:06d}_0.jpg")
                cv2.imwrite(p, make_crop(src))
                crops.append(p)
                dets.append({"frame_id": i, "timestamp": float(i),
                             "faces": [{"bbox": {"x": 0, "y": 0, "w": bw, "h": bh}, "confidence": conf,
                                        "is_main": True, "crop_path": p}]})
            state = {"face_detections": dets, "data_dir": tmp, "debug": False}
            res = v5.run(state)
            order = sorted(range(len(faces)), key=lambda i: faces[i][1] * faces[i][2] * faces[i][3], reverse=True)[:3]
            case = {"score": res["texture_ela_score"], "details": res["texture_ela_details"],
                    "faces": [{"src": f[0], "confidence": f[1], "w": f[2], "h": f[3]} for f in faces],
                    "selected": order, "ranks": []}
            ela_dir = os.path.join(tmp, "ela_analysis")
            keep = os.path.join(HERE, "node_" + name)
            if os.path.isdir(keep):
                shutil.rmtree(keep)
            os.makedirs(keep)
            for rank, idx in enumerate(order):
                original = np.asarray(Image.open(crops[idx]).convert("RGB"))
                t, e, f = (os.path.join(ela_dir, f"{k}_{rank}.jpg") for k in ("temp_ela", "ela", "fft"))
                comp = np.asarray(Image.open(t).convert("RGB"))
                resid = np.abs(original.astype(np.int16) - comp.astype(np.int16)).astype(np.uint8)
                gray = cv2.imread(crops[idx], cv2.IMREAD_GRAYSCALE)
                case["ranks"].append({
                    "face_index": idx,
                    "crop_file_sha": file_sha(crops[idx]),
                    "original_rgb_sha": sha(original),
                    "gray_sha": sha(gray),
                    "temp_ela_file_sha": file_sha(t), "ela_file_sha": file_sha(e), "fft_file_sha": file_sha(f),
                    "temp_ela_pixels_sha": sha(comp),
                    "ela_pixels_sha": sha(np.asarray(Image.open(e).convert("RGB"))),
                    "fft_pixels_sha": sha(cv2.imread(f, cv2.IMREAD_GRAYSCALE)),
                    "residual_sha": sha(resid),
                    "residual_max": [int(resid[..., c].max()) for c in range(3)],
                    "shape": list(original.shape),
                })
                if name == "case0_ref_fixture":      # tiny: keep the reference's artefacts themselves
                    for p in (crops[idx], t, e, f):
                        shutil.copy(p, os.path.join(keep, os.path.basename(p)))
            if not os.listdir(keep):
                os.rmdir(keep)
            out["cases"][name] = case
        finally:
            shutil.rmtree(tmp)
    with open(os.path.join(HERE, "node_golden.json"), "w") as fp:
        json.dump(out, fp, indent=1)
    return out


# ---------------------------------------------------------------------------------------------- frame goldens
def adversarial(kind: str, h: int, w: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "binary":
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    if kind == "checker":
        yy, xx = np.mgrid[0:h, 0:w]
        return np.repeat((((yy + xx) & 1) * 255).astype(np.uint8)[..., None], 3, axis=2)
    if kind == "flat":
        return np.full((h, w, 3), [200, 30, 90], dtype=np.uint8)
    if kind == "saturated":
        a = np.zeros((h, w, 3), np.uint8)
        a[..., 0] = 255 * ((np.arange(w)[None, :] // 3) & 1)
        a[..., 2] = 255 * ((np.arange(h)[:, None] // 5) & 1)
        return a
    raise ValueError(kind)


FRAME_CASES = (
    # (kind/gen args, h, w, qualities, keep full residual?)
    [(("gen", 0, 0), 720, 1280, (90,), False),
     (("gen", 0, 0), 1080, 1920, (75, 85, 90, 95), False),
     (("gen", 5, 0), 2160, 3840, (90,), False),
     (("gen", 1, 0), 211, 173, (75, 90), True),
     (("gen", 7, 2), 64, 96, (90,), True),
     (("gen", 2, 1), 33, 47, (50, 90, 100), True)]
    + [(("gen", 3, 4), h, w, (90,), True) for (h, w) in
       [(1, 1), (1, 9), (9, 1), (2, 2), (3, 5), (5, 4), (8, 8), (16, 16), (17, 17), (15, 33), (31, 9), (40, 3), (6, 40)]]
    + [((k, 11), 48, 80, (1, 30, 90, 100), True) for k in ("noise", "binary", "checker", "flat", "saturated")]
    + [(("noise", 12), 270, 480, (90,), False)]
)


def make_frame(spec, h, w) -> np.ndarray:
    if spec[0] == "gen":
        return gen_frame(spec[1], h, w, spec[2])
    return adversarial(spec[0], h, w, spec[1])


def run_frame_cases():
    out = {"versions": pil_oracle.versions(), "cases": []}
    blobs = {}
    for spec, h, w, quals, keep in FRAME_CASES:
        frame = make_frame(spec, h, w)
        for q in quals:
            rec, resid = pil_oracle.record(frame, q, with_residual=True)
            enh = pil_oracle.ela_enhanced(frame, q)
            entry = {"spec": list(spec), "h": h, "w": w, "q": q, "in_sha": sha(frame), "resid_sha": sha(resid),
                     "enhanced_sha": sha(enh), **rec_to_json(rec)}
            if keep:
                key = f"resid_{'_'.join(str(s) for s in spec)}_{h}x{w}_q{q}"
                blobs[key] = resid
                entry["resid_key"] = key
            out["cases"].append(entry)
    np.savez_compressed(os.path.join(HERE, "resid_small.npz"), **blobs)
    with open(os.path.join(HERE, "frames_golden.json"), "w") as fp:
        json.dump(out, fp, indent=1)
    return out


if __name__ == "__main__":
    n = run_node_cases()
    f = run_frame_cases()
    print("node cases:", list(n["cases"]), "frame cases:", len(f["cases"]), "versions:", f["versions"])

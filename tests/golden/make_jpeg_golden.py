"""Generates tests/golden/jpeg_golden.json: known answers for the codec rows.

Every value is produced by the libraries the reference calls, with the reference's own call forms
(/root/reference/nodes/V_nodes/v5_texture_ela.py and v1_keyframes_facetrack.py):
    encode  RGB : PIL  Image.save(buf, 'JPEG', quality=q)                 v5…:66-67 (q=90), :80-81 (default 75)
    encode  gray: cv2.imwrite / imencode('.jpg', gray[, quality])         v5…:90-91 (default 95)
    encode  crop: cv2.imwrite('.jpg', bgr)                                v1…:166 (default 95)
    decode  RGB : PIL  Image.open(buf).convert('RGB')                     v5…:64
    decode  gray: cv2.imread(path, IMREAD_GRAYSCALE)                      v5…:83
Run in the build container:  python tests/golden/make_jpeg_golden.py
"""
import hashlib
import io
import json
import os
import sys

import cv2
import numpy as np
from PIL import Image, features

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "fake-video-detection-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden_frame, sha  # noqa: E402


def gray_of(case):
    return np.ascontiguousarray(golden_frame(case)[..., 1])


CASES = []
for spec, h, w in [(["gen", 0, 0], 1, 1), (["gen", 1, 0], 7, 9), (["gen", 2, 0], 8, 8), (["gen", 3, 0], 16, 16),
                   (["gen", 4, 0], 17, 33), (["gen", 5, 1], 31, 15), (["noise", 6], 40, 56), (["binary", 7], 23, 41),
                   (["checker", 8], 32, 48), (["flat", 9], 19, 21), (["saturated", 10], 50, 70), (["gen", 11, 2], 211, 173),
                   (["gen", 12, 3], 257, 301), (["noise", 13], 100, 3), (["noise", 14], 3, 100), (["gen", 15, 0], 270, 480),
                   (["gen", 0, 0], 720, 1280), (["gen", 3, 0], 1080, 1920)]:
    for q in ([75, 90, 95] if h * w < 100000 else [90]):
        CASES.append({"spec": spec, "h": h, "w": w, "q": q})
for q in (1, 30, 100):
    CASES.append({"spec": ["noise", 20 + q], "h": 45, "w": 61, "q": q})


def main():
    out = {"versions": {"pillow": Image.__version__, "libjpeg_turbo": features.version_feature("libjpeg_turbo"),
                        "opencv": cv2.__version__}, "cases": []}
    for case in CASES:
        rgb = golden_frame(case)
        gray = gray_of(case)
        q = case["q"]
        buf = io.BytesIO()
        Image.fromarray(rgb, "RGB").save(buf, "JPEG", quality=q)
        rgb_file = buf.getvalue()
        ok, enc = cv2.imencode(".jpg", gray, [cv2.IMWRITE_JPEG_QUALITY, q])
        gray_file = enc.tobytes()
        ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(rgb[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, q])
        assert enc.tobytes() == rgb_file, "PIL and OpenCV disagree on the colour file"
        dec_rgb = np.asarray(Image.open(io.BytesIO(rgb_file)).convert("RGB"))
        dec_y = cv2.imdecode(np.frombuffer(rgb_file, np.uint8), cv2.IMREAD_GRAYSCALE)
        dec_gray_file = cv2.imdecode(np.frombuffer(gray_file, np.uint8), cv2.IMREAD_GRAYSCALE)
        c = dict(case)
        c.update(in_sha=sha(rgb), rgb_file_len=len(rgb_file), rgb_file_sha=hashlib.sha256(rgb_file).hexdigest()[:16],
                 gray_file_len=len(gray_file), gray_file_sha=hashlib.sha256(gray_file).hexdigest()[:16],
                 dec_rgb_sha=sha(dec_rgb), dec_y_sha=sha(dec_y), dec_gray_file_sha=sha(dec_gray_file))
        out["cases"].append(c)
    with open(os.path.join(HERE, "jpeg_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()

"""Generates tests/golden/layouts_golden.json: known answers for what the round's last session widened —
the decoder's other chroma layouts and restart intervals and the optional texture histogram (§8a).

Every value is produced by the libraries the reference calls:
    files        PIL Image.save(buf, 'JPEG', quality=q, subsampling=0|1|2[, restart_marker_blocks=n])   (v5…:66-67 call form)
                 cv2.imencode('.jpg', bgr, [IMWRITE_JPEG_QUALITY, q, IMWRITE_JPEG_RST_INTERVAL, n])     (v1…:166 call form)
    decode RGB   PIL Image.open(buf).convert('RGB')                                                     v5…:64
    decode gray  cv2.imdecode(.., IMREAD_GRAYSCALE)                                                     v5…:83
    tex_hist     np.bincount(np.minimum(np.abs(cv2.Laplacian(PIL convert('L'), CV_16S, ksize=1)), 255)) §8a record table
Run in the build container:  python tests/golden/make_layouts_golden.py
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

FRAMES = [(["gen", 0, 0], 1, 1), (["gen", 1, 0], 7, 9), (["gen", 3, 0], 16, 16), (["gen", 4, 0], 17, 33), (["noise", 6], 40, 56),
          (["binary", 7], 23, 41), (["saturated", 10], 50, 70), (["gen", 12, 3], 257, 301), (["noise", 14], 3, 100),
          (["gen", 15, 0], 270, 480)]


def layout_files(rgb, q):
    """(label, file bytes) for every writer option the decoder claims."""
    out = []
    for ss, name in ((0, "444"), (1, "422"), (2, "420")):
        for kw, tag in (({}, ""), ({"restart_marker_blocks": 3}, "_rst3"), ({"restart_marker_rows": 1}, "_rstrow")):
            if ss == 2 and not kw:
                continue                                    # the default files are jpeg_golden.json's business
            buf = io.BytesIO()
            Image.fromarray(rgb, "RGB").save(buf, "JPEG", quality=q, subsampling=ss, **kw)
            out.append((f"pil_{name}{tag}", buf.getvalue()))
    ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(rgb[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, 5])
    out.append(("cv2_420_rst5", enc.tobytes()))
    ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(rgb[..., 1]), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    out.append(("cv2_gray_rst2", enc.tobytes()))
    return out


def texture_hist(rgb):
    y = np.asarray(Image.fromarray(rgb, "RGB").convert("L"))
    lap = cv2.Laplacian(y, cv2.CV_16S, ksize=1).astype(np.int64)
    return np.bincount(np.minimum(np.abs(lap), 255).ravel(), minlength=256).astype(np.uint32)


def main():
    out = {"versions": {"pillow": Image.__version__, "libjpeg_turbo": features.version_feature("libjpeg_turbo"),
                        "opencv": cv2.__version__}, "cases": []}
    for i, (spec, h, w) in enumerate(FRAMES):
        case = {"spec": spec, "h": h, "w": w, "q": (95, 75, 50, 90)[i % 4]}
        rgb = golden_frame(case)
        files = []
        for label, data in layout_files(rgb, case["q"]):
            dec_rgb = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
            dec_y = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
            files.append({"label": label, "file_len": len(data), "file_sha": hashlib.sha256(data).hexdigest()[:16],
                          "dec_rgb_sha": sha(dec_rgb), "dec_y_sha": sha(dec_y)})
        th = texture_hist(rgb)
        case.update(in_sha=sha(rgb), files=files, tex_hist_sha=sha(th), tex_hist_head=[int(v) for v in th[:8]], tex_hist_last=int(th[255]))
        out["cases"].append(case)
    with open(os.path.join(HERE, "layouts_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "frames,", sum(len(c["files"]) for c in out["cases"]), "files")


if __name__ == "__main__":
    main()

"""CPU: V1's face-crop arithmetic (reference v1_keyframes_facetrack.py:117-162) as restated in v5ela.handoff."""
from v5ela import handoff


def ref_crop(face, fw, fh):
    # the reference's lines, verbatim in meaning (v1_keyframes_facetrack.py:154-160)
    x, y, w, h = face["x"], face["y"], face["w"], face["h"]
    pad_w = int(w * 0.2)
    pad_h = int(h * 0.2)
    return max(0, x - pad_w), max(0, y - pad_h), min(fw, x + w + pad_w), min(fh, y + h + pad_h)


def test_crop_box_rule():
    for fw, fh in ((640, 360), (1920, 1080), (33, 17)):
        for x in (0, 1, fw // 3, fw - 2):
            for y in (0, fh // 2, fh - 1):
                for w in (1, 4, 5, 9, fw // 2, fw - x):
                    for h in (1, 7, fh - y):
                        f = {"x": x, "y": y, "w": w, "h": h}
                        assert handoff.crop_box(f, fw, fh) == ref_crop(f, fw, fh)


def test_face_selection_keeps_v1_indices():
    boxes = [{"x": 0, "y": 0, "w": 10, "h": 10, "confidence": 0.6}, {"x": 0, "y": 0, "w": 300, "h": 300, "confidence": 0.7},
             {"x": 0, "y": 0, "w": 100, "h": 100, "confidence": 0.9}, {"x": 5, "y": 5, "w": 0, "h": 4, "confidence": 0.9}]
    sel = handoff.select_faces(boxes, 1920, 1080)
    # sorted by area (v1…:142); faces under 0.5 % of the frame (v1…:149: 10368 px at 1080p) are skipped; indices = area ranks
    assert [(i, f["w"]) for i, f in sel] == [(0, 300)]
    assert [(i, f["w"]) for i, f in handoff.select_faces(boxes, 640, 360)] == [(0, 300), (1, 100)]

"""V1 -> V5 hand-off on the device.

The reference's V1 node writes every keyframe and every padded face crop to disk as JPEG
(/root/reference/nodes/V_nodes/v1_keyframes_facetrack.py:110-112, :144-166) and V5 reads the crops back (v5_texture_ela.py:64,
:83). When the decoded frames already live on the GPU neither file is needed to compute V5's features: the crop is a strided
view of the frame tensor and goes straight into ``analyze_batch``. The files are still produced for the LLM payload and the
UI — by the GPU encoder, byte-identical to ``cv2.imwrite`` — so the ``face_detections`` state the rest of the graph consumes
is unchanged. The SSD face detector itself stays third-party (boxes are an input here).
"""
from __future__ import annotations

import os


def crop_box(face: dict, frame_w: int, frame_h: int):
    """V1's crop rectangle for a detection {x, y, w, h}: 20 % padding on every side, clamped to the frame
    (v1_keyframes_facetrack.py:154-160). Returns (x1, y1, x2, y2), exclusive ends."""
    x, y, w, h = int(face["x"]), int(face["y"]), int(face["w"]), int(face["h"])
    pad_w, pad_h = int(w * 0.2), int(h * 0.2)
    return max(0, x - pad_w), max(0, y - pad_h), min(frame_w, x + w + pad_w), min(frame_h, y + h + pad_h)


def select_faces(boxes, frame_w: int, frame_h: int):
    """V1's per-frame face list (v1…:117-152): boxes are dicts {x, y, w, h, confidence} already clamped to the frame; sorted by
    area, largest first; faces under 0.5 % of the frame area are dropped but keep their index (it names the crop file)."""
    faces = [dict(b, area=int(b["w"]) * int(b["h"])) for b in boxes if b["w"] > 0 and b["h"] > 0]
    faces.sort(key=lambda f: f["area"], reverse=True)
    return [(i, f) for i, f in enumerate(faces) if f["area"] >= frame_w * frame_h * 0.005]


def face_detections_on_device(frames, frame_ids, boxes_per_frame, data_dir=None, fps: float = 1.0, write_files: bool = True,
                              quality: int = 90):
    """frames: uint8 CUDA tensor (N, H, W, 3), RGB. boxes_per_frame: per frame, the detector's boxes (see select_faces).

    Returns (face_detections, features): ``face_detections`` has V1's structure (v1…:168-180), with ``crop_path`` /
    ``keyframe_path`` pointing at JPEG files encoded on the GPU (cv2.imwrite's bytes; skipped when write_files is False);
    ``features`` maps (frame index, face index) -> the V5F v1 feature dict computed from the crop view on the device.
    NOTE: V5 itself analyses the crop *file* after a JPEG round trip at quality 95; pass that decoded image instead when
    the reference's exact numbers are wanted — ``analyze_jpeg_files`` does that on the GPU.
    """
    import torch

    from . import jpeg
    from .batch import analyze_batch
    from .records import as_records, features

    n, fh, fw, _ = frames.shape
    detections, feats = [], {}
    faces_dir = keyframes_dir = None
    if write_files:
        faces_dir, keyframes_dir = os.path.join(data_dir, "faces"), os.path.join(data_dir, "keyframes")
        os.makedirs(faces_dir, exist_ok=True)
        os.makedirs(keyframes_dir, exist_ok=True)
        kf_files, kf_sizes = jpeg.encode_batch(frames, 95)                  # cv2.imwrite default quality
        kf_files, kf_sizes = kf_files.cpu(), kf_sizes.cpu()
    for k in range(n):
        frame_id = int(frame_ids[k])
        keyframe_path = None
        if write_files:
            keyframe_path = os.path.join(keyframes_dir, f"frame_{frame_id:06d}.jpg")
            with open(keyframe_path, "wb") as fh_:
                fh_.write(kf_files[k, :int(kf_sizes[k])].numpy().tobytes())
        in_frame = []
        for i, face in select_faces(boxes_per_frame[k], fw, fh):
            x1, y1, x2, y2 = crop_box(face, fw, fh)
            view = frames[k:k + 1, y1:y2, x1:x2]
            rec = analyze_batch(view, quality=quality)["records"]
            feats[(k, i)] = features(as_records(rec.cpu())[0], (y2 - y1) * (x2 - x1))
            crop_path = None
            if write_files:
                data, size = jpeg.encode_batch(view, 95)
                crop_path = os.path.join(faces_dir, f"face_{frame_id:06d}_{i}.jpg")
                with open(crop_path, "wb") as fh_:
                    fh_.write(data[0, :int(size[0])].cpu().numpy().tobytes())
            in_frame.append({"bbox": {"x": face["x"], "y": face["y"], "w": face["w"], "h": face["h"]},
                             "confidence": face["confidence"], "is_main": i == 0, "crop_path": crop_path})
        detections.append({"frame_id": frame_id, "timestamp": frame_id / fps if fps else 0.0, "faces": in_frame,
                           "keyframe_path": keyframe_path})
    torch.cuda.synchronize(frames.device)
    return detections, feats

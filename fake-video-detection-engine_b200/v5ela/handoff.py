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


def _encode_checked(images, quality: int):
    """jpeg.encode_batch with the overflow rule enforced: a file that did not fit the default capacity (sizes[i] > capacity, bytes
    undefined) makes the whole group go again with the capacity that always fits (v5ela_jpeg_bound). Returns host (files, sizes)."""
    from . import jpeg

    files, sizes = jpeg.encode_batch(images, quality)
    files_h, sizes_h = files.cpu(), sizes.cpu()
    if bool((sizes_h > files_h.shape[1]).any()):
        ch = 3 if images.dim() == 4 else 1
        files, sizes = jpeg.encode_batch(images, quality, capacity=jpeg.bound(int(images.shape[1]), int(images.shape[2]), ch))
        files_h, sizes_h = files.cpu(), sizes.cpu()
        if bool((sizes_h > files_h.shape[1]).any()):
            raise RuntimeError("JPEG encoder overflowed its bound capacity")
    return files_h, sizes_h


def face_detections_on_device(frames, frame_ids, boxes_per_frame, data_dir=None, fps: float = 1.0, write_files: bool = True,
                              quality: int = 90):
    """frames: uint8 CUDA tensor (N, H, W, 3), RGB. boxes_per_frame: per frame, the detector's boxes (see select_faces).

    Returns (face_detections, features): ``face_detections`` has V1's structure (v1…:168-180), with ``crop_path`` /
    ``keyframe_path`` pointing at JPEG files encoded on the GPU (cv2.imwrite's bytes; skipped when write_files is False);
    ``features`` maps (frame index, face index) -> the V5F v1 feature dict computed from the crop view on the device.
    All crops of the batch — whatever their sizes — are analysed by ONE ragged launch (``analyze_ragged``) and their records come
    back in one copy; the crop files are encoded one group of equally sized crops at a time (the encoder's batches are uniform).
    NOTE: V5 itself analyses the crop *file* after a JPEG round trip at quality 95; pass that decoded image instead when
    the reference's exact numbers are wanted — ``analyze_jpeg_files`` does that on the GPU.
    """
    import torch

    from .batch import analyze_ragged
    from .records import as_records, features

    n, fh, fw, _ = frames.shape
    faces_dir = keyframes_dir = None
    kf_files = kf_sizes = None
    if write_files:
        faces_dir, keyframes_dir = os.path.join(data_dir, "faces"), os.path.join(data_dir, "keyframes")
        os.makedirs(faces_dir, exist_ok=True)
        os.makedirs(keyframes_dir, exist_ok=True)
        kf_files, kf_sizes = _encode_checked(frames, 95)                    # cv2.imwrite default quality
    # every crop of the batch: (frame index, face index, face, box, view)
    crops = []
    for k in range(n):
        for i, face in select_faces(boxes_per_frame[k], fw, fh):
            x1, y1, x2, y2 = crop_box(face, fw, fh)
            if x2 > x1 and y2 > y1:
                crops.append((k, i, face, (x1, y1, x2, y2), frames[k, y1:y2, x1:x2]))
    feats = {}
    if crops:
        out = analyze_ragged([c[4] for c in crops], quality=quality)       # one launch, one download
        recs = as_records(out["records"].cpu())
        for j, (k, i, _, (x1, y1, x2, y2), _) in enumerate(crops):
            feats[(k, i)] = features(recs[j], (y2 - y1) * (x2 - x1))
    crop_bytes = {}
    if write_files and crops:
        groups = {}
        for j, c in enumerate(crops):
            groups.setdefault((c[4].shape[0], c[4].shape[1]), []).append(j)
        for (_, _), members in groups.items():
            batch = torch.stack([crops[j][4] for j in members])           # equally sized crops: one encoder batch
            files_h, sizes_h = _encode_checked(batch, 95)
            for m, j in enumerate(members):
                crop_bytes[j] = files_h[m, :int(sizes_h[m])].numpy().tobytes()
    detections = []
    by_frame = {}
    for j, c in enumerate(crops):
        by_frame.setdefault(c[0], []).append(j)
    for k in range(n):
        frame_id = int(frame_ids[k])
        keyframe_path = None
        if write_files:
            keyframe_path = os.path.join(keyframes_dir, f"frame_{frame_id:06d}.jpg")
            with open(keyframe_path, "wb") as fh_:
                fh_.write(kf_files[k, :int(kf_sizes[k])].numpy().tobytes())
        in_frame = []
        for j in by_frame.get(k, []):
            _, i, face, _, _ = crops[j]
            crop_path = None
            if write_files:
                crop_path = os.path.join(faces_dir, f"face_{frame_id:06d}_{i}.jpg")
                with open(crop_path, "wb") as fh_:
                    fh_.write(crop_bytes[j])
            in_frame.append({"bbox": {"x": face["x"], "y": face["y"], "w": face["w"], "h": face["h"]},
                             "confidence": face["confidence"], "is_main": i == 0, "crop_path": crop_path})
        detections.append({"frame_id": frame_id, "timestamp": frame_id / fps if fps else 0.0, "faces": in_frame,
                           "keyframe_path": keyframe_path})
    return detections, feats

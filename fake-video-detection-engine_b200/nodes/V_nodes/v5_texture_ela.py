"""Drop-in replacement for the reference node ``nodes/V_nodes/v5_texture_ela.py`` (same module path, same ``run``).

What is unchanged (reference line numbers refer to /root/reference/nodes/V_nodes/v5_texture_ela.py):
  * signature ``run(state: dict) -> dict``; reads ``face_detections``, ``debug``, ``data_dir`` (:16-18) and
    ``OPENAI_API_KEY`` (:49); writes ``texture_ela_score`` / ``texture_ela_details`` (:22-23, :29-30, :163-164, :176-177)
    with the same reason strings; top-3 selection by confidence*w*h of faces[0] (:33-42); artefacts
    ``ela_analysis/ela_{i}.jpg`` and ``fft_{i}.jpg`` named by selection rank (:80, :90); the GPT-4o request (:93-138);
    the score aggregation and ``V5_debug.json`` payload keys (:147-173); per-face errors are printed and swallowed (:140-144).
What is replaced: the error-level analysis itself (:66-78: save q=90 -> reopen -> ImageChops.difference -> getextrema ->
  Brightness.enhance) runs on the GPU through libv5ela.so (include/v5ela.h: v5ela_analyze_host) and is bit-exact, so
  ``ela_{i}.jpg`` is byte-identical to the reference's. There is no CPU fallback: if the library or a B200 is missing the
  per-face ``try`` reports the error exactly like any other analysis failure.
  The FFT log-magnitude spectrum image (:84-88) also runs on the GPU (v5ela_spectrum_host: float64 DFT + log + min/max
  normalisation); it is within one grey level of NumPy/OpenCV and bit-identical on every crop of the golden set, so
  ``fft_{i}.jpg`` matches the reference's file there too. ``state["v5_gpu_fft"] = False`` selects the reference's host code.
What is added (optional, defaults reproduce the reference): state keys ``v5_quality`` (90), ``v5_max_faces`` (3),
  ``v5_device`` (0), ``v5_gpu_fft`` (True), ``v5_keep_temp_jpeg`` (False: the reference's ``temp_ela_{i}.jpg`` scratch file
  is only written on request since nothing reads it); the per-face integer/float statistics of the V5F v1 record are attached as
  ``ela_features`` inside ``texture_ela_details`` entries and ``V5_debug.json`` (lr_node reads only ``avg_score``).
Still on the host, as in the reference: decoding the crop (:64, :83), JPEG-encoding the two artefacts (:81, :91), and
  the OpenAI call.
"""
import base64
import json
import os

import cv2
import numpy as np
from dotenv import load_dotenv
from openai import OpenAI
from PIL import Image

from nodes import dump_node_debug

load_dotenv()


def _analyze_crop(rgb: np.ndarray, quality: int, device: int):
    """GPU error-level analysis of one RGB crop -> (features dict, enhanced residual image HxWx3 uint8)."""
    from v5ela import host as v5host
    from v5ela.records import features

    recs, _, enhanced = v5host.analyze_frames_host(rgb[None], quality=quality, want_enhanced=True, device=device)
    return features(recs[0], rgb.shape[0] * rgb.shape[1]), enhanced[0]


def run(state: dict) -> dict:
    print("Node V5: Running Texture & ELA Analysis...")

    face_detections = state.get("face_detections", [])
    debug = state.get("debug", False)
    output_dir = state.get("data_dir")
    quality = int(state.get("v5_quality", 90))
    max_faces = int(state.get("v5_max_faces", 3))
    device = int(state.get("v5_device", 0))

    if not face_detections:
        print("Node V5: No faces detected to analyze.")
        state["texture_ela_score"] = 0.0
        state["texture_ela_details"] = {"reason": "No faces found"}
        return state

    valid_faces = [f for f in face_detections if f.get("faces")]
    if not valid_faces:
        print("Node V5: Face detections present but no crops were generated.")
        state["texture_ela_score"] = 0.0
        state["texture_ela_details"] = {"reason": "No face crops available"}
        return state

    sorted_faces = sorted(
        valid_faces,
        key=lambda x: x["faces"][0]["confidence"] * x["faces"][0]["bbox"]["w"] * x["faces"][0]["bbox"]["h"],
        reverse=True,
    )
    selected_faces = sorted_faces[:max_faces]

    ela_dir = os.path.join(output_dir, "ela_analysis")
    os.makedirs(ela_dir, exist_ok=True)

    analysis_results = []
    ela_features = []

    api_key = os.getenv("OPENAI_API_KEY")
    client = None
    if api_key:
        client = OpenAI(api_key=api_key)
    else:
        print("Node V5: OPENAI_API_KEY not found. Skipping OpenAI analysis.")

    for i, face_data in enumerate(selected_faces):
        try:
            face_info = face_data["faces"][0]
            crop_path = face_info["crop_path"]

            if not os.path.exists(crop_path):
                continue

            original = Image.open(crop_path).convert("RGB")
            if state.get("v5_keep_temp_jpeg", False):
                original.save(os.path.join(ela_dir, f"temp_ela_{i}.jpg"), "JPEG", quality=quality)

            feats, enhanced = _analyze_crop(np.asarray(original), quality, device)
            feats["rank"] = i
            ela_features.append(feats)

            ela_output_path = os.path.join(ela_dir, f"ela_{i}.jpg")
            Image.fromarray(enhanced, "RGB").save(ela_output_path)

            gray_image = cv2.imread(crop_path, cv2.IMREAD_GRAYSCALE)
            if state.get("v5_gpu_fft", True):
                from v5ela import host as v5host

                magnitude_spectrum = v5host.spectrum_host(gray_image, device=device)
            else:                                   # the reference's host path, verbatim (v5_texture_ela.py:84-88)
                f = np.fft.fft2(gray_image)
                fshift = np.fft.fftshift(f)
                magnitude_spectrum = 20 * np.log(np.abs(fshift) + 1)
                magnitude_spectrum = cv2.normalize(magnitude_spectrum, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U)
            fft_output_path = os.path.join(ela_dir, f"fft_{i}.jpg")
            cv2.imwrite(fft_output_path, magnitude_spectrum)

            if client:
                def encode_image(image_path):
                    with open(image_path, "rb") as image_file:
                        return base64.b64encode(image_file.read()).decode("utf-8")

                base64_original = encode_image(crop_path)
                base64_ela = encode_image(ela_output_path)
                base64_fft = encode_image(fft_output_path)

                response = client.chat.completions.create(
                    model="gpt-4o",
                    messages=[
                        {
                            "role": "system",
                            "content": (
                                "You are a forensic image analyst specializing in deepfake detection. "
                                "You MUST return a JSON object (nothing else) with keys 'fake_probability' "
                                "and 'reasoning'."
                            ),
                        },
                        {
                            "role": "user",
                            "content": [
                                {"type": "text", "text": "Analyze this face for manipulation. Return JSON."},
                                {"type": "image_url", "image_url": {"url": f"data:image/jpeg;base64,{base64_original}"}},
                                {"type": "image_url", "image_url": {"url": f"data:image/jpeg;base64,{base64_ela}"}},
                                {"type": "image_url", "image_url": {"url": f"data:image/jpeg;base64,{base64_fft}"}},
                            ],
                        },
                    ],
                    response_format={"type": "json_object"},
                    timeout=30.0,
                )

                content = response.choices[0].message.content
                if not content:
                    if debug:
                        print(f"[DEBUG] V5: Empty response content for face {i}, skipping.")
                    continue
                try:
                    result_json = json.loads(content)
                    if isinstance(result_json, dict):
                        result_json.setdefault("ela_features", feats)
                    analysis_results.append(result_json)
                except Exception as parse_err:
                    print(f"Error parsing OpenAI response for face {i}: {parse_err}")
                    if debug:
                        print(f"[DEBUG] V5: Raw content: {content}")

        except Exception as e:
            print(f"Error analyzing face {i}: {e}")
            if debug:
                import traceback

                traceback.print_exc()

    def _safe_float(val, default=0.0):
        try:
            return float(val)
        except Exception:
            return default

    scores = []
    for r in analysis_results:
        if isinstance(r, dict):
            scores.append(_safe_float(r.get("fake_probability"), None))
        else:
            scores.append(_safe_float(r, None))
    scores = [s for s in scores if s is not None]

    if scores:
        avg_score = sum(scores) / len(scores)
        state["texture_ela_score"] = avg_score
        state["texture_ela_details"] = analysis_results
        print(f"Node V5: Analysis complete. Score: {avg_score:.2f}")
        dump_node_debug(
            state,
            "V5",
            {
                "faces_analyzed": len(analysis_results),
                "avg_score": avg_score,
                "ela_features": ela_features,
            },
        )
    else:
        print("Node V5: No analysis results generated.")
        state["texture_ela_score"] = 0.0
        details = {"reason": "Analysis failed or no keys"}
        if ela_features:
            details["ela_features"] = ela_features
        state["texture_ela_details"] = details

    return state

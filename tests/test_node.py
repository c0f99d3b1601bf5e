"""The drop-in node: the reference's own three tests (tests/test_v5_texture_ela.py) restated against the replacement
module, plus byte-level comparison of its artefacts with the files the UNMODIFIED reference node wrote
(tests/golden/node_golden.json, produced by tests/golden/make_golden.py)."""
import hashlib
import os
import shutil
import tempfile
from unittest.mock import MagicMock, patch

import cv2
import numpy as np
import pytest

from helpers import GOLDEN, load_json, sha

from nodes.V_nodes.v5_texture_ela import run

NODE_GOLDEN = load_json("node_golden.json")


def file_sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


@pytest.fixture
def fixture_state():
    """Same fixture as the reference test's setUp (tests/test_v5_texture_ela.py:17-43)."""
    test_dir = tempfile.mkdtemp()
    face_dir = os.path.join(test_dir, "faces")
    os.makedirs(face_dir)
    img = np.zeros((100, 100, 3), dtype=np.uint8)
    cv2.rectangle(img, (25, 25), (75, 75), (255, 255, 255), -1)
    face_path = os.path.join(face_dir, "face_0.jpg")
    cv2.imwrite(face_path, img)
    state = {
        "face_detections": [{"faces": [{"confidence": 0.99, "bbox": {"w": 100, "h": 100}, "crop_path": face_path}]}],
        "data_dir": test_dir,
        "debug": True,
    }
    yield state
    shutil.rmtree(test_dir)


def test_v5_no_faces():
    state = {"face_detections": [], "data_dir": "dummy"}
    result = run(state)
    assert result["texture_ela_score"] == 0.0
    assert result["texture_ela_details"]["reason"] == "No faces found"


def test_v5_faces_without_crops():
    result = run({"face_detections": [{"faces": []}], "data_dir": "dummy"})
    assert result["texture_ela_score"] == 0.0
    assert result["texture_ela_details"]["reason"] == "No face crops available"


@pytest.mark.gpu
@patch("nodes.V_nodes.v5_texture_ela.OpenAI")
def test_v5_run_basic(mock_openai, fixture_state):
    mock_client = MagicMock()
    mock_openai.return_value = mock_client
    mock_response = MagicMock()
    mock_response.choices[0].message.content = '{"fake_probability": 0.85, "reasoning": "Test reasoning"}'
    mock_client.chat.completions.create.return_value = mock_response
    with patch.dict(os.environ, {"OPENAI_API_KEY": "test_key"}):
        result = run(fixture_state)
    assert "texture_ela_score" in result
    assert result["texture_ela_score"] == 0.85
    assert len(result["texture_ela_details"]) == 1
    assert result["texture_ela_details"][0]["fake_probability"] == 0.85
    ela_dir = os.path.join(fixture_state["data_dir"], "ela_analysis")
    assert os.path.exists(os.path.join(ela_dir, "ela_0.jpg"))
    assert os.path.exists(os.path.join(ela_dir, "fft_0.jpg"))
    # artefacts are byte-identical to what the reference node wrote for this fixture
    gold = NODE_GOLDEN["cases"]["case0_ref_fixture"]["ranks"][0]
    assert file_sha(fixture_state["face_detections"][0]["faces"][0]["crop_path"]) == gold["crop_file_sha"]
    assert file_sha(os.path.join(ela_dir, "ela_0.jpg")) == gold["ela_file_sha"]
    assert file_sha(os.path.join(ela_dir, "fft_0.jpg")) == gold["fft_file_sha"]
    with open(os.path.join(GOLDEN, "node_case0_ref_fixture", "ela_0.jpg"), "rb") as f, \
            open(os.path.join(ela_dir, "ela_0.jpg"), "rb") as g:
        assert f.read() == g.read()
    # the reference's max_diff for this crop (golden residual maxima) is what the GPU record reports
    feats = result["texture_ela_details"][0]["ela_features"]
    assert feats["ela_max_rgb"] == gold["residual_max"] and feats["ela_max"] == max(gold["residual_max"])
    import json

    with open(os.path.join(fixture_state["data_dir"], "V5_debug.json")) as f:
        dbg = json.load(f)
    assert dbg["faces_analyzed"] == 1 and dbg["avg_score"] == 0.85


@pytest.mark.gpu
@patch("nodes.V_nodes.v5_texture_ela.OpenAI")
def test_v5_openai_failure(mock_openai, fixture_state):
    mock_client = MagicMock()
    mock_openai.return_value = mock_client
    mock_client.chat.completions.create.side_effect = Exception("API Error")
    with patch.dict(os.environ, {"OPENAI_API_KEY": "test_key"}):
        result = run(fixture_state)
    assert result["texture_ela_score"] == 0.0
    assert result["texture_ela_details"]["reason"] == "Analysis failed or no keys"


@pytest.mark.gpu
def test_v5_top3_selection_matches_reference_artefacts():
    """case1 of the node goldens: 4 faces, the reference keeps the top 3 by confidence*w*h and names files by rank."""
    from v5ela.synth import gen_frame

    case = NODE_GOLDEN["cases"]["case1_top3_of_4"]
    tmp = tempfile.mkdtemp()
    try:
        fdir = os.path.join(tmp, "faces")
        os.makedirs(fdir)
        dets = []
        for i, face in enumerate(case["faces"]):
            _, n, h, w, seed = face["src"]
            p = os.path.join(fdir, f"face_{i:06d}_0.jpg")
            cv2.imwrite(p, np.ascontiguousarray(gen_frame(n, h, w, seed)[..., ::-1]))
            dets.append({"frame_id": i, "timestamp": float(i),
                         "faces": [{"bbox": {"x": 0, "y": 0, "w": face["w"], "h": face["h"]},
                                    "confidence": face["confidence"], "is_main": True, "crop_path": p}]})
        env = {k: v for k, v in os.environ.items() if k != "OPENAI_API_KEY"}
        with patch.dict(os.environ, env, clear=True):
            res = run({"face_detections": dets, "data_dir": tmp, "debug": False, "v5_keep_temp_jpeg": True})
        assert res["texture_ela_score"] == case["score"] == 0.0
        assert res["texture_ela_details"]["reason"] == case["details"]["reason"]
        ela_dir = os.path.join(tmp, "ela_analysis")
        assert sorted(os.listdir(ela_dir)) == sorted(f"{k}_{r}.jpg" for k in ("temp_ela", "ela", "fft") for r in range(3))
        for rank, gold in enumerate(case["ranks"]):
            assert file_sha(os.path.join(fdir, f"face_{gold['face_index']:06d}_0.jpg")) == gold["crop_file_sha"]
            assert file_sha(os.path.join(ela_dir, f"ela_{rank}.jpg")) == gold["ela_file_sha"], rank
            assert file_sha(os.path.join(ela_dir, f"fft_{rank}.jpg")) == gold["fft_file_sha"], rank
            assert file_sha(os.path.join(ela_dir, f"temp_ela_{rank}.jpg")) == gold["temp_ela_file_sha"], rank
            feats = res["texture_ela_details"]["ela_features"][rank]
            assert feats["ela_max_rgb"] == gold["residual_max"]
    finally:
        shutil.rmtree(tmp)


def test_missing_gpu_is_reported_like_any_face_error(fixture_state, capsys):
    """Without a CUDA device the per-face try/except swallows the library error (reference :140-144 semantics)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    env = {k: v for k, v in os.environ.items() if k != "OPENAI_API_KEY"}
    with patch.dict(os.environ, env, clear=True):
        result = run(fixture_state)
    assert result["texture_ela_score"] == 0.0
    assert result["texture_ela_details"] == {"reason": "Analysis failed or no keys"}
    assert "Error analyzing face 0" in capsys.readouterr().out


@pytest.mark.gpu
def test_unsupported_crop_is_a_face_error_not_a_fallback(fixture_state, capsys):
    """A progressive JPEG is outside the GPU decoder's set: the node reports it like any per-face failure (reference
    :140-144) instead of silently reading it on the CPU; v5_gpu_codec=False is the caller's explicit switch."""
    from PIL import Image

    crop = fixture_state["face_detections"][0]["faces"][0]["crop_path"]
    Image.open(crop).save(crop, "JPEG", quality=95, progressive=True)
    env = {k: v for k, v in os.environ.items() if k != "OPENAI_API_KEY"}
    with patch.dict(os.environ, env, clear=True):
        result = run(dict(fixture_state))
        out = capsys.readouterr().out
        assert result["texture_ela_details"] == {"reason": "Analysis failed or no keys"}
        assert "Error analyzing face 0" in out and "-5" in out
        result = run(dict(fixture_state, v5_gpu_codec=False))
    assert "ela_features" in result["texture_ela_details"]
    assert os.path.exists(os.path.join(fixture_state["data_dir"], "ela_analysis", "ela_0.jpg"))

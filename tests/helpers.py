"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def sha(a) -> str:
    if isinstance(a, np.ndarray):
        a = np.ascontiguousarray(a).tobytes()
    return hashlib.sha256(a).hexdigest()[:16]


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_frame(case) -> np.ndarray:
    """Rebuild the input frame of a frames_golden.json case (same code as tests/golden/make_golden.py)."""
    from v5ela.synth import gen_frame

    spec, h, w = case["spec"], case["h"], case["w"]
    if spec[0] == "gen":
        return gen_frame(spec[1], h, w, spec[2])
    rng = np.random.default_rng(spec[1])
    kind = spec[0]
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


def record_matches_golden(rec, case) -> list:
    """List of field names where a structured record differs from a golden case (empty = match)."""
    bad = []
    for k in ("ela_max", "ela_sum", "ela_sumsq"):
        if [int(v) for v in rec[k]] != case[k]:
            bad.append(k)
    for k in ("tex_sumabs", "tex_sumsq", "tex_maxabs"):
        if int(rec[k]) != case[k]:
            bad.append(k)
    if sha(rec["ela_hist"]) != case["ela_hist_sha"]:
        bad.append("ela_hist")
    if sha(rec.tobytes()) != case["record_sha"]:
        bad.append("record_bytes")
    return bad


def layout_golden_files():
    """[(case, label, file bytes, golden entry)] of tests/golden/layouts_golden.json: the files are written again with the
    generator's own code (tests/golden/make_layouts_golden.py) and must hash to what it recorded."""
    import hashlib
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_layouts_golden", os.path.join(GOLDEN, "make_layouts_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out = []
    for case in load_json("layouts_golden.json")["cases"]:
        rgb = golden_frame(case)
        assert sha(rgb) == case["in_sha"]
        written = dict(gen.layout_files(rgb, case["q"]))
        for entry in case["files"]:
            data = written[entry["label"]]
            assert (len(data), hashlib.sha256(data).hexdigest()[:16]) == (entry["file_len"], entry["file_sha"]), (case["spec"], entry["label"])
            out.append((case, entry["label"], data, entry))
    return out


def check_layout_goldens(decode):
    """decode(list of file bytes) -> list of dict(rgb, gray); every file of layouts_golden.json must decode to the recorded pixels."""
    items = layout_golden_files()
    outs = decode([data for _, _, data, _ in items])
    assert len(outs) == len(items) == 100
    for (case, label, _, entry), o in zip(items, outs):
        assert sha(o["rgb"]) == entry["dec_rgb_sha"], (case["spec"], case["h"], case["w"], label)
        assert sha(o["gray"]) == entry["dec_y_sha"], (case["spec"], case["h"], case["w"], label)

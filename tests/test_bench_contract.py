"""CPU: bench.py's command-line contract for the arm that runs without a GPU (--impl reference): exactly one JSON line on
stdout with the keys the driver reads; and the native arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

from helpers import HERE

ROOT = os.path.dirname(HERE)


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p_keyframes_per_sec_v5_ela_texture" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("256 synthetic 1920x1080")


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_native_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and p.stdout.strip() == ""
    assert "no CUDA device" in p.stderr


def test_profile_source_hash_ignores_comments_and_matches_the_committed_profile():
    """bench.py says whether the ncu figures it quotes (roofline.traffic, int_issue) describe the build it timed by comparing a
    hash of csrc/ with the one stamped into profiles/fused_kernel_*.json. The hash ignores comments and white space, changes with
    any token, and the committed profiles carry the hash of the committed sources."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("v5_source_hash", os.path.join(ROOT, "profiles", "source_hash.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a = 'int f(int x) { return x /* not a comment in "s" */ + 1; }  // tail\nconst char *s = "a // b /* c */";\n'
    b = 'int f(int x)\n{\n    return x + 1;\n}\nconst char *s = "a // b /* c */";'
    assert mod.strip_code(a).strip() == mod.strip_code(b).strip()
    assert mod.strip_code(a) != mod.strip_code(a.replace("+ 1", "+ 2"))
    assert mod.strip_code(a) != mod.strip_code(a.replace('"a // b', '"a / b'))
    h = mod.source_hash(os.path.join(ROOT, "fake-video-detection-engine_b200", "csrc"))
    for name in ("fused_kernel_dram.json", "fused_kernel_issue.json"):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            assert json.load(f)["source_hash"] == h, f"{name}: re-capture (profiles/final_run_r02.sh) and profiles/derive_fused_json.py after a kernel change"

"""GPU: the BENCHMARKED kernel instantiation — ela_fused_kernel<FAST> (width % 16 == 0, aligned frames, records only: what every
BASELINE.json config and bench.py drive) — against the oracle, directly: every call here goes through ``analyze_batch`` WITHOUT a
residual map and asserts, through ``v5ela_last_instantiation``, that the FAST instantiation really ran (round-1 VERDICT, weak #1:
the general instantiation was the only one compared with the oracle on the GPU).

Oracles: oracle/c_oracle (restatement of libjpeg, pinned on the reference's goldens), oracle/pil_oracle (the reference's own Pillow
calls, v5_texture_ela.py:66-73) and SURVEY.md Appendix B's known-answer table. Bar: records byte-for-byte."""
import numpy as np
import pytest

from oracle import c_oracle, pil_oracle
from v5ela.records import as_records, combine
from v5ela.synth import gen_batch_torch, gen_frame

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("block_stage")]   # every test under both block-stage builds

# SURVEY.md Appendix B rows (Pillow 12.2.0 / libjpeg-turbo 3.1.4.1 at survey time)
APPENDIX_B = {
    (0, 720, 1280, 90): dict(ela_max=[15, 13, 15], ela_sum=[2908862, 2466142, 3163775], ela_sumsq=[13616386, 9810180, 16022737],
                             hist_r=[83745, 164503, 159362, 146352], lap=(8456825, 115570645, 40)),
    (0, 1080, 1920, 75): dict(ela_max=[16, 13, 21], ela_sum=[7025028, 6363248, 7484461], ela_sumsq=[34903406, 28143590, 40001245],
                              hist_r=[172689, 341145, 333915, 315306], lap=(19016872, 259819512, 40)),
    (0, 1080, 1920, 85): dict(ela_max=[16, 13, 17], ela_sum=[6779848, 5931197, 7209212], ela_sumsq=[32749462, 24801821, 36857400],
                              hist_r=[181602, 355788, 345566, 322940], lap=(19016872, 259819512, 40)),
    (0, 1080, 1920, 90): dict(ela_max=[15, 13, 16], ela_sum=[6545756, 5547703, 7120886], ela_sumsq=[30661720, 22057949, 36084286],
                              hist_r=[188552, 370860, 357776, 329172], lap=(19016872, 259819512, 40)),
    (0, 1080, 1920, 95): dict(ela_max=[16, 11, 17], ela_sum=[6313117, 4550693, 7264677], ela_sumsq=[29355535, 15598263, 38468685],
                              hist_r=[208654, 398266, 369102, 323498], lap=(19016872, 259819512, 40)),
    (5, 2160, 3840, 90): dict(ela_max=[16, 13, 17], ela_sum=[26182514, 22179221, 28471595], ela_sumsq=[122685342, 88155219, 144241389],
                              hist_r=[758371, 1482740, 1425684, 1314899], lap=(76098742, 1039975130, 44)),
}


def run_fast(t, q=90):
    """records (host, structured) of a device batch through the records-only call; asserts the FAST instantiation ran."""
    import torch
    import v5ela
    from v5ela.batch import get_handle

    out = v5ela.analyze_batch(t, quality=q)
    torch.cuda.synchronize()
    assert "residual" not in out
    assert get_handle(t.device.index).last_instantiation == "fast"
    return as_records(out["records"])


def check_appendix_b(rec, key):
    g = APPENDIX_B[key]
    assert rec["ela_max"].tolist() == g["ela_max"]
    assert rec["ela_sum"].tolist() == g["ela_sum"]
    assert rec["ela_sumsq"].tolist() == g["ela_sumsq"]
    assert rec["ela_hist"][0][:4].tolist() == g["hist_r"]
    assert (int(rec["tex_sumabs"]), int(rec["tex_sumsq"]), int(rec["tex_maxabs"])) == g["lap"]


def test_instantiation_selector(block_stage):
    """The host's choice (csrc/v5ela_host.h fast_path_ok) as the ABI reports it."""
    import torch
    import v5ela
    from v5ela.batch import get_handle

    hd = get_handle(torch.cuda.current_device())
    assert hd.block_stage == block_stage
    t = gen_batch_torch(0, 2, 64, 64, seed=0)
    v5ela.analyze_batch(t)
    assert hd.last_instantiation == "fast"
    v5ela.analyze_batch(t, want_residual=True)
    assert hd.last_instantiation == "general"
    v5ela.analyze_batch(t, want_tex_hist=True)
    assert hd.last_instantiation == "texhist"
    v5ela.analyze_batch(gen_batch_torch(0, 2, 64, 72, seed=0))         # width not a multiple of 16
    assert hd.last_instantiation == "general"
    v5ela.analyze_batch(t[:, :, 1:49])                                   # 48 wide but unaligned rows
    assert hd.last_instantiation == "general"
    torch.cuda.synchronize()


def test_config1_16x720p_fast_vs_oracle():
    """BASELINE.json configs[0] shape: 16 x 1280x720, q=90 (the reference's CPU-runnable case)."""
    t = gen_batch_torch(0, 16, 720, 1280, seed=0)
    recs = run_fast(t)
    host = t.cpu().numpy()
    orecs, _ = c_oracle.analyze(host, 90)
    assert recs.tobytes() == orecs.tobytes()
    check_appendix_b(recs[0], (0, 720, 1280, 90))
    for i in (0, 9):                                                     # the reference's own Pillow calls
        assert recs[i].tobytes() == pil_oracle.record(host[i], 90).tobytes()


def test_config2_256x1080p_fast_vs_oracle():
    """BASELINE.json configs[1], the benchmarked batch itself: 256 x 1920x1080, q=90. 64 of its frames (frame 0 = Appendix B, then
    every fourth) against the C oracle byte for byte, 2 against Pillow; all 256 against the general instantiation."""
    import torch
    import v5ela

    t = gen_batch_torch(0, 256, 1080, 1920, seed=0)
    recs = run_fast(t)
    check_appendix_b(recs[0], (0, 1080, 1920, 90))
    for i in range(0, 256, 4):
        o = c_oracle.analyze_frame(t[i].cpu().numpy(), 90)
        assert recs[i].tobytes() == o["record"].tobytes(), i
    for i in (1, 255):
        assert recs[i].tobytes() == pil_oracle.record(t[i].cpu().numpy(), 90).tobytes(), i
    for lo in range(0, 256, 64):                                         # general instantiation (residual map wanted), 64 frames at a time
        gen = v5ela.analyze_batch(t[lo:lo + 64], want_residual=True)
        torch.cuda.synchronize()
        assert as_records(gen["records"]).tobytes() == recs[lo:lo + 64].tobytes()
        del gen
    assert (recs["ela_hist"].sum(axis=2) == 1080 * 1920).all()


def test_config3_4k_fast_vs_oracle():
    """BASELINE.json configs[2] shape: 3840x2160, q=90 — frames 4..7 of the generator, frame 5 = Appendix B."""
    t = gen_batch_torch(4, 4, 2160, 3840, seed=0)
    recs = run_fast(t)
    check_appendix_b(recs[1], (5, 2160, 3840, 90))
    for i in range(4):
        o = c_oracle.analyze_frame(t[i].cpu().numpy(), 90)
        assert recs[i].tobytes() == o["record"].tobytes(), i


@pytest.mark.parametrize("q", [75, 85, 90, 95])
def test_config5_quality_sweep_1080p_fast_vs_oracle(q):
    """BASELINE.json configs[4]: q in {75, 85, 90, 95} at 1080p (quant-table variants), Appendix B rows + oracle + Pillow."""
    t = gen_batch_torch(0, 8, 1080, 1920, seed=0)
    recs = run_fast(t, q)
    check_appendix_b(recs[0], (0, 1080, 1920, q))
    host = t.cpu().numpy()
    orecs, _ = c_oracle.analyze(host, q)
    assert recs.tobytes() == orecs.tobytes()
    assert recs[3].tobytes() == pil_oracle.record(host[3], q).tobytes()


def test_config4_per_video_reduce_1080p_fast_vs_oracle():
    """BASELINE.json configs[3] shape at full frame size: videos x 32 keyframes of 1080p, video v frame k = gen_frame(32 v + k,
    seed = v), per-video aggregation on the device == combine() of the oracle's per-frame records."""
    import torch
    import v5ela
    from v5ela.batch import get_handle

    videos, per = 2, 32
    t = torch.cat([gen_batch_torch(per * v, per, 1080, 1920, seed=v) for v in range(videos)])
    out = v5ela.analyze_batch(t)
    agg = as_records(v5ela.reduce_records(out["records"], per))
    torch.cuda.synchronize()
    assert get_handle(t.device.index).last_instantiation == "fast"
    recs = as_records(out["records"])
    host = t.cpu().numpy()
    assert np.array_equal(host[per + 3], gen_frame(per + 3, 1080, 1920, 1))      # the device generator is the NumPy generator
    orecs, _ = c_oracle.analyze(host, 90)
    assert recs.tobytes() == orecs.tobytes()
    for v in range(videos):
        assert agg[v].tobytes() == combine(orecs[per * v:per * (v + 1)]).tobytes(), v


@pytest.mark.parametrize("hw", [(16, 16), (16, 32), (17, 48), (8, 496), (33, 512), (250, 976), (1, 1936), (1080, 16), (64, 4096)])
def test_fast_small_and_odd_heights_vs_oracle(hw):
    """FAST only constrains the width: odd heights, one-row frames, one-MCU-wide frames, several strips, noise content."""
    import torch

    h, w = hw
    rng = np.random.default_rng(h * 7919 + w)
    frames = np.stack([gen_frame(2, h, w, 5), rng.integers(0, 256, (h, w, 3), dtype=np.uint8),
                       np.where(rng.integers(0, 2, (h, w, 3)) > 0, 255, 0).astype(np.uint8)])
    for q in (90, 1, 100):
        recs = run_fast(torch.from_numpy(frames).cuda(), q)
        orecs, _ = c_oracle.analyze(frames, q)
        assert recs.tobytes() == orecs.tobytes(), q


def test_fast_goldens_with_width_multiple_of_16():
    """Every frames_golden.json case whose width is a multiple of 16, through the FAST instantiation."""
    import torch
    from helpers import golden_frame, load_json, record_matches_golden

    n = 0
    for case in load_json("frames_golden.json")["cases"]:
        if case["w"] % 16:
            continue
        recs = run_fast(torch.from_numpy(golden_frame(case)[None]).cuda(), case["q"])
        assert record_matches_golden(recs[0], case) == [], (case["spec"], case["h"], case["w"], case["q"])
        n += 1
    assert n >= 5

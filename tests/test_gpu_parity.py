"""GPU: the CUDA path (through the C ABI, via v5ela.analyze_batch) against the oracle and the reference goldens.

Bar: residual maps, histograms and every integer record field bit-exact; derived float statistics within 1e-5 relative
(they are computed in float64 from identical integers, so they are in fact equal).
"""
import numpy as np
import pytest

from helpers import golden_frame, load_json, record_matches_golden, sha
from oracle import c_oracle, pil_oracle
from v5ela.records import as_records, features
from v5ela.synth import gen_batch, gen_batch_torch, gen_frame

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("block_stage")]   # every test under both block-stage builds

FRAMES = load_json("frames_golden.json")


def run_gpu(frames_np, q=90, **kw):
    import torch
    import v5ela

    t = torch.from_numpy(np.ascontiguousarray(frames_np)).cuda()
    out = v5ela.analyze_batch(t, quality=q, want_residual=True, **kw)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res["records"] = as_records(res["records"])
    return res


def test_native_library_is_loaded():
    import torch
    from v5ela import _abi

    lib = _abi.load()
    assert lib.v5ela_record_bytes() == 3144
    with open("/proc/self/maps") as f:
        assert "libv5ela.so" in f.read()
    h = _abi.Handle(torch.cuda.current_device())
    lu, ch = h.quant_tables()
    assert lu[0].tolist() == [3, 2, 2, 3, 5, 8, 10, 12] and ch[0].tolist() == [3, 4, 5, 9, 20, 20, 20, 20]  # q=90, A.1
    h.close()


@pytest.mark.parametrize("case", FRAMES["cases"], ids=lambda c: f"{c['spec'][0]}{c['spec'][1]}_{c['h']}x{c['w']}_q{c['q']}")
def test_reference_goldens(case):
    frame = golden_frame(case)
    out = run_gpu(frame[None], case["q"], want_enhanced=True)
    assert sha(out["residual"][0]) == case["resid_sha"]
    assert record_matches_golden(out["records"][0], case) == []
    assert sha(out["enhanced"][0]) == case["enhanced_sha"]


def test_small_batch_vs_oracle_all_fields():
    frames = gen_batch(10, 6, 211, 173, seed=3)
    out = run_gpu(frames, 90)
    recs, resid = c_oracle.analyze(frames, 90, want_residual=True)
    assert np.array_equal(out["residual"], resid)
    assert out["records"].tobytes() == recs.tobytes()


@pytest.mark.parametrize("hw", [(1, 1), (1, 2), (2, 1), (2, 3), (3, 4), (4, 5), (7, 7), (8, 9), (15, 17), (16, 16), (17, 33),
                                (32, 48), (33, 497), (500, 31), (100, 1000), (271, 481), (16, 512), (48, 960), (49, 961)])
@pytest.mark.parametrize("q", [90, 30])
def test_ragged_sizes_vs_oracle(hw, q):
    h, w = hw
    rng = np.random.default_rng(h * 1000 + w)
    frames = np.stack([gen_frame(4, h, w, 9), rng.integers(0, 256, (h, w, 3), dtype=np.uint8)])
    out = run_gpu(frames, q)
    recs, resid = c_oracle.analyze(frames, q, want_residual=True)
    assert np.array_equal(out["residual"], resid)
    assert out["records"].tobytes() == recs.tobytes()


@pytest.mark.parametrize("q", [1, 10, 50, 75, 85, 90, 95, 100])
def test_quality_sweep_vs_pil(q):
    frame = gen_frame(2, 270, 480, 1)
    out = run_gpu(frame[None], q)
    rec, resid = pil_oracle.record(frame, q, with_residual=True)
    assert np.array_equal(out["residual"][0], resid)
    assert out["records"][0].tobytes() == rec.tobytes()


def test_padded_strides_and_unaligned_views():
    import torch
    import v5ela

    base = torch.from_numpy(gen_batch(0, 3, 100, 150, seed=2)).cuda()
    big = torch.zeros((3, 120, 170, 3), dtype=torch.uint8, device="cuda")
    big[:, 7:107, 5:155] = base
    view = big[:, 7:107, 5:155]                                  # padded row/frame strides, unaligned start
    out = v5ela.analyze_batch(view, want_residual=True)
    ref = v5ela.analyze_batch(base, want_residual=True)
    torch.cuda.synchronize()
    assert torch.equal(out["residual"], ref["residual"]) and torch.equal(out["records"], ref["records"])
    recs, resid = c_oracle.analyze(base.cpu().numpy(), 90, want_residual=True)
    assert np.array_equal(ref["residual"].cpu().numpy(), resid)


def test_records_only_mode_equals_residual_mode():
    import torch
    import v5ela

    t = gen_batch_torch(0, 4, 360, 640, seed=5)
    a = v5ela.analyze_batch(t)
    b = v5ela.analyze_batch(t, want_residual=True)
    torch.cuda.synchronize()
    assert "residual" not in a and torch.equal(a["records"], b["records"])


def test_torch_generator_matches_numpy():
    t = gen_batch_torch(3, 2, 97, 131, seed=4).cpu().numpy()
    assert np.array_equal(t, gen_batch(3, 2, 97, 131, seed=4))


def test_1080p_known_answers_and_properties():
    """BASELINE.json config 2 shape: 1080p, q=90. Frame 0 == Appendix B hash; size-independent properties on all."""
    import torch
    import v5ela

    n = 16
    t = gen_batch_torch(0, n, 1080, 1920, seed=0)
    out = v5ela.analyze_batch(t, want_residual=True, want_enhanced=True)
    torch.cuda.synchronize()
    recs = as_records(out["records"])
    resid0 = out["residual"][0].cpu().numpy()
    assert sha(resid0) == "7c7db2c090891020"                      # SURVEY Appendix B, gen_frame(0,1080,1920,0) q=90
    assert recs[0]["ela_max"].tolist() == [15, 13, 16]
    assert recs[0]["ela_sum"].tolist() == [6545756, 5547703, 7120886]
    assert (int(recs[0]["tex_sumabs"]), int(recs[0]["tex_sumsq"]), int(recs[0]["tex_maxabs"])) == (19016872, 259819512, 40)
    # properties at full size, every frame: histogram <-> residual map <-> sums consistency (checksum of checksums)
    res = out["residual"]
    for i in range(n):
        for c in range(3):
            ch = res[i, :, :, c].reshape(-1)
            hist = torch.bincount(ch.to(torch.int64), minlength=256).cpu().numpy()
            assert np.array_equal(hist, recs[i]["ela_hist"][c])
            assert int(recs[i]["ela_hist"][c].sum()) == 1080 * 1920
            assert int(ch.to(torch.int64).sum()) == int(recs[i]["ela_sum"][c])
            assert int(ch.max()) == int(recs[i]["ela_max"][c])
    # three more frames against the C oracle in full
    for i in (1, 7, 15):
        o = c_oracle.analyze_frame(t[i].cpu().numpy(), 90)
        assert np.array_equal(out["residual"][i].cpu().numpy(), o["residual"])
        assert recs[i].tobytes() == o["record"].tobytes()
    # idempotence of the enhancement LUT definition: enhanced == lut[max][residual]
    for i in (0, 5):
        m = int(recs[i]["ela_max"].max())
        lut = c_oracle.enhance_lut(m)
        assert np.array_equal(out["enhanced"][i].cpu().numpy(), lut[out["residual"][i].cpu().numpy()])
    f = features(recs[0], 1080 * 1920)
    ref = features(c_oracle.analyze_frame(t[0].cpu().numpy(), 90)["record"], 1080 * 1920)
    for k, v in f.items():
        assert np.allclose(v, ref[k], rtol=1e-5, atol=0), k       # north_star tolerance for float statistics


def test_4k_frame_known_answer():
    out = run_gpu(gen_frame(5, 2160, 3840, 0)[None], 90)
    assert sha(out["residual"][0]) == "3f81ee414328ef0c"
    assert out["records"][0]["ela_sum"].tolist() == [26182514, 22179221, 28471595]
    assert int(out["records"][0]["tex_sumsq"]) == 1039975130


def test_noise_frames_full_range_residuals():
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (2, 360, 640, 3), dtype=np.uint8)
    out = run_gpu(frames, 90)
    recs, resid = c_oracle.analyze(frames, 90, want_residual=True)
    assert np.array_equal(out["residual"], resid) and out["records"].tobytes() == recs.tobytes()
    assert int(out["records"]["ela_max"].max()) > 200


def test_reduce_records_per_video():
    import torch
    import v5ela
    from v5ela.records import combine

    t = gen_batch_torch(0, 12, 180, 320, seed=1)
    out = v5ela.analyze_batch(t)
    agg = v5ela.reduce_records(out["records"], 4)
    torch.cuda.synchronize()
    recs, aggr = as_records(out["records"]), as_records(agg)
    for g in range(3):
        assert aggr[g].tobytes() == combine(recs[4 * g:4 * g + 4]).tobytes()


def test_analyze_host_entry_point():
    import ctypes

    import torch
    from v5ela import _abi

    frames = gen_batch(2, 3, 120, 200, seed=6)
    h = _abi.Handle(torch.cuda.current_device())
    recs = np.zeros(3, dtype=as_records(np.zeros((1, 3144), np.uint8)).dtype)
    resid = np.zeros_like(frames)
    enh = np.zeros_like(frames)
    h.analyze_host(frames.ctypes.data, 3, 120, 200, recs.ctypes.data, resid.ctypes.data, enh.ctypes.data)
    orecs, oresid = c_oracle.analyze(frames, 90, want_residual=True)
    assert np.array_equal(resid, oresid) and recs.tobytes() == orecs.tobytes()
    for i in range(3):
        assert np.array_equal(enh[i], pil_oracle.ela_enhanced(frames[i], 90))
    assert h.launch_count >= 3
    h.close()


def test_error_codes_do_not_raise_cuda_faults():
    import torch
    from v5ela import _abi

    h = _abi.Handle(torch.cuda.current_device())
    with pytest.raises(_abi.V5ElaError):
        h.set_quality(0)
    with pytest.raises(_abi.V5ElaError):
        h.analyze(0, 1, 16, 16, 768, 48, 0, None, None)          # null pointers
    t = torch.zeros((1, 16, 16, 3), dtype=torch.uint8, device="cuda")
    r = torch.zeros((1, 3144), dtype=torch.uint8, device="cuda")
    with pytest.raises(_abi.V5ElaError):
        h.analyze(t.data_ptr(), 1, 16, 16, 768, 47, r.data_ptr(), None, None)   # row stride < 3*W
    h.analyze(t.data_ptr(), 1, 16, 16, 768, 48, r.data_ptr(), None, None)
    torch.cuda.synchronize()
    h.close()


def test_8k_frame_and_empty_batch():
    """Maximum-size style input (7680x4320, 99.5 MB) and the empty batch."""
    import torch
    import v5ela

    frame = gen_frame(2, 4320, 7680, 1)
    out = run_gpu(frame[None], 90)
    o = c_oracle.analyze_frame(frame, 90)
    assert np.array_equal(out["residual"][0], o["residual"])
    assert out["records"][0].tobytes() == o["record"].tobytes()
    empty = v5ela.analyze_batch(torch.empty((0, 64, 64, 3), dtype=torch.uint8, device="cuda"), want_residual=True)
    assert empty["records"].shape == (0, 3144) and empty["residual"].shape == (0, 64, 64, 3)


def test_many_small_frames_dynamic_tickets():
    """More work items than resident CTAs, of unequal size (ragged strips and segments): every item is done exactly once."""
    frames = gen_batch(0, 700, 40, 1000, seed=2)                 # 700 frames x 3 strips (21+21+21 MCUs, ragged edge)
    out = run_gpu(frames, 90)
    recs, resid = c_oracle.analyze(frames, 90, want_residual=True)
    assert np.array_equal(out["residual"], resid)
    assert out["records"].tobytes() == recs.tobytes()


def test_repeated_calls_are_deterministic():
    import torch
    import v5ela

    t = gen_batch_torch(0, 8, 360, 640, seed=9)
    a = v5ela.analyze_batch(t)["records"].clone()
    for _ in range(5):
        b = v5ela.analyze_batch(t)["records"]
        torch.cuda.synchronize()
        assert torch.equal(a, b)


def test_one_handle_on_two_streams():
    """A thread's cached handle used alternately on two CUDA streams (round-1 ADVICE): the handle-owned scratch — the work-item
    ticket of the fused kernel, the codec and spectrum workspaces — is reused by every call, so a call on another stream has to
    wait for the previous call's kernels. Long launches on one stream, short ones on the other, results against serial runs."""
    import torch
    import v5ela
    from v5ela import jpeg

    big = gen_batch_torch(0, 48, 720, 1280, seed=3)
    small = gen_batch_torch(7, 3, 96, 160, seed=4)
    gray = small[..., 1].contiguous()
    ref_big = v5ela.analyze_batch(big)["records"].clone()
    ref_small = v5ela.analyze_batch(small)["records"].clone()
    ref_spec = v5ela.spectrum_batch(gray).clone()
    ref_enc, ref_sizes = jpeg.encode_batch(small, 90)
    ref_enc, ref_sizes = ref_enc.clone(), ref_sizes.clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for it in range(12):
        with torch.cuda.stream(s1):
            a = v5ela.analyze_batch(big)["records"]
        with torch.cuda.stream(s2):
            b = v5ela.analyze_batch(small)["records"]
            sp = v5ela.spectrum_batch(gray) if it % 3 == 0 else None
            enc = jpeg.encode_batch(small, 90) if it % 3 == 1 else None
        outs.append((a, b, sp, enc))
    torch.cuda.synchronize()
    for a, b, sp, enc in outs:
        assert torch.equal(a, ref_big) and torch.equal(b, ref_small)
        if sp is not None:
            assert torch.equal(sp, ref_spec)
        if enc is not None:
            assert torch.equal(enc[1], ref_sizes)
            for i in range(3):
                n = int(ref_sizes[i])
                assert torch.equal(enc[0][i, :n], ref_enc[i, :n])


def test_two_host_threads_with_their_own_handles():
    """The node may run next to V2/V3/V4 on LangGraph's worker threads: one handle per thread (thread-local in
    v5ela.batch / v5ela.host), concurrent calls, identical results."""
    import threading

    import torch
    import v5ela
    from v5ela import host as v5host, jpeg

    frames = gen_batch(20, 6, 200, 312, seed=9)
    ref_recs, ref_resid = c_oracle.analyze(frames, 90, want_residual=True)
    ref_file = c_oracle.jpeg_encode(frames[0], 75)
    errors = []

    def worker(kind):
        try:
            for _ in range(8):
                if kind == 0:
                    t = torch.from_numpy(frames).cuda()
                    out = v5ela.analyze_batch(t, quality=90, want_residual=True)
                    torch.cuda.synchronize()
                    assert np.array_equal(out["residual"].cpu().numpy(), ref_resid)
                    assert as_records(out["records"].cpu()).tobytes() == ref_recs.tobytes()
                elif kind == 1:
                    recs, resid, _ = v5host.analyze_frames_host(frames, quality=90, want_residual=True)
                    assert recs.tobytes() == ref_recs.tobytes() and np.array_equal(resid, ref_resid)
                else:
                    data = jpeg.encode_host(frames[:1], 75)[0]
                    assert data == ref_file
                    assert np.array_equal(jpeg.decode_host([data])[0]["rgb"], c_oracle.jpeg_decode(ref_file)["rgb"])
        except Exception as e:  # noqa: BLE001
            errors.append((kind, repr(e)))

    threads = [threading.Thread(target=worker, args=(k % 3,)) for k in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert errors == []


def test_texture_histogram_optional_output():
    """tex_hist[256] (optional): per frame bincount(min(|cv2.Laplacian(Y)|, 255)) through v5ela_analyze_ex;
    records identical to the call without it (which takes another kernel instantiation); ragged sizes, full size, a batch."""
    import torch
    from oracle import pil_oracle
    from v5ela import analyze_batch
    from v5ela.synth import gen_frame

    rng = np.random.default_rng(9)
    for h, w in [(1, 1), (9, 1), (17, 33), (64, 64), (270, 481), (1080, 1920)]:
        frames = np.stack([gen_frame(0, h, w, 4), rng.integers(0, 256, (h, w, 3), dtype=np.uint8),
                           np.where(rng.integers(0, 2, (h, w, 1)) > 0, 255, 0).astype(np.uint8).repeat(3, axis=2)])
        t = torch.from_numpy(frames).cuda()
        out = analyze_batch(t, want_tex_hist=True)
        plain = analyze_batch(t)
        with_res = analyze_batch(t, want_tex_hist=True, want_residual=True)
        th = out["tex_hist"].cpu().numpy().view(np.uint32)
        for i in range(3):
            assert np.array_equal(th[i], pil_oracle.texture_hist(frames[i])), (h, w, i)
        assert int(th.sum()) == 3 * h * w
        assert torch.equal(out["records"], plain["records"]) and torch.equal(with_res["records"], plain["records"])
        assert torch.equal(with_res["tex_hist"], out["tex_hist"])


def test_texture_histogram_goldens():
    """tests/golden/layouts_golden.json: tex_hist recorded from OpenCV's Laplacian of Pillow's luma, through v5ela_analyze_ex."""
    import torch
    from helpers import golden_frame, load_json, sha
    from v5ela import analyze_batch

    for case in load_json("layouts_golden.json")["cases"]:
        t = torch.from_numpy(golden_frame(case)[None]).cuda()
        th = analyze_batch(t, want_tex_hist=True)["tex_hist"].cpu().numpy().view(np.uint32)[0]
        assert sha(np.ascontiguousarray(th)) == case["tex_hist_sha"], (case["spec"], case["h"], case["w"])
        assert int(th[255]) == case["tex_hist_last"]

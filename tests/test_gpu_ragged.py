"""GPU: the ragged-batch entry points — frames of different sizes in ONE launch (v5ela_analyze_ragged / _host, include/v5ela.h),
the form the reference node's real input has: at most three face crops of different sizes per call (v5_texture_ela.py:42, 56-64,
crops cut at v1_keyframes_facetrack.py:144-166). Against the C oracle frame by frame, bit for bit; both block-stage builds."""
import numpy as np
import pytest

from oracle import c_oracle
from v5ela.records import as_records
from v5ela.synth import gen_frame

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("block_stage")]

SIZES = [(1, 1), (2, 3), (9, 1), (16, 16), (17, 33), (64, 96), (257, 301), (100, 1000), (33, 497), (360, 640), (271, 481), (1080, 1920)]


def mixed_frames(seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for k, (h, w) in enumerate(SIZES):
        if k % 3 == 1:
            out.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        elif k % 3 == 2:
            out.append(np.where(rng.integers(0, 2, (h, w, 3)) > 0, 255, 0).astype(np.uint8))
        else:
            out.append(gen_frame(k, h, w, seed))
    return out


@pytest.mark.parametrize("q", [90, 30])
def test_ragged_mixed_sizes_vs_oracle(q):
    import torch
    import v5ela
    from v5ela.batch import get_handle

    frames = mixed_frames(q)
    dev = [torch.from_numpy(f).cuda() for f in frames]
    out = v5ela.analyze_ragged(dev, quality=q, want_residual=True, want_enhanced=True)
    torch.cuda.synchronize()
    assert get_handle(torch.cuda.current_device()).last_instantiation == "general"
    recs = as_records(out["records"])
    for i, f in enumerate(frames):
        o = c_oracle.analyze_frame(f, q)
        assert recs[i].tobytes() == o["record"].tobytes(), (i, f.shape)
        assert np.array_equal(out["residual"][i].cpu().numpy(), o["residual"]), (i, f.shape)
        lut = c_oracle.enhance_lut(int(o["record"]["ela_max"].max()))
        assert np.array_equal(out["enhanced"][i].cpu().numpy(), lut[o["residual"]]), (i, f.shape)
    # records only: same records, nothing else written
    plain = v5ela.analyze_ragged(dev, quality=q)
    torch.cuda.synchronize()
    assert torch.equal(plain["records"], out["records"]) and "residual" not in plain


def test_ragged_views_into_one_keyframe():
    """Crops as strided, unaligned views of a device-resident frame (what v5ela.handoff feeds): no copies, one launch."""
    import torch
    import v5ela

    frame = gen_frame(3, 720, 1280, 2)
    t = torch.from_numpy(frame).cuda()
    boxes = [(0, 0, 720, 1280), (13, 7, 300, 411), (400, 1000, 720, 1280), (100, 100, 101, 101), (5, 640, 250, 1279), (719, 0, 720, 1280),
             (0, 1279, 720, 1280), (333, 222, 590, 523)]
    views = [t[y1:y2, x1:x2] for (y1, x1, y2, x2) in boxes]
    out = v5ela.analyze_ragged(views, want_residual=True)
    torch.cuda.synchronize()
    recs = as_records(out["records"])
    for i, (y1, x1, y2, x2) in enumerate(boxes):
        o = c_oracle.analyze_frame(np.ascontiguousarray(frame[y1:y2, x1:x2]), 90)
        assert recs[i].tobytes() == o["record"].tobytes(), boxes[i]
        assert np.array_equal(out["residual"][i].cpu().numpy(), o["residual"]), boxes[i]


def test_ragged_equals_uniform_batches():
    """A ragged call over equally sized frames == the uniform call (records), and a large ragged batch fills the GPU correctly."""
    import torch
    import v5ela
    from v5ela.synth import gen_batch_torch

    t = gen_batch_torch(0, 40, 200, 304, seed=6)
    uni = v5ela.analyze_batch(t)["records"]
    rag = v5ela.analyze_ragged([t[i] for i in range(40)])["records"]
    torch.cuda.synchronize()
    assert torch.equal(uni, rag)


def test_ragged_host_entry_point_vs_oracle():
    from v5ela import host as v5host

    frames = mixed_frames(5)[:9]
    recs, resid, enh = v5host.analyze_ragged_host(frames, quality=90, want_residual=True, want_enhanced=True)
    for i, f in enumerate(frames):
        o = c_oracle.analyze_frame(f, 90)
        assert recs[i].tobytes() == o["record"].tobytes(), (i, f.shape)
        assert np.array_equal(resid[i], o["residual"])
        assert np.array_equal(enh[i], c_oracle.enhance_lut(int(o["record"]["ela_max"].max()))[o["residual"]])
    recs2, r2, e2 = v5host.analyze_ragged_host(frames, quality=90, want_enhanced=True)      # enhanced alone: residual formed in place
    assert recs2.tobytes() == recs.tobytes() and r2 is None
    for a, b in zip(e2, enh):
        assert np.array_equal(a, b)


def test_three_face_crops_gpu_time():
    """VERDICT r01 task 5: the node's case — three 257x301 crops — costs one launch sequence and <= 80 us of fused-kernel time."""
    import torch
    import v5ela
    from v5ela.batch import get_handle

    crops = [torch.from_numpy(gen_frame(i, 257, 301, 1)).cuda() for i in range(3)]
    hd = get_handle(torch.cuda.current_device())
    for _ in range(3):
        v5ela.analyze_ragged(crops)
    torch.cuda.synchronize()
    hd.profile_enable(True)
    hd.profile_read(True)
    l0 = hd.launch_count
    for _ in range(20):
        v5ela.analyze_ragged(crops)
    ms, cnt = hd.profile_read(True)
    hd.profile_enable(False)
    assert cnt == 20 and hd.launch_count - l0 == 40                  # fused + finalize per call, nothing per crop
    assert ms / cnt * 1e3 <= 80.0, f"{ms / cnt * 1e3:.1f} us per call"


def test_ragged_argument_errors():
    import ctypes

    import torch
    from v5ela import _abi
    from v5ela.batch import get_handle

    hd = get_handle(torch.cuda.current_device())
    t = torch.zeros((16, 16, 3), dtype=torch.uint8, device="cuda")
    r = torch.zeros((1, 3144), dtype=torch.uint8, device="cuda")
    d = (_abi.FrameDesc * 1)()
    d[0].rgb, d[0].height, d[0].width, d[0].row_stride_bytes = t.data_ptr(), 16, 16, 47      # stride < 3 * width
    with pytest.raises(_abi.V5ElaError):
        hd.analyze_ragged(d, 1, r.data_ptr(), None)
    d[0].row_stride_bytes, d[0].rgb = 48, None
    with pytest.raises(_abi.V5ElaError):
        hd.analyze_ragged(d, 1, r.data_ptr(), None)
    d[0].rgb = t.data_ptr()
    hd.analyze_ragged(d, 1, r.data_ptr(), None)
    hd.analyze_ragged(d, 0, r.data_ptr(), None)                      # empty batch: nothing to do
    torch.cuda.synchronize()

"""CPU: host-side logic — sharding, record parsing/aggregation, the world_size-2 gather over gloo."""
import os
import socket

import numpy as np
import pytest

from oracle import c_oracle
from v5ela.records import RECORD_BYTES, as_records, combine, features
from v5ela.shard import shard_range, shard_videos
from v5ela.synth import gen_batch


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 1000, 2048):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert shard_videos(64, 32, 3, 8) == (24 * 32, 32 * 32)
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_features_and_combine():
    frames = gen_batch(0, 4, 48, 64, seed=1)
    recs, resid = c_oracle.analyze(frames, 90, want_residual=True)
    raw = np.frombuffer(recs.tobytes(), np.uint8).reshape(4, RECORD_BYTES)
    parsed = as_records(raw)
    assert parsed.tobytes() == recs.tobytes()
    f = features(parsed[0], 48 * 64)
    d = resid[0].astype(np.float64)
    assert f["ela_max"] == int(resid[0].max()) and f["ela_scale"] == 255.0 / max(int(resid[0].max()), 1)
    np.testing.assert_allclose(f["ela_mean_rgb"], d.reshape(-1, 3).mean(0), rtol=1e-12)
    np.testing.assert_allclose(f["ela_var_rgb"], d.reshape(-1, 3).var(0), rtol=1e-9)
    agg = combine(parsed)
    assert int(agg["ela_hist"].sum()) == 4 * 48 * 64 * 3
    assert agg["ela_sum"].tolist() == parsed["ela_sum"].sum(0).tolist()
    assert int(agg["tex_maxabs"]) == int(parsed["tex_maxabs"].max())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist

    from v5ela.shard import gather_records, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    local = torch.zeros((hi - lo, RECORD_BYTES), dtype=torch.uint8)
    for i in range(lo, hi):
        local[i - lo] = torch.arange(RECORD_BYTES, dtype=torch.int64).add(i).remainder(251).to(torch.uint8)
    got = gather_records(local, total, dst=0)
    if rank == 0:
        ok = got.shape == (total, RECORD_BYTES)
        for i in range(total):
            ok = ok and bool((got[i] == torch.arange(RECORD_BYTES, dtype=torch.int64).add(i).remainder(251).to(torch.uint8)).all())
        q.put(ok)
    else:
        q.put(got is None)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_records_world2_gloo(total):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert results == [True, True]


def test_log_magnitude_algorithm_error_bound():
    """csrc/v5ela_fft.cuh log_magnitude20 — 20 ln(sqrt(s) + 1) through a 128-entry reciprocal / logarithm table and a degree-6
    polynomial — restated in NumPy (no FMA, so slightly pessimistic): the absolute error against the library functions stays below
    the 2e-13 the header claims, over the whole range the spectrum path can produce (|F|^2 < (255 H W)^2 at 4K) and around 0 and 1.
    The GPU tests (tests/test_gpu_spectrum.py: <= 1 grey level, byte-identical goldens) are the gate for the CUDA code itself."""
    import numpy as np

    i = np.arange(128)
    inv_c = 1.0 / (1.0 + (i + 0.5) / 128.0)
    neg_ln = -np.log(inv_c)

    def log_magnitude20(s):
        s = np.maximum(s, 1e-300)
        m = s * (1.0 / np.sqrt(s))
        b = (m + 1.0).view(np.int64)
        e = (b >> 52) - 1023
        k = (b >> 45) & 127
        mant = ((b & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000).view(np.float64)
        q = mant * inv_c[k] - 1.0
        assert np.abs(q).max() <= 2.0 ** -8 + 1e-12
        p = q * (-1.0 / 6.0) + 0.2
        p = p * q - 0.25
        p = p * q + 1.0 / 3.0
        p = p * q - 0.5
        p = p * q + 1.0
        return 20.0 * np.abs((e * 0.6931471805599453 + neg_ln[k]) + p * q)

    rng = np.random.default_rng(3)
    s = np.concatenate([10.0 ** rng.uniform(-12, 2 * np.log10(255.0 * 2160 * 3840), 1_000_000), rng.uniform(0.0, 4.0, 100_000),
                        np.array([0.0, 1e-310, 1.0, (255.0 * 2160 * 3840) ** 2])])
    ref = 20.0 * np.log(np.sqrt(s) + 1.0)
    got = log_magnitude20(s)
    assert got.min() >= 0.0 and np.abs(got - ref).max() < 2e-13
    assert log_magnitude20(np.array([0.0]))[0] < 1e-15              # |F| = 0: 20 ln 1, the two halves cancel to rounding

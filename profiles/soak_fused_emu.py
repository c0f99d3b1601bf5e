"""CPU soak of the fused kernel's device code (tests/emu build, threads emulated per barrier phase) in the instantiations the
host picks for records-only calls on widths that are a multiple of 16 (paired rows) and for everything else, and with the
texture histogram: random sizes, segment lengths, qualities and contents against the C oracle. Development aid.
usage: python profiles/soak_fused_emu.py [cases] [seed] [emulator build: 2cta | 2cta_mma | 2cta_mma_np2 | 3cta | 3cta_mma]
The builds are the ones tests/test_kernel_emu.py compiles (run it once first). Every case also goes through the small-batch
decomposition (short segments, narrow strips) when `seg` is 0, and the range counters of the block stage must stay at zero."""
import ctypes, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "fake-video-detection-engine_b200"))
from oracle import c_oracle, pil_oracle
from v5ela.records import RECORD_DTYPE

BUILD = sys.argv[3] if len(sys.argv) > 3 else "2cta"
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", f"libv5ela_emu_{BUILD}.so"))
lib.v5emu_range_violations.restype = ctypes.c_longlong
u8p = ctypes.POINTER(ctypes.c_uint8)
lib.v5emu_analyze.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                              ctypes.c_void_p, u8p, ctypes.c_int, ctypes.c_void_p]


def emu(frame, q, seg, want_residual, want_tex):
    h, w, _ = frame.shape
    recs = np.zeros(1, RECORD_DTYPE)
    res = np.zeros((h, w, 3), np.uint8) if want_residual else None
    th = np.zeros(256, np.uint32) if want_tex else None
    rc = lib.v5emu_analyze(frame.ctypes.data_as(u8p), 1, h, w, h * w * 3, w * 3, q, recs.ctypes.data_as(ctypes.c_void_p),
                           res.ctypes.data_as(u8p) if want_residual else None, seg,
                           th.ctypes.data_as(ctypes.c_void_p) if want_tex else None)
    assert rc == 0
    return recs[0], res, th


def main(cases=400, seed=1):
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(cases):
        h = int(rng.integers(1, 150))
        w = 16 * int(rng.integers(1, 40)) if it % 4 else int(rng.integers(1, 600))
        q = int(rng.integers(1, 101))
        seg = int(rng.choice([0, 1, 2, 3, 5]))
        kind = it % 5
        if kind == 0:
            f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == 1:
            f = np.full((h, w, 3), int(rng.integers(256)), np.uint8)
        elif kind == 2:
            f = np.where(rng.integers(0, 2, (h, w, 3)) > 0, 255, 0).astype(np.uint8)
        elif kind == 3:
            yy, xx = np.mgrid[0:h, 0:w]
            f = np.stack([(xx * 3 + yy) % 256, (yy * 5 + xx // 3) % 256, (xx + yy * 2) % 256], -1).astype(np.uint8)
        else:
            f = np.repeat(rng.integers(0, 256, (h, w, 1), dtype=np.uint8), 3, axis=2)
        f = np.ascontiguousarray(f)
        o = c_oracle.analyze_frame(f, q)
        lib.v5emu_set_target_items(592 if (seg == 0 and it % 2) else 0)
        mode = it % 3                                            # 0: records only (paired rows when w % 16 == 0), 1: + residual, 2: + tex_hist
        rec, res, th = emu(f, q, seg, mode == 1, mode == 2)
        ok = rec.tobytes() == o["record"].tobytes()
        if mode == 1:
            ok = ok and np.array_equal(res, o["residual"])
        if mode == 2:
            ok = ok and np.array_equal(th, pil_oracle.texture_hist(f))
        if not ok:
            bad += 1
            print("MISMATCH", h, w, q, seg, kind, mode)
    print(f"build {BUILD}: cases {cases} mismatches {bad} range violations {lib.v5emu_range_violations()}")


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]))

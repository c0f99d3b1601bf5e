"""Small end-to-end case for compute-sanitizer: multi-strip, multi-segment, ragged edges, residual + enhance + spectrum."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fake-video-detection-engine_b200"))
import numpy as np
from v5ela import host
from v5ela.synth import gen_batch
from oracle import c_oracle
for (n, h, w) in [(2, 100, 1000), (1, 300, 497), (3, 33, 47), (1, 272, 496)]:
    frames = gen_batch(0, n, h, w, 1)
    recs, resid, enh = host.analyze_frames_host(frames, want_residual=True, want_enhanced=True)
    orecs, oresid = c_oracle.analyze(frames, 90, want_residual=True)
    assert np.array_equal(resid, oresid) and recs.tobytes() == orecs.tobytes(), (n, h, w)
spec = host.spectrum_host(gen_batch(0, 1, 65, 99, 0)[0, ..., 1])
print("sanitize case ok", spec.shape)

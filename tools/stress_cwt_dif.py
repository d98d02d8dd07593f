"""Randomised cross-check of the two / four-warps-per-row CWT kernels against the generic kernel.

    python tools/stress_cwt_dif.py [cases] [seed]

Random row lengths in (1024, 4096], batch sizes, scale grids, f0 and cone masks; FP32 gate 1e-4 norm-wise
(NaN patterns of the masked planes must be identical).  Prints the worst error per FFT length.
"""
import os
import sys
from pathlib import Path

import numpy as np

os.environ["WTB_CWT_MIN_BATCH"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wavelet_transformer_b200 import _shim  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
_shim.init(0)
worst = {2048: 0.0, 4096: 0.0}
for c in range(cases):
    n0 = int(rng.integers(1025, 4097))
    nfft = 2048 if n0 <= 2048 else 4096
    batch = int(rng.choice([1, 3, 12, 13, 37, 149, 300, 611]))
    dt = float(rng.choice([1 / 12, 0.25, 1.0]))
    dj = float(rng.choice([1 / 2, 1 / 4, 1 / 8, 1 / 12]))
    s0 = dt * float(rng.choice([1.0, 2.0, 4.0]))
    f0 = float(rng.choice([5.5, 6.0, 6.0, 9.0]))
    jmax = int(np.floor(np.log2(n0 * dt / s0) / dj))
    J = int(rng.integers(max(1, jmax // 3), min(jmax, 127) + 1))
    coi = bool(rng.integers(0, 2))
    x = rng.standard_normal((batch, n0)) * float(rng.choice([1e-3, 1.0, 1e3])) + 0.05 * rng.standard_normal((batch, n0)).cumsum(axis=1)
    got, _ = _shim.cwt_morlet(x, dt, dj, s0, J, f0, f64=False, coi_mask=coi)
    ref, _ = _shim.cwt_morlet(x, dt, dj, s0, J, f0, f64=False, coi_mask=coi, generic_only=True)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), (c, n0, batch, "mask differs")
    m = ~np.isnan(ref)
    for b in range(batch):
        mb = m[b]
        if not mb.any():
            continue
        scale = np.abs(ref[b][mb]).max()
        err = float((np.abs(got[b][mb] - ref[b][mb]) / (1e-4 * np.abs(ref[b][mb]) + 1e-4 * scale)).max())
        worst[nfft] = max(worst[nfft], err)
        assert err <= 1.0, (c, n0, batch, dj, J, f0, coi, b, err)
    print(f"case {c}: n0={n0} batch={batch} dj={dj:.3f} J={J} f0={f0} coi={coi} ok", flush=True)
print("worst error / gate:", worst)

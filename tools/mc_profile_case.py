"""Small Monte-Carlo significance run for profiling (ncu launch list / --set full captures).

    python tools/mc_profile_case.py [realisations] [repeats]

Runs `wtb_wct_mc_hist` (device RNG, FP32, cfg5 shape) `repeats` times on `realisations` pairs and
prints the rate; under ncu each repeat shows the five kernels of the pipeline once per chunk.
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wavelet_transformer_b200 import _shim  # noqa: E402

DT = 1 / 12
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
_shim.init(0)
for r in range(reps):
    t0 = time.perf_counter()
    h = _shim.wct_mc_hist(0.989, 0.966, DT, 1 / 8, 2 * DT, 65, mc_first=r * n, mc_count=n, seed=2024, f64=False)
    dt = time.perf_counter() - t0
    print(f"rep {r}: {n} realisations in {dt * 1e3:.1f} ms -> {n / dt:.0f} pairs/s, counts {int(h.sum())}")

#!/usr/bin/env python
"""Latency of the FP32 fused CWT+power call on small device-resident batches: the shipped
dispatch (warp kernels with each series' rows split over up to 16 warps, generic kernel below
their break-even) against the generic kernel alone.

    python tools/bench_small_batches.py
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from wavelet_transformer_b200 import _shim

    _shim.init(0)
    dev = torch.device("cuda", 0)
    dt = 1 / 12

    def run(n0, dj, J, batch, generic):
        x = torch.randn((batch, n0), dtype=torch.float32, device=dev)
        out = torch.empty((batch, J + 1, n0), dtype=torch.float32, device=dev)

        def step():
            _shim.cwt_power_device(x.data_ptr(), batch, n0, dt, dj, 2 * dt, J, 6.0, out.data_ptr(), f64=False,
                                   generic_only=generic)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    for n0, dj, J in ((1024, 1 / 12, 119), (565, 1 / 8, 65), (400, 1 / 12, 91), (1346, 1 / 12, 84)):
        for batch in (1, 4, 16, 64, 256, 1024):
            print(json.dumps({"n0": n0, "scales": J + 1, "batch": batch,
                              "dispatch_ms": round(run(n0, dj, J, batch, False), 4),
                              "generic_ms": round(run(n0, dj, J, batch, True), 4)}), flush=True)


if __name__ == "__main__":
    main()

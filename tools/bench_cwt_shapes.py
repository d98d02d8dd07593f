#!/usr/bin/env python
"""Device-resident throughput of the fused CWT+power paths by series shape, against the roofline north_star
names: the slower of FFT flops at the nominal FP32 peak and coefficient bytes at the measured HBM bandwidth.

    python tools/bench_cwt_shapes.py

Unit = one series; algorithmic bytes 4*n0*(1 + S) (FP32 series in, power plane out), algorithmic flop
5 N log2 N + S (5 N log2 N + 5 N) with N the FFT length.  Shapes: BASELINE
cfg1 (1346 samples -> nfft 2048, 85 scales), full 2048, cfg4 (1024 x 120) and the nfft-4096 rows;
each also through the generic Stockham kernel for reference.
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
FP32_PEAK_TF = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal non-tensor FP32
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from wavelet_transformer_b200 import _shim

    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="substring of the shape label to run (default: all)")
    ap.add_argument("--no-generic", action="store_true")
    args = ap.parse_args()
    _shim.init(0)
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    dev = torch.device("cuda", 0)
    dt = 1 / 12
    shapes = [("cfg1 shape", 1346, 1 / 12, 84, 20000), ("nfft 2048 full", 2048, 1 / 12, 84, 20000),
              ("odd rows", 1345, 1 / 12, 84, 20000), ("cfg4", 1024, 1 / 12, 119, 20000),
              ("nfft 4096", 3351, 1 / 8, 65, 4000), ("nfft 512", 400, 1 / 12, 91, 40000)]
    for label, n0, dj, J, batch in shapes:
        if args.only and args.only not in label:
            continue
        S = J + 1
        x = torch.randn((batch, n0), dtype=torch.float32, device=dev)
        out = torch.empty((batch, S, n0), dtype=torch.float32, device=dev)
        for generic in ((False,) if args.no_generic else (False, True)):
            nb = batch if not generic else batch // 4

            def step():
                _shim.cwt_power_device(x.data_ptr(), nb, n0, dt, dj, 2 * dt, J, 6.0, out.data_ptr(), f64=False,
                                       generic_only=generic)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                step()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3 / 5
            gbs = 4.0 * n0 * (1 + S) * nb / t / 1e9
            nfft = 1 << (n0 - 1).bit_length()
            lg = nfft.bit_length() - 1
            flop = 5.0 * nfft * lg + S * (5.0 * nfft * lg + 5.0 * nfft)
            tfs = flop * nb / t / 1e12
            t_hbm, t_fp32 = 4.0 * n0 * (1 + S) / (peak * 1e9), flop / (FP32_PEAK_TF * 1e12)
            print(json.dumps({"shape": label, "n0": n0, "nfft": nfft, "scales": S, "batch": nb,
                              "kernel": "generic" if generic else "fast", "ms": t * 1e3, "coeff_per_s": nb * S * n0 / t,
                              "achieved_GBs": gbs, "frac_hbm": gbs / peak, "achieved_TFLOPs": tfs,
                              "frac_fp32_nominal": tfs / FP32_PEAK_TF, "bound": "fp32" if t_fp32 > t_hbm else "hbm",
                              "frac_roofline": max(t_hbm, t_fp32) * nb / t}))
        del x, out


if __name__ == "__main__":
    main()

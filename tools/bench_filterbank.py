#!/usr/bin/env python
"""Device-resident throughput of the MODWT / DWT filterbank kernels against the HBM roofline.

    python tools/bench_filterbank.py [--batch 100000]

Unit = one series.  Algorithmic bytes: MODWT analysis sz*N*(1 + J+1), synthesis
sz*N*(J+1 + 1), MRA sz*N*2*(J+1), wavedec / waverec sz*(N + sum(lens)).  Peak = MEASURED_PEAKS.json hbm_gbs.
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from wavelet_transformer_b200 import _shim
    from wavelet_transformer_b200 import pywt_compat as pywt

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, nargs="*", default=[1333, 1024, 4096])
    ap.add_argument("--dtype", choices=["f64", "f32", "both"], default="both")
    args = ap.parse_args()
    _shim.init(0)
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    dev = torch.device("cuda", 0)
    w = pywt.Wavelet("sym4")
    st = torch.cuda.current_stream().cuda_stream
    J = 6

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / args.steps

    for n in args.n:
        for dtype, f64, sz in ((torch.float64, True, 8), (torch.float32, False, 4)):
            if args.dtype != "both" and f64 != (args.dtype == "f64"):
                continue
            B = args.batch if n <= 2048 else args.batch // 4
            x = torch.randn((B, n), dtype=dtype, device=dev)
            out = torch.empty((B, J + 1, n), dtype=dtype, device=dev)
            rec = torch.empty((B, n), dtype=dtype, device=dev)
            t = timed(lambda: _shim.modwt_device(x.data_ptr(), B, n, w.dec_lo, w.dec_hi, J, out.data_ptr(), f64=f64, stream=st))
            by = sz * n * (1 + J + 1) * B
            print(json.dumps({"kernel": "k_modwt", "n": n, "J": J, "dtype": str(dtype), "batch": B, "ms": t * 1e3,
                              "coeff_per_s": B * (J + 1) * n / t, "achieved_GBs": by / t / 1e9, "frac_hbm": by / t / 1e9 / peak}))
            t = timed(lambda: _shim.imodwt_device(out.data_ptr(), B, n, w.dec_lo, w.dec_hi, J, rec.data_ptr(), f64=f64, stream=st))
            print(json.dumps({"kernel": "k_imodwt", "n": n, "J": J, "dtype": str(dtype), "batch": B, "ms": t * 1e3,
                              "achieved_GBs": by / t / 1e9, "frac_hbm": by / t / 1e9 / peak,
                              "max_abs_err": float((rec - x).abs().max())}))
            mra = torch.empty_like(out)
            t = timed(lambda: _shim.modwtmra_taps_device(out.data_ptr(), B, n, w.dec_lo, w.dec_hi, J, mra.data_ptr(), f64=f64, stream=st))
            by3 = sz * n * 2 * (J + 1) * B
            print(json.dumps({"kernel": "k_mra", "n": n, "J": J, "dtype": str(dtype), "batch": B, "ms": t * 1e3,
                              "achieved_GBs": by3 / t / 1e9, "frac_hbm": by3 / t / 1e9 / peak,
                              "max_abs_err": float((mra.sum(dim=1) - x).abs().max())}))
            del mra
            level = pywt.dwt_max_level(n, 8)
            lens = _shim.dwt_coeff_lens(n, 8, level)
            packed = torch.empty((B, int(lens.sum())), dtype=dtype, device=dev)
            t = timed(lambda: _shim.wavedec_device(x.data_ptr(), B, n, w.dec_lo, w.dec_hi, level, packed.data_ptr(), f64=f64, stream=st))
            by2 = sz * (n + int(lens.sum())) * B
            print(json.dumps({"kernel": "k_wavedec", "n": n, "level": int(level), "dtype": str(dtype), "batch": B,
                              "ms": t * 1e3, "achieved_GBs": by2 / t / 1e9, "frac_hbm": by2 / t / 1e9 / peak}))
            nrec = _shim.waverec_len(lens, 8)
            xr = torch.empty((B, nrec), dtype=dtype, device=dev)
            t = timed(lambda: _shim.waverec_device(packed.data_ptr(), B, lens, w.rec_lo, w.rec_hi, xr.data_ptr(), f64=f64, stream=st))
            by4 = sz * (nrec + int(lens.sum())) * B
            print(json.dumps({"kernel": "k_waverec", "n": n, "level": int(level), "dtype": str(dtype), "batch": B,
                              "ms": t * 1e3, "achieved_GBs": by4 / t / 1e9, "frac_hbm": by4 / t / 1e9 / peak,
                              "max_abs_err": float((xr[:, :n] - x).abs().max()) if n % 2 == 0 else None}))
            del x, out, rec, packed, xr


if __name__ == "__main__":
    main()

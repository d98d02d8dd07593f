#!/usr/bin/env python
"""Latency of the reference's three real-data configurations (BASELINE.json configs 1-3)
through the drop-in entry points, next to the CPU oracle on the same inputs.

    python tools/bench_real_data.py        # on a B200 box; prints one JSON line per config

These are single-series calls: kernel time is microseconds and the figures are dominated
by launch latency and host<->device copies ("I/O dominates").
"""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import modwt_oracle as mo  # noqa: E402
from oracle import pycwt_oracle as po  # noqa: E402
from wavelet_transformer_b200 import _shim  # noqa: E402
from wavelet_transformer_b200 import pycwt_compat as wavelet  # noqa: E402
from wavelet_transformer_b200.api import modwt as gmodwt  # noqa: E402

DT = 1 / 12


def best_of(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    _shim.init(0)
    s = dict(np.load(ROOT / "tests" / "golden" / "sample_series.npz"))
    out = []
    # cfg1: Morlet CWT (dj=1/12, s0=2dt, J=84) of cpi (n0=1346 -> 2048), power plane 85 x 1346
    x = (s["cpi_value"] - s["cpi_value"].mean()) / s["cpi_value"].std()
    for prec in ("fp64", "fp32"):
        _shim.set_precision(prec)
        t_gpu = best_of(lambda: _shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, 84))
        t_cpu = best_of(lambda: np.abs(po.cwt(x, DT, 1 / 12, 2 * DT, 84)[0]) ** 2, 3)
        out.append({"config": "cfg1 CWT cpi.csv 85x1346", "precision": prec, "gpu_ms": 1e3 * t_gpu,
                    "cpu_oracle_ms": 1e3 * t_cpu, "coeff_per_s_gpu": 85 * 1346 / t_gpu})
    _shim.set_precision("fp64")
    # cfg2: MODWT LA8 (sym4) J=6 + MRA of inflation and expectation
    for name in ("inflation", "expectation"):
        v = s[f"{name}_value"]
        t_gpu = best_of(lambda: gmodwt.modwtmra(gmodwt.modwt(v, "sym4", 6), "sym4"))
        t_cpu = best_of(lambda: mo.modwtmra(mo.modwt(v, "sym4", 6), "sym4"), 3)
        out.append({"config": f"cfg2 MODWT+MRA sym4 J=6 {name}.csv (N={v.size})", "precision": "fp64",
                    "gpu_ms": 1e3 * t_gpu, "cpu_oracle_ms": 1e3 * t_cpu})
    # cfg3: WCT expectation vs diff-log CPI (n0=565, 66 scales) + 300-realisation Monte Carlo
    y1 = (100 * np.diff(np.log(s["cpi_value"])))[-565:]
    y2 = s["expectation_value"]
    t_gpu = best_of(lambda: wavelet.wct(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1, sig=False), 5)
    t_cpu = best_of(lambda: po.wct(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1, sig=False), 3)
    out.append({"config": "cfg3 WCT 66x565, no significance", "precision": "fp64", "gpu_ms": 1e3 * t_gpu,
                "cpu_oracle_ms": 1e3 * t_cpu})
    a1, a2 = po.ar1(y1)[0], po.ar1(y2)[0]
    for prec in ("fp64", "fp32"):
        _shim.set_precision(prec)
        t_gpu = best_of(lambda: wavelet.wct_significance(a1, a2, DT, 1 / 8, 2 * DT, 65, mc_count=300, cache=False), 3)
        t0 = time.perf_counter()
        po.wct_significance(a1, a2, DT, 1 / 8, 2 * DT, 65, mc_count=6, rng=np.random.default_rng(0), faithful_loop=True)
        t_cpu = (time.perf_counter() - t0) * 50
        out.append({"config": "cfg3 wct_significance 300 realisations (N=3351->4096, 66 scales)", "precision": prec,
                    "gpu_ms": 1e3 * t_gpu, "cpu_oracle_ms": 1e3 * t_cpu,
                    "cpu_note": "6 realisations timed (pycwt-style Python histogram loop), scaled x50",
                    "surrogates_per_s_gpu": 300 / t_gpu})
    # the reference's own entry points end to end (dataclasses in, dataclasses out), default precision
    import os
    import shutil
    import tempfile
    from src import cwt as rcwt, dwt as rdwt, wct as rwct, xwt as rxwt
    from wavelet_transformer_b200 import pywt_compat as pywt
    _shim.set_precision("fp64")
    t = np.arange("1978-01", "2025-02", dtype="datetime64[M]")[:565]
    cache = tempfile.mkdtemp(prefix="wtb_cache_")
    os.environ["WTB_CACHE_DIR"] = cache

    def fresh_wct():
        shutil.rmtree(cache, ignore_errors=True)             # no cached significance: the Monte Carlo runs
        return rwct.run_wct(rwct.DataForWCT(y1, y2, rwct.MOTHER, DT, 1 / 8, 2 * DT, [1, 2, 4, 8, 16]))

    runs = {
        "run_cwt (diff-log cpi, 85 x 1345, significance)": lambda: rcwt.run_cwt(
            rcwt.DataForCWT(np.arange("1913-02", "2025-03", dtype="datetime64[M]")[:1345],
                            100 * np.diff(np.log(s["cpi_value"])), rcwt.MOTHER, DT, 1 / 12, 2 * DT, 7 * 12)),
        "run_wct (66 x 565) + 300-realisation Monte Carlo significance, cold cache": fresh_wct,
        "run_xwt (66 x 565)": lambda: rxwt.run_xwt(rxwt.DataForXWT(y1, y2, rxwt.MOTHER, DT, 1 / 8, 2 * DT, [1, 2, 4, 8, 16])),
        "run_dwt db4 + smooth_signal (N=565)": lambda: rdwt.run_dwt(rdwt.DataForDWT(y2, pywt.Wavelet("db4"))).smooth_signal(
            y2, pywt.Wavelet("db4")),
    }
    for name, fn in runs.items():
        try:
            out.append({"config": "entry point: " + name, "precision": "fp64 (Monte Carlo fp32)", "gpu_ms": 1e3 * best_of(fn, 5)})
        except Exception as exc:  # a signature drift would show here rather than abort the listing
            out.append({"config": "entry point: " + name, "error": repr(exc)})
    shutil.rmtree(cache, ignore_errors=True)
    # One call, every GPU of the box: pycwt_compat.wct_significance (what run_wct reaches) with the
    # BASELINE cfg5 count, the library's own worker pool spreading the realisations (wtb_init_multi).
    if "--mc-scaling" in sys.argv:
        _shim.set_precision("fp32")
        ref = None
        n = 1
        while n <= _shim.device_count():
            wavelet.wct_significance(a1, a2, DT, 1 / 8, 2 * DT, 65, mc_count=2000, cache=False, n_gpus=n)   # pool + scratch warm-up
            t_gpu = best_of(lambda: wavelet.wct_significance(a1, a2, DT, 1 / 8, 2 * DT, 65, mc_count=100_000, cache=False,
                                                             seed=2024, n_gpus=n), 3)
            sig = wavelet.wct_significance(a1, a2, DT, 1 / 8, 2 * DT, 65, mc_count=100_000, cache=False, seed=2024, n_gpus=n)
            ref = sig if ref is None else ref
            out.append({"config": f"one call: wct_significance mc_count=100000 on {n} GPU(s)", "precision": "fp32",
                        "gpu_ms": 1e3 * t_gpu, "surrogates_per_s_gpu": 100_000 / t_gpu,
                        "same_thresholds_as_one_gpu": bool(np.array_equal(sig, ref, equal_nan=True))})
            n *= 2
        _shim.init_multi(1)
    for line in out:
        print(json.dumps(line))


if __name__ == "__main__":
    main()

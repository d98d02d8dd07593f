"""Regenerate the committed golden fixtures (run in the BUILD container only).

    python tests/golden/make_golden.py

Reads /root/reference (absent on the GPU box; tests only read the outputs):

* ``sample_series.npz``  -- the ``value`` columns (and dates as int64 days) of
  sample_data/{cpi,inflation,expectation}.csv plus the inflation/expectation
  left-merge the WCT demo uses.
* ``modwt_reference.npz`` -- outputs of the reference's OWN MODWT arithmetic
  (src/modwt.py:56-194, 232-251), extracted by AST so its plotting / network
  imports are never executed, with ``pywt.Wavelet`` stubbed by the tap tables.
* ``helpers_reference.npz`` -- outputs of the reference's
  ``standardize_series`` / ``normalize_xwt_results`` / ``calculate_phase_difference``
  (src/utils/wavelet_helpers.py:22-78, src/wct.py:143-158).

No reference SOURCE is copied into the repo: only numeric inputs/outputs.
"""

from __future__ import annotations

import ast
import logging
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

from oracle.pywt_oracle import Wavelet  # noqa: E402  (tap tables for the pywt stub)


def _extract(path: Path, names: set[str], namespace: dict):
    tree = ast.parse(path.read_text())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = names - {n.name for n in body}
    assert not missing, f"{path}: missing {missing}"
    code = compile(ast.Module(body=body, type_ignores=[]), str(path), "exec")
    exec(code, namespace)
    return namespace


def sample_series():
    out = {}
    frames = {}
    for name in ("cpi", "inflation", "expectation"):
        df = pd.read_csv(REF / "sample_data" / f"{name}.csv", parse_dates=["date"])
        frames[name] = df
        out[f"{name}_value"] = df["value"].to_numpy(dtype=float)
        out[f"{name}_days"] = df["date"].to_numpy().astype("datetime64[D]").astype(np.int64)
    merged = frames["expectation"].merge(frames["inflation"], on="date", how="left",
                                         suffixes=("_exp", "_inf")).dropna()
    out["pair_expectation"] = merged["value_exp"].to_numpy(dtype=float)
    out["pair_inflation"] = merged["value_inf"].to_numpy(dtype=float)
    out["pair_days"] = merged["date"].to_numpy().astype("datetime64[D]").astype(np.int64)
    np.savez_compressed(HERE / "sample_series.npz", **out)
    return out


def modwt_reference(series):
    from scipy.ndimage import convolve1d

    pywt_stub = types.SimpleNamespace(Wavelet=Wavelet)
    ns = {"np": np, "convolve1d": convolve1d, "pywt": pywt_stub, "print": lambda *a, **k: None}
    _extract(REF / "src" / "modwt.py",
             {"upArrow_op", "period_list", "circular_convolve_mra", "circular_convolve_d",
              "circular_convolve_s", "modwt", "imodwt", "modwtmra", "smooth_signal"}, ns)
    out = {}
    rng = np.random.default_rng(7)
    cases = {
        "inflation": series["inflation_value"],
        "expectation": series["expectation_value"],
        "short37": rng.standard_normal(37),      # dilated kernel longer than N at J>=4
        "pow2_256": rng.standard_normal(256),
    }
    for cname, x in cases.items():
        for filt in ("db4", "sym4", "haar"):
            J = 6 if x.size > 100 else 4
            w = ns["modwt"](x, filt, J)
            out[f"{cname}|{filt}|J"] = np.int64(J)
            out[f"{cname}|{filt}|x"] = x
            out[f"{cname}|{filt}|modwt"] = w
            out[f"{cname}|{filt}|imodwt"] = ns["imodwt"](w, filt)
            out[f"{cname}|{filt}|mra"] = ns["modwtmra"](w, filt)
            sm = ns["smooth_signal"](w, filt, J)
            out[f"{cname}|{filt}|smooth{J}"] = sm[J]["signal"]
            out[f"{cname}|{filt}|smooth1"] = sm[1]["signal"]
    np.savez_compressed(HERE / "modwt_reference.npz", **out)


def helpers_reference(series):
    ns = {"np": np, "npt": types.SimpleNamespace(NDArray=np.ndarray),
          "logger": logging.getLogger("golden"), "Tuple": tuple}
    _extract(REF / "src" / "utils" / "wavelet_helpers.py",
             {"standardize_series", "normalize_xwt_results", "align_series"}, ns)
    _extract(REF / "src" / "wct.py", {"calculate_phase_difference"}, ns)
    y = series["pair_inflation"]
    out = {
        "y": y,
        "std_detrend": ns["standardize_series"](y, detrend=True),
        "std_mean": ns["standardize_series"](y, detrend=False, remove_mean=True),
        "std_raw": ns["standardize_series"](y, detrend=False, standardize=False),
    }
    rng = np.random.default_rng(11)
    S, n = 9, 40
    xw = rng.standard_normal((S, n)) + 1j * rng.standard_normal((S, n))
    coi = np.abs(rng.standard_normal(n)) + 0.3
    freqs = 1.0 / (0.2 * 2 ** (np.arange(S) / 2))
    signif = np.abs(rng.standard_normal(S)) + 0.5
    period, power, sig95, coi_plot = ns["normalize_xwt_results"](
        n, xw, coi, np.log2(0.25), freqs, signif)
    out.update(xw=xw, coi=coi, freqs=freqs, signif=signif, coi_min=np.log2(0.25),
               nx_period=period, nx_power=power, nx_sig95=sig95, nx_coi_plot=coi_plot)
    phase = rng.uniform(-np.pi, np.pi, (S, n))
    u, v = ns["calculate_phase_difference"](phase)
    out.update(phase=phase, phase_u=u, phase_v=v)
    out["align"] = ns["align_series"](np.arange(10), np.arange(12.0))
    np.savez_compressed(HERE / "helpers_reference.npz", **out)


if __name__ == "__main__":
    s = sample_series()
    modwt_reference(s)
    helpers_reference(s)
    print("golden fixtures written to", HERE)

"""Where does the FP32 coherence error sit?  ("what's weak" 1.)

Runs `wtb_xwt_wct` in FP32 (fast nfft = 4096 kernels where they apply, and the generic kernels)
against the float64 oracle on BASELINE cfg3's pair, the cfg5 surrogate shape and the test
batch of tests/test_gpu_wct.py, and prints for every case

  * max / p99.9 / p99 / mean of |wct - ref| and the BASELINE.md bound 1e-4 |ref| + 1e-4 max|ref|,
  * the error split by the size of the denominator S1*S2 relative to its row median
    (the smoothed spectra come out of FP32 FFTs whose error is relative to the ROW, so small
    denominators are where a ratio can lose digits),
  * the location of the worst sample.

    python tools/wct_fp32_error.py [--out profiles/r2_wct_fp32_error.jsonl]

Test infrastructure: imports oracle/ as the checker.
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import pycwt_oracle as po  # noqa: E402
from wavelet_transformer_b200 import _shim  # noqa: E402

DT = 1 / 12


def _norm(y):
    return (y - y.mean()) / y.std()


def reference(y1, y2, dt, dj, s0, J):
    wav = po.Morlet()
    W1, sj, *_ = po.cwt(y1, dt, dj, s0, J, wav)
    W2 = po.cwt(y2, dt, dj, s0, J, wav)[0]
    S1, S2, S12, _ = po.smoothed_spectra(W1, W2, sj, dt, dj, wav)
    return np.abs(S12) ** 2 / (S1 * S2), S1 * S2


def analyse(tag, got, ref, den):
    err = np.abs(got.astype(np.float64) - ref)
    bound = 1e-4 * np.abs(ref) + 1e-4 * np.abs(ref).max()
    rel_den = den / np.median(den, axis=1, keepdims=True)
    i = np.unravel_index(err.argmax(), err.shape)
    rec = {
        "case": tag, "shape": list(ref.shape),
        "max": float(err.max()), "p999": float(np.percentile(err, 99.9)), "p99": float(np.percentile(err, 99)),
        "mean": float(err.mean()), "violations_of_1e-4_gate": int((err > bound).sum()),
        "worst_at": [int(i[0]), int(i[1])], "worst_den_over_row_median": float(rel_den[i]),
        "by_denominator": {},
    }
    for lo, hi in [(0, 1e-4), (1e-4, 1e-3), (1e-3, 1e-2), (1e-2, 1e-1), (1e-1, np.inf)]:
        m = (rel_den >= lo) & (rel_den < hi)
        if m.any():
            rec["by_denominator"][f"[{lo:g},{hi:g})"] = {"n": int(m.sum()), "max": float(err[m].max()),
                                                          "mean": float(err[m].mean())}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    _shim.init(0)
    cases = []
    gold = dict(np.load(ROOT / "tests" / "golden" / "sample_series.npz"))
    cases.append(("cfg3 pair n0=565 (nfft 1024)", _norm(gold["pair_inflation"]), _norm(gold["pair_expectation"]),
                  1 / 8, -1))
    rng = np.random.default_rng(41)
    for k in range(2):
        cases.append((f"cfg5 surrogate n0=3351 (nfft 4096) #{k}", _norm(po.rednoise(3351, 0.989, 1, rng)),
                      _norm(po.rednoise(3351, 0.966, 1, rng)), 1 / 8, 65))
    rng = np.random.default_rng(17)
    a = rng.standard_normal((4, 700)).cumsum(axis=1)
    b = a * 0.3 + rng.standard_normal((4, 700)) * 3
    for dj in (1 / 8, 1 / 12, 1 / 4):
        for k in range(4):
            cases.append((f"test batch n0=700 dj=1/{round(1 / dj)} #{k}", _norm(a[k]), _norm(b[k]), dj, -1))
    rng = np.random.default_rng(5)
    cases.append(("white pair n0=4096", _norm(rng.standard_normal(4096)), _norm(rng.standard_normal(4096)), 1 / 8, 65))
    out = []
    for tag, y1, y2, dj, J in cases:
        ref, den = reference(y1, y2, DT, dj, 2 * DT, J)
        for mode, kw in (("fp32 dispatch", {}), ("fp32 generic", {"generic_only": True})):
            got, _, _ = _shim.xwt_wct(y1, y2, DT, dj, 2 * DT, J, f64=False, want_phase=False, **kw)
            rec = analyse(f"{tag} | {mode}", got, ref, den)
            out.append(rec)
            print(json.dumps(rec))
    worst = max(r["max"] for r in out)
    print(f"# worst max error over {len(out)} runs: {worst:.3e}; gate violations: "
          f"{sum(r['violations_of_1e-4_gate'] for r in out)}")
    if args.out:
        with open(args.out, "w") as f:
            for r in out:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()

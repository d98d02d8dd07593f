"""The MODWT oracle against the reference's own arithmetic (fixtures made by
tests/golden/make_golden.py from /root/reference/src/modwt.py) plus invariants."""

import numpy as np
import pytest

from oracle import modwt_oracle as mo


def _cases(g):
    return sorted({k.rsplit("|", 1)[0] for k in g})


def test_oracle_matches_reference_code(modwt_golden):
    g = modwt_golden
    assert len(_cases(g)) == 12
    for case in _cases(g):
        filt = case.split("|")[1]
        x, J = g[f"{case}|x"], int(g[f"{case}|J"])
        tol = 1e-12 * max(1.0, np.abs(x).max())
        w = g[f"{case}|modwt"]
        assert np.abs(mo.modwt(x, filt, J) - w).max() < tol
        assert np.abs(mo.imodwt(w, filt) - g[f"{case}|imodwt"]).max() < tol
        assert np.abs(mo.modwtmra(w, filt) - g[f"{case}|mra"]).max() < tol
        sm = mo.smooth_signal(w, filt, J)
        assert np.abs(sm[J]["signal"] - g[f"{case}|smooth{J}"]).max() < tol
        assert np.abs(sm[1]["signal"] - g[f"{case}|smooth1"]).max() < tol


@pytest.mark.parametrize("filt", ["db4", "sym4", "haar"])
def test_invariants(series, filt):
    for name in ("inflation_value", "expectation_value"):
        x = series[name]
        w = mo.modwt(x, filt, 6)
        assert w.shape == (7, x.size)
        assert np.abs(mo.imodwt(w, filt) - x).max() < 1e-9
        assert np.abs(mo.modwtmra(w, filt).sum(axis=0) - x).max() < 1e-9
        assert (w ** 2).sum() == pytest.approx((x ** 2).sum(), rel=1e-10)


def test_impulse_support_is_causal_and_circular():
    x = np.zeros(128); x[10] = 1
    w = mo.modwt(x, "db4", 3)
    for j, last in ((0, 17), (1, 31), (2, 59)):
        nz = np.nonzero(np.abs(w[j]) > 1e-14)[0]
        assert (nz.min(), nz.max()) == (10, last)
    x = np.zeros(16); x[15] = 1
    assert np.nonzero(np.abs(mo.modwt(x, "haar", 1)[0]) > 0)[0].tolist() == [0, 15]

"""Pins for the statsmodels-OLS restatement (oracle/ols_oracle.py)."""

import numpy as np
from scipy import stats

from oracle import ols_oracle


def test_ols_oracle_matches_linregress_and_lstsq():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(200)
    y = 0.3 + 1.7 * x + 0.5 * rng.standard_normal(200)
    fit = ols_oracle.ols(y, x)
    lr = stats.linregress(x, y)
    assert np.allclose(fit["params"], [lr.intercept, lr.slope], rtol=1e-12)
    assert np.allclose(fit["bse"], [lr.intercept_stderr, lr.stderr], rtol=1e-10)
    assert np.isclose(fit["pvalues"][1], lr.pvalue, rtol=1e-8)
    assert np.isclose(fit["rsquared"], lr.rvalue ** 2, rtol=1e-12)
    nc = ols_oracle.ols(y, x, add_constant=False)
    assert np.allclose(nc["params"], np.linalg.lstsq(x[:, None], y, rcond=None)[0], rtol=1e-12)
    assert np.isclose(nc["rsquared"], 1 - nc["ssr"] / (y @ y))          # uncentred without a constant
    assert nc["df_resid"] == 199 and fit["df_resid"] == 198


def _stats_like_kernel(x, y, add_constant):
    """What wtb_rowwise_ols writes per row, computed with NumPy (host logic test: no GPU)."""
    n = x.size
    mx, my = (x.mean(), y.mean()) if add_constant else (0.0, 0.0)
    dx, dy = x - mx, y - my
    sxx, sxy, syy = dx @ dx, dx @ dy, dy @ dy
    slope = sxy / sxx
    return [n, (my - slope * mx) if add_constant else 0.0, slope, max(syy - slope * sxy, 0.0), syy, sxx, x.mean(), y.mean()]


def test_fits_from_stats_host_closed_forms():
    """Standard errors, t, p, R^2 and adjusted R^2 formed on the host from the kernel's eight
    numbers equal the statsmodels-OLS restatement, with and without a constant."""
    from wavelet_transformer_b200.api import regression as reg
    rng = np.random.default_rng(3)
    for add_constant in (True, False):
        rows = []
        pairs = []
        for _ in range(4):
            x = rng.standard_normal(120) + 2.0
            y = -1.0 + 0.4 * x + rng.standard_normal(120)
            rows.append(_stats_like_kernel(x, y, add_constant))
            pairs.append((x, y))
        fits = reg.fits_from_stats(np.array(rows), add_constant)
        for fit, (x, y) in zip(fits, pairs):
            ref = ols_oracle.ols(y, x, add_constant)
            assert np.allclose(fit.params, ref["params"], rtol=1e-10)
            assert np.allclose(fit.bse, ref["bse"], rtol=1e-10)
            assert np.allclose(fit.pvalues, ref["pvalues"], rtol=1e-8, atol=1e-300)
            assert fit.rsquared == ref["rsquared"] or abs(fit.rsquared - ref["rsquared"]) < 1e-12
            assert abs(fit.rsquared_adj - ref["rsquared_adj"]) < 1e-12
            assert fit.df_resid == ref["df_resid"] and fit.param_names == (["const", "x1"] if add_constant else ["x1"])
    summary = reg.RegressionSummary(zip(["S_2", "D_2", "D_1"], fits[:3]))
    text = summary.as_text()
    assert text.count("==") == 6 and "x1" in text and summary.as_frame().shape[1] == 3

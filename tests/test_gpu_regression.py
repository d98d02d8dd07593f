"""GPU parity: per-component regressions vs the statsmodels-OLS oracle."""

import numpy as np
import pytest

from oracle import modwt_oracle as mo
from oracle import ols_oracle
from oracle import pywt_oracle as pw

pytestmark = pytest.mark.gpu


def _check(fit, ref, tol=1e-9):
    assert np.allclose(fit.params, ref["params"], rtol=tol, atol=tol)
    assert np.allclose(fit.bse, ref["bse"], rtol=tol, atol=tol)
    assert np.allclose(fit.tvalues, ref["tvalues"], rtol=1e-7, atol=1e-7)
    assert np.allclose(fit.pvalues, ref["pvalues"], rtol=1e-6, atol=1e-12)
    assert np.isclose(fit.rsquared, ref["rsquared"], atol=tol) and np.isclose(fit.rsquared_adj, ref["rsquared_adj"], atol=tol)
    assert fit.nobs == ref["nobs"] and fit.df_resid == ref["df_resid"]


@pytest.mark.parametrize("add_constant", [True, False])
def test_rowwise_ols_matches_oracle(shim, add_constant):
    from wavelet_transformer_b200.api import regression as reg
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, 333)) + 5.0                      # a large mean stresses the centring
    y = 2.0 - 0.7 * x + rng.standard_normal((9, 333))
    fits = reg.fits_from_stats(shim.rowwise_ols(x, y, add_constant=add_constant, f64=True), add_constant)
    for j, fit in enumerate(fits):
        _check(fit, ols_oracle.ols(y[j], x[j], add_constant))
    one = reg.fits_from_stats(shim.rowwise_ols(x, y[0], add_constant=add_constant, f64=True), add_constant)
    for j, fit in enumerate(one):                                 # one y row against every x row
        _check(fit, ols_oracle.ols(y[0], x[j], add_constant))
    f32 = reg.fits_from_stats(shim.rowwise_ols(x, y, add_constant=add_constant, f64=False), add_constant)
    assert np.allclose(f32[3].params, fits[3].params, rtol=1e-5)
    with pytest.raises(ValueError):
        shim.rowwise_ols(x[:, :2], y[:, :2], add_constant=True)


def test_time_scale_regression_dwt(series):
    """regression.time_scale_regression on the reference's sample series (regression.py:91-126)."""
    from src import regression
    n = min(series["inflation_value"].size, series["expectation_value"].size)
    a, b = series["inflation_value"][:n], series["expectation_value"][:n]
    res = regression.time_scale_regression(a, b, 5, "db4")
    assert list(res) == ["S_5", "D_5", "D_4", "D_3", "D_2", "D_1"]
    ca, cb = pw.wavedec(a, "db4", level=5), pw.wavedec(b, "db4", level=5)
    for j, name in enumerate(res):
        only = lambda cs: pw.waverec([c if i == j else np.zeros_like(c) for i, c in enumerate(cs)], "db4")
        _check(res[name], ols_oracle.ols(only(cb), only(ca)), tol=1e-8)
    text = res.as_text()
    assert "S_5" in text and "R-squared" in text and res.as_frame().shape[1] == 6


def test_time_scale_regression_modwt_and_approximation(series):
    from src import modwt, regression
    n = min(series["inflation_value"].size, series["expectation_value"].size)
    a, b = series["inflation_value"][:n], series["expectation_value"][:n]
    wa, wb = modwt.modwtmra(modwt.modwt(a, "sym4", 4), "sym4"), modwt.modwtmra(modwt.modwt(b, "sym4", 4), "sym4")
    res = modwt.time_scale_regression(wa[::-1], wb[::-1], 4)      # rows S_4, D_4, .., D_1
    ra, rb = mo.modwtmra(mo.modwt(a, "sym4", 4), "sym4")[::-1], mo.modwtmra(mo.modwt(b, "sym4", 4), "sym4")[::-1]
    for j, name in enumerate(res):
        _check(res[name], ols_oracle.ols(rb[j], ra[j]), tol=1e-8)
    smooth = modwt.smooth_signal(modwt.modwt(a, "sym4", 4), "sym4", 4)
    fits = regression.wavelet_approximation(smooth, b, 4)
    ref_smooth = mo.smooth_signal(mo.modwt(a, "sym4", 4), "sym4", 4)
    for c, fit in fits.items():
        _check(fit, ols_oracle.ols(b, ref_smooth[c]["signal"]), tol=1e-8)
    single = regression.simple_regression(a, b)
    _check(single, ols_oracle.ols(b, a))
    # the reference's own form: simple_regression(data: DataFrame, x_var, y_var, add_constant)
    import pandas as pd
    df = pd.DataFrame({"infl": a, "expect": b})
    named = regression.simple_regression(df, "infl", "expect")
    _check(named, ols_oracle.ols(b, a))
    assert named.param_names == ["const", "infl"]
    no_const = regression.simple_regression(df, "infl", "expect", add_constant=False)
    assert no_const.params.shape == (1,) and no_const.param_names == ["infl"]
    assert no_const.params[0] == pytest.approx(float(a @ b / (a @ a)), rel=1e-10)

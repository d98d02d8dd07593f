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

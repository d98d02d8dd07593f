"""float64 NumPy restatement of the regressions the reference runs through statsmodels.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  statsmodels (requirements.txt) is not
installable here; ``sm.OLS(y, X).fit()`` is restated from its documented definitions --
``params = pinv(X) y``, ``bse = sqrt(diag(ssr / df_resid * (X'X)^-1))``, centred total sum of
squares when X holds a constant, uncentred otherwise -- and pinned in tests against
``scipy.stats.linregress`` / ``numpy.linalg.lstsq``.  Call sites: src/regression.py:54-64,
76-81, 118-121; src/modwt.py:218-222.
"""

from __future__ import annotations

import numpy as np
from scipy import stats


def ols(y, x, add_constant=True):
    y = np.asarray(y, dtype=float)
    x = np.asarray(x, dtype=float)
    X = np.column_stack([np.ones_like(x), x]) if add_constant else x[:, None]
    params = np.linalg.pinv(X) @ y
    resid = y - X @ params
    nobs, k = X.shape
    df = nobs - k
    ssr = float(resid @ resid)
    cov = ssr / df * np.linalg.inv(X.T @ X)
    bse = np.sqrt(np.diag(cov))
    tvalues = params / bse
    tss = float(((y - y.mean()) ** 2).sum()) if add_constant else float(y @ y)
    r2 = 1.0 - ssr / tss
    r2_adj = 1.0 - (nobs - (1 if add_constant else 0)) / df * (1.0 - r2)
    return {"params": params, "bse": bse, "tvalues": tvalues, "pvalues": 2 * stats.t.sf(np.abs(tvalues), df),
            "rsquared": r2, "rsquared_adj": r2_adj, "nobs": nobs, "ssr": ssr, "df_resid": df}

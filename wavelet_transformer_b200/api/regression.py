"""Per-component regressions -- mirrors src/regression.py:54-126 and src/modwt.py:197-229.

The reference fits ``sm.OLS(output_j, sm.add_constant(input_j))`` once per component vector
(S_J, D_J, .., D_1) and prints the fits side by side with ``summary_col``.  Each of those is a
one-regressor least-squares problem, i.e. five sums per row: all components of both series are
reconstructed in ONE batched ``waverec`` launch and all regressions run in ONE
``wtb_rowwise_ols`` launch (csrc/regress.cu); standard errors, t statistics and p values are
closed forms evaluated on the host.

statsmodels is not a dependency of this package, so the result objects are small stand-ins that
carry the attributes the reference's tables show (``params, bse, tvalues, pvalues, rsquared,
rsquared_adj, nobs``) and render a ``summary_col``-like text table with ``as_text()``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np
import numpy.typing as npt

from .. import _shim
from .. import pywt_compat as pywt

MOTHER = pywt.Wavelet("db4")


def _t_sf2(t: np.ndarray, df: float) -> np.ndarray:
    """Two-sided p value of a Student t statistic."""
    from scipy import stats  # host-side closed form only
    return 2.0 * stats.t.sf(np.abs(t), df)


@dataclass
class ComponentFit:
    """The part of statsmodels' RegressionResults the reference prints."""

    params: npt.NDArray          # [const, x1] (or [x1] without a constant)
    bse: npt.NDArray
    tvalues: npt.NDArray
    pvalues: npt.NDArray
    rsquared: float
    rsquared_adj: float
    nobs: int
    ssr: float
    df_resid: int
    param_names: List[str] = field(default_factory=list)

    def summary_lines(self) -> List[str]:
        out = []
        for name, b, se, p in zip(self.param_names, self.params, self.bse, self.pvalues):
            stars = "***" if p < 0.01 else "**" if p < 0.05 else "*" if p < 0.1 else ""
            out.append(f"{name:<8}{b:>12.4f}{stars:<3} ({se:.4f})")
        out.append(f"R-squared      {self.rsquared:.4f}   adj. {self.rsquared_adj:.4f}   N {self.nobs}")
        return out


def fits_from_stats(stats: npt.NDArray, add_constant: bool) -> List[ComponentFit]:
    """[rows, 8] output of wtb_rowwise_ols -> one ComponentFit per row."""
    fits = []
    for nobs, icpt, slope, ssr, tss, sxx, mean_x, _ in np.atleast_2d(stats):
        k = 2 if add_constant else 1
        df = int(nobs) - k
        sigma2 = ssr / df
        se_slope = np.sqrt(sigma2 / sxx)
        if add_constant:
            se_icpt = np.sqrt(sigma2 * (1.0 / nobs + mean_x * mean_x / sxx))
            params, bse, names = np.array([icpt, slope]), np.array([se_icpt, se_slope]), ["const", "x1"]
        else:
            params, bse, names = np.array([slope]), np.array([se_slope]), ["x1"]
        with np.errstate(divide="ignore", invalid="ignore"):
            tv = params / bse
            r2 = 1.0 - ssr / tss
        # statsmodels: 1 - (nobs - k_constant) / df_resid * (1 - R^2)
        r2_adj = 1.0 - (nobs - (1 if add_constant else 0)) / df * (1.0 - r2)
        fits.append(ComponentFit(params, bse, tv, _t_sf2(tv, df), float(r2), float(r2_adj), int(nobs),
                                 float(ssr), df, names))
    return fits


class RegressionSummary(dict):
    """``{component name: ComponentFit}`` in the reference's column order, with the text table
    the reference prints through ``summary_col(...).as_text()``."""

    def as_text(self) -> str:
        blocks = []
        for name, fit in self.items():
            blocks.append(f"== {name} ==")
            blocks.extend(fit.summary_lines())
        blocks.append("Standard errors in parentheses.  * p<.1, ** p<.05, *** p<.01")
        return "\n".join(blocks)

    def as_frame(self):
        import pandas as pd
        cols = {}
        for name, fit in self.items():
            col = {}
            for pname, b, se, p in zip(fit.param_names, fit.params, fit.bse, fit.pvalues):
                col[pname], col[f"{pname}_se"], col[f"{pname}_p"] = b, se, p
            col["R-squared"], col["R-squared Adj."], col["N"] = fit.rsquared, fit.rsquared_adj, fit.nobs
            cols[name] = col
        return pd.DataFrame(cols)

    def __str__(self) -> str:
        return self.as_text()


def _component_names(levels: int) -> List[str]:
    return [f"S_{levels}" if j == 0 else f"D_{levels - j + 1}" for j in range(levels + 1)]


def simple_regression(data, x_var, y_var=None, add_constant: bool = True) -> ComponentFit:
    """``sm.OLS(data[y_var], add_constant(data[x_var])).fit()`` (src/regression.py:54-64).

    The reference's form is ``simple_regression(data: DataFrame, x_var: str, y_var: str,
    add_constant=True)``; anything indexable by column name (DataFrame, dict of arrays) works.
    ``simple_regression(x, y)`` with two array-likes is kept as a shorthand.  The result is a
    ``ComponentFit`` stand-in for statsmodels' results object (params, bse, tvalues, pvalues,
    rsquared, rsquared_adj, nobs -- no ``.summary()``)."""
    if isinstance(x_var, str):
        if not isinstance(y_var, str):
            raise TypeError("simple_regression(data, x_var, y_var): y_var must be a column name")
        x, y, names = data[x_var], data[y_var], ["const", x_var]
    else:
        if isinstance(y_var, bool):          # simple_regression(x, y, False), the array shorthand
            add_constant, y_var = y_var, None
        if y_var is not None:
            raise TypeError("simple_regression(x, y[, add_constant]) takes two arrays, or (data, x_var, y_var)")
        x, y, names = data, x_var, None
    fit = fits_from_stats(_shim.rowwise_ols(np.asarray(x, dtype=float), np.asarray(y, dtype=float),
                                            add_constant=add_constant, f64=True), add_constant)[0]
    if names:
        fit.param_names = names if add_constant else names[1:]
    return fit


def _rowwise(xs: Sequence[npt.NDArray], ys: Sequence[npt.NDArray], add_constant: bool) -> List[ComponentFit]:
    """Regress ys[j] on xs[j]; rows of one length go to the device together."""
    fits: List[ComponentFit] = [None] * len(xs)
    by_len: Dict[int, List[int]] = {}
    for j, (a, b) in enumerate(zip(xs, ys)):
        if len(a) != len(b):
            raise ValueError(f"component {j}: {len(a)} input vs {len(b)} output samples")
        by_len.setdefault(len(a), []).append(j)
    for idx in by_len.values():
        stats = _shim.rowwise_ols(np.stack([np.asarray(xs[j], dtype=float) for j in idx]),
                                  np.stack([np.asarray(ys[j], dtype=float) for j in idx]),
                                  add_constant=add_constant, f64=True)
        for j, fit in zip(idx, fits_from_stats(stats, add_constant)):
            fits[j] = fit
    return fits


def time_scale_regression_components(input_coeffs, output_coeffs, levels: int,
                                     add_constant: bool = True) -> RegressionSummary:
    """Regress output on input for each component vector S_J, D_J, .., D_1 given as rows
    (src/modwt.py:197-229: ``input_coeffs[j]`` against ``output_coeffs[j]``, j = 0 is S_J)."""
    fits = _rowwise([input_coeffs[j] for j in range(levels + 1)],
                    [output_coeffs[j] for j in range(levels + 1)], add_constant)
    return RegressionSummary(zip(_component_names(levels), fits))


def component_signals(coeffs: list, wavelet) -> npt.NDArray:
    """``reconstruct_signal_component(coeffs, wavelet, j)`` for every j in one batched launch:
    row j is the reconstruction from block j alone (src/dwt.py:110-120)."""
    w = wavelet if hasattr(wavelet, "rec_lo") else pywt.Wavelet(wavelet)
    parts = [np.asarray(c, dtype=float).ravel() for c in coeffs]
    lens = np.array([p.size for p in parts], dtype=np.int32)
    edges = np.concatenate([[0], np.cumsum(lens)])
    packed = np.zeros((len(parts), int(edges[-1])))
    for j, part in enumerate(parts):
        packed[j, edges[j]:edges[j + 1]] = part
    if len(parts) == 1:
        return packed
    return np.asarray(_shim.waverec(packed, lens, w.rec_lo, w.rec_hi, f64=True), dtype=float)


def time_scale_regression(input_data: npt.NDArray, output_data: npt.NDArray, levels: int,
                          mother_wavelet: str, add_constant: bool = True) -> RegressionSummary:
    """DWT both series, rebuild every component vector on its own and regress output on input
    component by component (src/regression.py:91-126)."""
    wavelet = pywt.Wavelet(mother_wavelet) if isinstance(mother_wavelet, str) else mother_wavelet
    comp_in = component_signals(pywt.wavedec(input_data, wavelet, level=levels), wavelet)
    comp_out = component_signals(pywt.wavedec(output_data, wavelet, level=levels), wavelet)
    stats = _shim.rowwise_ols(comp_in, comp_out, add_constant=add_constant, f64=True)
    return RegressionSummary(zip(_component_names(levels), fits_from_stats(stats, add_constant)))


def wavelet_approximation(smooth_t_dict: Dict[int, Dict[str, npt.NDArray]], original_y: npt.NDArray,
                          levels: int, add_constant: bool = True, verbose: bool = False) -> Dict[int, ComponentFit]:
    """Regress the original series on each smoothed input (src/regression.py:67-88): the
    single y row is broadcast against the ``levels`` smooth rows in one launch."""
    crystals = list(range(1, levels + 1))
    xs = np.stack([np.asarray(smooth_t_dict[c]["signal"], dtype=float) for c in crystals])
    stats = _shim.rowwise_ols(xs, np.asarray(original_y, dtype=float), add_constant=add_constant, f64=True)
    out = dict(zip(crystals, fits_from_stats(stats, add_constant)))
    if verbose:
        for c, fit in out.items():
            print(f"\n-----Smoothed model, Removing D_{list(range(1, c + 1))}-----\n")
            print("\n".join(fit.summary_lines()))
    return out

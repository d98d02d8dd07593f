"""Pre-/post-processing helpers with the reference's names and semantics
(src/utils/wavelet_helpers.py:13-78).  O(n) host arithmetic on either side of
the GPU transforms."""

from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger(__name__)


def align_series(t_values, series_vlaues):
    """Drop leading samples so the series is as long as ``t_values``
    (wavelet_helpers.py:13-19; argument spelling kept)."""
    extra = abs(len(series_vlaues) - len(t_values))
    if extra:
        logger.warning("Trimming series signal")
        return series_vlaues[extra:]
    return series_vlaues


def standardize_series(series, detrend: bool = True, standardize: bool = True, remove_mean: bool = False):
    """Detrend (degree-1 fit) or de-mean, then divide by the RAW series' standard
    deviation (wavelet_helpers.py:22-57)."""
    series = np.asarray(series)
    if detrend and remove_mean:
        raise ValueError("Only standardize by either removing secular trend or mean, not both.")
    raw_std, raw_mean = series.std(), series.mean()
    out = series
    if detrend:
        x = np.arange(0, series.size)
        out = series - np.polyval(np.polyfit(x, series, 1), x)
    if remove_mean:
        out = out - raw_mean
    if standardize:
        out = out / raw_std
    return out


def normalize_xwt_results(signal_size, xwt_coeffs, coi, coi_min, freqs, signif):
    """period, |W12|^2, power/signif ratio and the clipped log2 COI polygon
    (wavelet_helpers.py:60-78)."""
    period = 1 / freqs
    power = np.abs(xwt_coeffs) ** 2
    sig95 = power / (np.ones([1, signal_size]) * signif[:, None])
    tail = np.log2(period[-1:])
    coi_plot = np.concatenate([np.log2(coi), [1e-9], tail, tail, [1e-9]]).clip(min=coi_min)
    return period, power, sig95, coi_plot

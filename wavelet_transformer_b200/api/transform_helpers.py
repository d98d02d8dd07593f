"""Dict-of-transforms helpers -- mirrors src/utils/transform_helpers.py:21-140.

Same names, arguments and return shapes as the reference.  The ``create_*_dict`` builders are
host glue (DataFrame columns -> dataclasses).  The ``create_*_results_dict`` runners are where
the reference loops ``pywt.wavedec`` / ``run_cwt`` / ``run_xwt`` over measures one at a time;
here measures of equal shape go to the device as ONE batch (one launch per group) and are
split back into the reference's per-measure result objects.
"""

from __future__ import annotations

from collections import defaultdict
from typing import Dict, Hashable, List, Tuple

import numpy as np

from .. import _shim
from .. import pycwt_compat as wavelet
from .. import pywt_compat as pywt
from . import cwt, dwt, wavelet_helpers, xwt
from .cwt import DataForCWT, ResultsFromCWT
from .dwt import DataForDWT, ResultsFromDWT
from .xwt import DataForXWT, ResultsFromXWT

DATE = "date"   # constants/ids.py:38


def create_dwt_dict(data_for_dwt, measures_list: List[str], **kwargs) -> Dict[str, DataForDWT]:
    """One ``DataForDWT`` per column at the maximum useful level (transform_helpers.py:21-46)."""
    mother_wavelet = kwargs.get("mother_wavelet", dwt.MOTHER)
    out = {}
    for measure in measures_list:
        y_values = data_for_dwt[measure].to_numpy()
        max_level = pywt.dwt_max_level(len(y_values), mother_wavelet.dec_len)
        out[measure] = DataForDWT(y_values=y_values, mother_wavelet=mother_wavelet, levels=max_level)
    return out


def create_cwt_dict(data_for_cwt, measures_list: List[str], **kwargs) -> Dict[str, DataForCWT]:
    """One ``DataForCWT`` per column: rows where the column is NaN dropped, series detrended
    and divided by its standard deviation (transform_helpers.py:49-63)."""
    out = {}
    for measure in measures_list:
        keep = data_for_cwt[measure].notna()
        t_values = data_for_cwt[keep][DATE].to_numpy()
        y_values = wavelet_helpers.standardize_series(data_for_cwt[keep][measure].to_numpy())
        out[measure] = DataForCWT(t_values=t_values, y_values=y_values, **kwargs)
    return out


def create_xwt_dict(data_for_xwt, xwt_list: List[Tuple[str, str]], **kwargs) -> Dict[Tuple[str, str], DataForXWT]:
    """One ``DataForXWT`` per pair of columns (rows with any NaN dropped), with the reference's
    XWT constants (transform_helpers.py:66-86, constants/results_configs.py:48-58)."""
    out = {}
    for comparison in xwt_list:
        complete = data_for_xwt.dropna()
        y1 = wavelet_helpers.standardize_series(complete[comparison[0]].to_numpy(), **kwargs)
        y2 = wavelet_helpers.standardize_series(complete[comparison[1]].to_numpy(), **kwargs)
        out[comparison] = DataForXWT(y1_values=y1, y2_values=y2, mother_wavelet=xwt.MOTHER_DICT[xwt.MOTHER],
                                     delta_t=xwt.DT, delta_j=xwt.DJ, initial_scale=xwt.S0, levels=xwt.LEVELS)
    return out


def _groups(keys: List[Hashable], shape_of) -> Dict[Hashable, List[Hashable]]:
    """Keys grouped by transform shape, first-seen order kept inside each group."""
    groups = defaultdict(list)
    for k in keys:
        groups[shape_of(k)].append(k)
    return groups


def _wavedec_batched(dwt_data_dict, measures_list, level_of) -> Dict[str, list]:
    """``pywt.wavedec`` of every measure; equal (length, filter, level) share one launch."""
    def shape_of(m):
        d = dwt_data_dict[m]
        w = pywt._w(d.mother_wavelet)
        return (len(d.y_values), tuple(w.dec_lo), tuple(w.dec_hi), level_of(d))

    coeffs = {}
    for (n, lo, hi, level), members in _groups(list(dict.fromkeys(measures_list)), shape_of).items():
        if level is None:
            level = pywt.dwt_max_level(n, len(lo))
        if level < 0:
            raise ValueError(f"Level value of {level} is too low . Minimum level is 0.")
        batch = np.stack([np.asarray(dwt_data_dict[m].y_values, dtype=float) for m in members])
        packed, lens = _shim.wavedec(batch, lo, hi, int(level), f64=True)
        packed = np.asarray(packed, dtype=float).reshape(len(members), -1)
        edges = np.concatenate([[0], np.cumsum(lens)])
        for row, m in enumerate(members):
            coeffs[m] = [np.array(packed[row, edges[i]:edges[i + 1]]) for i in range(len(lens))]
    return coeffs


def create_dwt_results_dict(dwt_data_dict: Dict[str, DataForDWT], measures_list: List[str],
                            **kwargs) -> Dict[str, ResultsFromDWT]:
    """Coefficients only (transform_helpers.py:89-104): ``ResultsFromDWT(wavedec(...), levels)``."""
    coeffs = _wavedec_batched(dwt_data_dict, measures_list, lambda d: d.levels)
    return {m: ResultsFromDWT(coeffs[m], dwt_data_dict[m].levels) for m in measures_list}


def create_dwt_regression_dict(dwt_data_dict: Dict[str, DataForDWT], measures_list: List[str],
                               **kwargs) -> Dict[str, ResultsFromDWT]:
    """``run_dwt`` of every measure (transform_helpers.py:107-114): like the plain results, but a
    ``levels=None`` entry reports the maximum useful level (src/dwt.py:93-100)."""
    coeffs = _wavedec_batched(dwt_data_dict, measures_list, lambda d: d.levels)
    out = {}
    for m in measures_list:
        d = dwt_data_dict[m]
        levels = d.levels
        if levels is None:
            levels = pywt.dwt_max_level(data_len=len(d.y_values), filter_len=d.mother_wavelet.dec_len)
        out[m] = ResultsFromDWT(coeffs[m], levels)
    return out


def create_cwt_results_dict(cwt_data_dict: Dict[str, DataForCWT], measures_list: List[str],
                            **kwargs) -> Dict[str, ResultsFromCWT]:
    """``run_cwt`` of every measure (transform_helpers.py:117-124).  Morlet series of equal length
    share one fused CWT+power launch; other mothers go through ``run_cwt`` one by one."""
    out = {}

    def shape_of(m):
        d = cwt_data_dict[m]
        mother = wavelet._as_mother(d.mother_wavelet)
        return (len(d.y_values), mother.f0) if isinstance(mother, wavelet.Morlet) else ("single", m)

    for shape, members in _groups(list(dict.fromkeys(measures_list)), shape_of).items():
        if shape[0] == "single":
            out[members[0]] = cwt.run_cwt(cwt_data_dict[members[0]], **kwargs)
        else:
            out.update(zip(members, cwt.run_cwt_batch([cwt_data_dict[m] for m in members], **kwargs)))
    return {m: out[m] for m in measures_list}


def create_xwt_results_dict(xwt_data_dict: Dict[Tuple[str, str], DataForXWT], xwt_list: List[Tuple[str, str]],
                            **kwargs) -> Dict[Tuple[str, str], ResultsFromXWT]:
    """``run_xwt`` of every comparison (transform_helpers.py:127-140); comparisons of one shape
    share the two launches of ``run_xwt_batch``."""
    def shape_of(c):
        d = xwt_data_dict[c]
        mother = wavelet._as_mother(d.mother_wavelet)
        if not isinstance(mother, wavelet.Morlet) or len(d.y1_values) != len(d.y2_values):
            return ("single", c)
        return (len(d.y1_values), d.delta_t, d.delta_j, d.initial_scale, mother.f0)

    out = {}
    for shape, members in _groups(list(dict.fromkeys(xwt_list)), shape_of).items():
        if shape[0] == "single" or len(members) == 1:
            for c in members:
                out[c] = xwt.run_xwt(xwt_data_dict[c], **kwargs)
        else:
            out.update(zip(members, xwt.run_xwt_batch([xwt_data_dict[c] for c in members], **kwargs)))
    return {c: out[c] for c in xwt_list}

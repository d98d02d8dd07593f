"""MODWT entry points -- mirror src/modwt.py:56-194, 232-251.

The reference carries its own circular a-trous arithmetic on
``scipy.ndimage.convolve1d(mode="wrap")``; here the analysis / synthesis
pyramids and the MRA correlations run as CUDA kernels, and only the (tiny)
equivalent-filter construction for the MRA stays on the host.
"""

from __future__ import annotations

import numpy as np
import numpy.typing as npt

from .. import _shim
from .. import pywt_compat as pywt

MOTHER = pywt.Wavelet("db4")


def _bank(filters):
    w = filters if hasattr(filters, "dec_lo") else pywt.Wavelet(filters)
    return np.asarray(w.dec_lo, dtype=float), np.asarray(w.dec_hi, dtype=float)


def upArrow_op(li, j):
    """Insert 2^(j-1)-1 zeros between taps (modwt.py:56-63); j == 0 -> [1]."""
    if j == 0:
        return [1]
    step = 2 ** (j - 1)
    out = np.zeros(step * (len(li) - 1) + 1)
    out[::step] = li
    return out


def period_list(li, N):
    """Periodise a filter to length N (modwt.py:66-78): zero-pad to the next
    multiple of N -- a whole extra N when already a multiple -- then fold."""
    f = np.concatenate([np.asarray(li, dtype=float), np.zeros(N - len(li) % N)])
    return f if f.size < 2 * N else f.reshape(-1, N).sum(axis=0)


def _correlate_rows(rows, filters_periodised):
    """out[r, t] = sum_m f[r, m] rows[r, (t + m) mod N] on the device (wtb_modwtmra); the kernel
    wants at least two rows, so a single row is sent twice."""
    rows = np.atleast_2d(np.asarray(rows, dtype=float))
    filt = np.atleast_2d(np.asarray(filters_periodised, dtype=float))
    if rows.shape[0] == 1:
        return np.asarray(_shim.modwtmra(np.vstack([rows, rows]), np.vstack([filt, filt])), dtype=float)[:1]
    return np.asarray(_shim.modwtmra(rows, filt), dtype=float)


def _dilated_periodised(taps, j, N, sign):
    """f[(sign * 2^(j-1) * l) mod N] += taps[l]: the level-j a-trous filter folded onto N samples."""
    f = np.zeros(N)
    np.add.at(f, (sign * 2 ** (j - 1) * np.arange(len(taps))) % N, np.asarray(taps, dtype=float))
    return f


def circular_convolve_mra(h_j_o, w_j):
    """One MRA row: sum_l h_j_o[l] w_j[(t + l) mod N] (modwt.py:81-83)."""
    w_j = np.asarray(w_j, dtype=float)
    f = np.zeros(w_j.size)
    f[: len(h_j_o)] = h_j_o
    return _correlate_rows(w_j, f)[0]


def circular_convolve_d(h_t, v_j_1, j):
    """Level-j analysis step w_j[t] = sum_l h_t[l] v_{j-1}[(t - 2^(j-1) l) mod N] (modwt.py:86-102)."""
    v = np.asarray(v_j_1, dtype=float)
    return _correlate_rows(v, _dilated_periodised(h_t, j, v.size, -1))[0]


def circular_convolve_s(h_t, g_t, w_j, v_j, j):
    """Level-j synthesis step v_{j-1}[t] = sum_l h_t[l] w_j[(t + 2^(j-1) l) mod N]
    + g_t[l] v_j[(t + 2^(j-1) l) mod N] (modwt.py:105-123)."""
    w = np.asarray(w_j, dtype=float)
    both = _correlate_rows(np.vstack([w, np.asarray(v_j, dtype=float)]),
                           np.vstack([_dilated_periodised(h_t, j, w.size, 1),
                                      _dilated_periodised(g_t, j, w.size, 1)]))
    return both[0] + both[1]


def modwt(x, filters, level):
    """Rows w_1..w_J, v_J of the maximal-overlap DWT (modwt.py:126-144)."""
    g, h = _bank(filters)
    return np.asarray(_shim.modwt(np.asarray(x, dtype=float), g, h, int(level)), dtype=float)


def imodwt(w, filters):
    """Inverse MODWT (modwt.py:147-160)."""
    g, h = _bank(filters)
    return np.asarray(_shim.imodwt(np.asarray(w, dtype=float), g, h), dtype=float)


def mra_filters(filters, level, N):
    """Periodised equivalent filters [h_1 .. h_J, g_J] (modwt.py:172-193)."""
    g, h = _bank(filters)
    bank = []
    g_part = np.array([1.0])
    for j in range(level):
        g_part = np.convolve(g_part, upArrow_op(g, j))
        h_j = np.convolve(g_part, upArrow_op(h, j + 1)) / 2 ** ((j + 1) / 2.0)
        if j == 0:
            h_j = h / np.sqrt(2)
        bank.append(period_list(h_j, N))
    g_j = np.convolve(g_part, upArrow_op(g, level)) / 2 ** (level / 2.0)
    bank.append(period_list(g_j, N))
    return np.vstack(bank)


def modwtmra(w, filters):
    """Multiresolution analysis: details D_1..D_J and smooth S_J (modwt.py:163-194)."""
    w = np.asarray(w, dtype=float)
    g, h = _bank(filters)
    # synthesis cascade on the device: same rows as the equivalent-filter correlation
    return np.asarray(_shim.modwtmra_taps(w, g, h), dtype=float)


def smooth_signal(modwt_coeffs: npt.NDArray, mother_wavelet: str, levels: int):
    """For l = levels..1 zero detail rows 0..l-1 and invert (modwt.py:232-251)."""
    out = {}
    stack = []
    for l in range(levels, 0, -1):
        kept = np.array(modwt_coeffs, dtype=float, copy=True)
        kept[:l] = 0.0
        out[l] = {"coeffs": kept}
        stack.append(kept)
    g, h = _bank(mother_wavelet)
    signals = _shim.imodwt(np.stack(stack), g, h)  # one batched launch for all levels
    for i, l in enumerate(range(levels, 0, -1)):
        out[l]["signal"] = np.asarray(signals[i], dtype=float)
    return out


def time_scale_regression(input_coeffs, output_coeffs, levels: int, add_constant: bool = True):
    """Regress output on input for each component vector S_J, D_J, .., D_1 (modwt.py:197-229);
    all ``levels + 1`` regressions run in one device launch."""
    from .regression import time_scale_regression_components
    return time_scale_regression_components(input_coeffs, output_coeffs, levels, add_constant)

"""float64 NumPy restatement of the PyWavelets 1.9.0 calls the reference makes.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: PyWavelets
(requirements.txt:39, uv.lock:941-947) is an un-vendored native dependency that
cannot be installed here.  What is restated is the published algorithm of its C
core (``convolution.c: downsampling_convolution`` in MODE_SYMMETRIC and
``upsampling_convolution_valid_sf``; ``_multilevel.py: wavedec/waverec``;
``_dwt.py: dwt_max_level``), following SURVEY.md Appendix B.  Call sites:

* ``pywt.Wavelet(name)``   <- src/dwt.py:28, src/modwt.py:30,132,150,166
* ``pywt.dwt_max_level``   <- src/dwt.py:95, src/utils/transform_helpers.py:41
* ``pywt.wavedec``         <- src/dwt.py:104, transform_helpers.py:96, regression.py:103-104
* ``pywt.waverec``         <- src/dwt.py:71,120
"""

from __future__ import annotations

import math

import numpy as np

# Scaling (low-pass decomposition) filters exactly as tabulated by PyWavelets.
_DEC_LO = {
    "haar": [0.7071067811865476, 0.7071067811865476],
    "db2": [-0.12940952255126037, 0.2241438680420134,
            0.8365163037378079, 0.48296291314453416],
    "db3": [0.035226291882100656, -0.08544127388224149, -0.13501102001039084,
            0.4598775021193313, 0.8068915093133388, 0.3326705529509569],
    "db5": [0.003335725285001549, -0.012580751999015526, -0.006241490213011705,
            0.07757149384006515, -0.03224486958502952, -0.24229488706619015,
            0.13842814590110342, 0.7243085284385744, 0.6038292697974729, 0.160102397974125],
    "db4": [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965,
            -0.18703481171888114, -0.02798376941698385, 0.6308807679295904,
            0.7148465705525415, 0.23037781330885523],
    "sym4": [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545,
             0.8037387518059161, 0.29785779560527736, -0.09921954357684722,
             -0.012603967262037833, 0.0322231006040427],
}
_DEC_LO["db1"] = _DEC_LO["haar"]
_DEC_LO["la8"] = _DEC_LO["sym4"]  # Percival & Walden "LA(8)" == PyWavelets sym4


class Wavelet:
    """pywt.Wavelet for orthogonal families: the four filter banks + dec_len."""

    def __init__(self, name: str):
        key = name.lower()
        if key not in _DEC_LO:
            raise ValueError(f"Unknown wavelet name '{name}'")
        lo = np.asarray(_DEC_LO[key], dtype=float)
        L = lo.size
        self.name = name
        self.dec_lo = lo.tolist()
        # quadrature mirror: dec_hi[k] = (-1)^(k+1) dec_lo[L-1-k]
        self.dec_hi = [(-1.0) ** (k + 1) * lo[L - 1 - k] for k in range(L)]
        self.rec_lo = self.dec_lo[::-1]
        self.rec_hi = self.dec_hi[::-1]
        self.dec_len = self.rec_len = L


def _as_wavelet(w):
    return w if hasattr(w, "dec_lo") else Wavelet(w)


def dwt_max_level(data_len: int, filter_len: int) -> int:
    """floor(log2(data_len / (filter_len - 1))), 0 when the signal is too short."""
    if filter_len < 2 or data_len < filter_len - 1:
        return 0
    return max(int(math.floor(math.log2(data_len / (filter_len - 1.0)))), 0)


def dwt_coeff_len(n: int, L: int) -> int:
    """Output length of one symmetric-mode analysis step."""
    return (n + L - 1) // 2


def _sym_ext(x, pad):
    """Half-sample symmetric extension by ``pad`` on both sides (repeats when
    pad exceeds the signal length, as PyWavelets' MODE_SYMMETRIC does)."""
    n = x.size
    idx = np.arange(-pad, n + pad)
    period = 2 * n
    m = np.mod(idx, period)
    m = np.where(m >= n, period - 1 - m, m)
    return x[m]


def dwt(x, wavelet):
    """One analysis step, mode='symmetric': cA[i] = sum_j lo[j]*xe[2i+1-j]."""
    w = _as_wavelet(wavelet)
    x = np.asarray(x, dtype=float)
    L = w.dec_len
    n = x.size
    xe = _sym_ext(x, L - 1)           # xe[p] == x_ext[p - (L-1)]
    nout = dwt_coeff_len(n, L)
    lo = np.asarray(w.dec_lo)
    hi = np.asarray(w.dec_hi)
    cA = np.empty(nout)
    cD = np.empty(nout)
    for i in range(nout):
        seg = xe[2 * i + 1 + (L - 1) - np.arange(L)]
        cA[i] = lo.dot(seg)
        cD[i] = hi.dot(seg)
    return cA, cD


def idwt(cA, cD, wavelet):
    """One synthesis step: the 'valid' part (length 2*len-L+2) of the
    zero-stuffed convolution with rec_lo / rec_hi."""
    w = _as_wavelet(wavelet)
    cA = np.asarray(cA, dtype=float)
    cD = np.asarray(cD, dtype=float)
    L = w.dec_len
    m = cA.size
    up_a = np.zeros(2 * m)
    up_d = np.zeros(2 * m)
    up_a[::2] = cA
    up_d[::2] = cD
    full = np.convolve(up_a, w.rec_lo) + np.convolve(up_d, w.rec_hi)
    nout = 2 * m - L + 2
    return full[L - 2:L - 2 + nout]


def wavedec(data, wavelet, mode="symmetric", level=None):
    """[cA_L, cD_L, ..., cD_1]."""
    if mode != "symmetric":
        raise NotImplementedError("oracle covers mode='symmetric' (the pywt default)")
    w = _as_wavelet(wavelet)
    a = np.asarray(data, dtype=float)
    if level is None:
        level = dwt_max_level(a.size, w.dec_len)
    out = []
    for _ in range(level):
        a, d = dwt(a, w)
        out.append(d)
    out.append(a)
    return out[::-1]


def waverec(coeffs, wavelet, mode="symmetric"):
    w = _as_wavelet(wavelet)
    a = np.asarray(coeffs[0], dtype=float)
    for d in coeffs[1:]:
        d = np.asarray(d, dtype=float)
        if a.size == d.size + 1:
            a = a[:-1]
        a = idwt(a, d, w)
    return a

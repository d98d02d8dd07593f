"""PyWavelets-shaped facade over the B200 engine.

The reference imports ``pywt`` for filter taps and the decimated DWT
(src/dwt.py:14,28,95,104,71,120; src/modwt.py:14,30,132;
src/utils/transform_helpers.py:4,41,96; src/regression.py:10,100-104;
constants/results_configs.py:3,28).  ``from wavelet_transformer_b200 import
pywt_compat as pywt`` keeps those call sites unchanged; analysis and synthesis
run in libwavelet_sm100a.so (mode='symmetric', the PyWavelets default).
"""

from __future__ import annotations

import numpy as np

from . import _shim

__all__ = ["Wavelet", "wavedec", "waverec", "dwt", "idwt", "dwt_max_level", "wavelist"]

# decomposition low-pass taps, as tabulated by PyWavelets
_SCALING = {
    "haar": (0.7071067811865476, 0.7071067811865476),
    "db2": (-0.12940952255126037, 0.2241438680420134, 0.8365163037378079, 0.48296291314453416),
    "db3": (0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313,
            0.8068915093133388, 0.3326705529509569),
    "db5": (0.003335725285001549, -0.012580751999015526, -0.006241490213011705, 0.07757149384006515,
            -0.03224486958502952, -0.24229488706619015, 0.13842814590110342, 0.7243085284385744,
            0.6038292697974729, 0.160102397974125),
    "db4": (-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
            -0.02798376941698385, 0.6308807679295904, 0.7148465705525415, 0.23037781330885523),
    "sym4": (-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161,
             0.29785779560527736, -0.09921954357684722, -0.012603967262037833, 0.0322231006040427),
}
_ALIASES = {"db1": "haar", "la8": "sym4", "sym2": "db2", "sym3": "db3"}


def wavelist():
    return sorted(list(_SCALING) + list(_ALIASES))


class Wavelet:
    """Orthogonal wavelet filter bank: dec_lo/dec_hi/rec_lo/rec_hi, dec_len."""

    def __init__(self, name: str):
        key = name.lower()
        key = _ALIASES.get(key, key)
        if key not in _SCALING:
            raise ValueError(f"Unknown wavelet name '{name}', check wavelist() for the list of available builtin wavelets.")
        lo = list(_SCALING[key])
        n = len(lo)
        self.name = name
        self.dec_lo = lo
        self.dec_hi = [(-1.0) ** (k + 1) * lo[n - 1 - k] for k in range(n)]
        self.rec_lo = lo[::-1]
        self.rec_hi = self.dec_hi[::-1]
        self.dec_len = self.rec_len = n
        self.orthogonal = True

    @property
    def filter_bank(self):
        return self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi

    def __repr__(self):
        return f"Wavelet({self.name!r})"


def _w(wavelet) -> Wavelet:
    return wavelet if hasattr(wavelet, "dec_lo") else Wavelet(wavelet)


def _mode(mode):
    if mode not in ("symmetric", "sym"):
        raise NotImplementedError("only mode='symmetric' (the PyWavelets default) is implemented")


def dwt_max_level(data_len, filter_len):
    if hasattr(filter_len, "dec_len"):
        filter_len = filter_len.dec_len
    return _shim.dwt_max_level(int(data_len), int(filter_len))


def wavedec(data, wavelet, mode="symmetric", level=None, axis=-1):
    """Multilevel decomposition -> ``[cA_n, cD_n, ..., cD_1]`` (float64 arrays)."""
    _mode(mode)
    w = _w(wavelet)
    x = np.asarray(data, dtype=float)
    if x.ndim != 1:
        raise NotImplementedError("wavedec facade takes 1-D data (use _shim.wavedec for batches)")
    if level is None:
        level = dwt_max_level(x.size, w.dec_len)
    if level < 0:
        raise ValueError(f"Level value of {level} is too low . Minimum level is 0.")
    packed, lens = _shim.wavedec(x, w.dec_lo, w.dec_hi, int(level))
    edges = np.concatenate([[0], np.cumsum(lens)])
    return [np.array(packed[edges[i]:edges[i + 1]], dtype=float) for i in range(len(lens))]


def waverec(coeffs, wavelet, mode="symmetric", axis=-1):
    """Multilevel reconstruction from ``[cA_n, cD_n, ..., cD_1]``."""
    _mode(mode)
    w = _w(wavelet)
    if len(coeffs) < 1:
        raise ValueError("Coefficient list too short (minimum 1 arrays required).")
    parts = [np.asarray(c, dtype=float).ravel() for c in coeffs]
    if len(parts) == 1:
        return parts[0].copy()
    lens = np.array([p.size for p in parts], dtype=np.int32)
    out = _shim.waverec(np.concatenate(parts), lens, w.rec_lo, w.rec_hi)
    return np.asarray(out, dtype=float)


def dwt(data, wavelet, mode="symmetric"):
    cA, cD = wavedec(data, wavelet, mode, level=1)
    return cA, cD


def idwt(cA, cD, wavelet, mode="symmetric"):
    return waverec([cA, cD], wavelet, mode)

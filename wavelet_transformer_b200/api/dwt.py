"""Discrete wavelet transform entry point -- mirrors src/dwt.py:27-120."""

from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Dict, Type

import numpy as np
import numpy.typing as npt

from .. import _shim
from .. import pywt_compat as pywt

logger = logging.getLogger(__name__)

MOTHER = pywt.Wavelet("db4")


@dataclass
class DataForDWT:
    """Inputs of a DWT run (src/dwt.py:31-37)."""

    y_values: npt.NDArray
    mother_wavelet: Type
    levels: int = None


def trim_signal(original_signal, reconstructed):
    """Odd-length inputs reconstruct one sample long: drop the FIRST sample
    (src/dwt.py:76-85)."""
    if len(original_signal) % 2 != 0:
        logger.warning("Trimming signal at beginning")
        return reconstructed[1:]
    return reconstructed


@dataclass
class ResultsFromDWT:
    """Coefficients ``[cA_L, cD_L, ..., cD_1]``, level count and, after
    ``smooth_signal``, the per-level smoothed signals (src/dwt.py:40-73)."""

    coeffs: npt.NDArray
    levels: int
    smoothed_signal_dict: Dict[int, Dict[str, npt.NDArray]] = field(default_factory=dict)

    def smooth_signal(self, y_values: npt.NDArray, mother_wavelet: Type) -> None:
        """For l = levels..1 zero the l finest detail blocks and reconstruct;
        entry l holds the signal with detail levels <= l removed."""
        out = {}
        w = mother_wavelet if hasattr(mother_wavelet, "rec_lo") else pywt.Wavelet(mother_wavelet)
        parts = [np.asarray(c, dtype=float).ravel() for c in self.coeffs]
        lens = np.array([p.size for p in parts], dtype=np.int32)
        edges = np.concatenate([[0], np.cumsum(lens)])
        order = list(range(self.levels, 0, -1))
        packed = np.tile(np.concatenate(parts), (len(order), 1))
        for row, l in enumerate(order):                       # zero the l finest detail blocks
            packed[row, edges[len(parts) - l]:] = 0.0
        # every smoothing level in ONE batched reconstruction launch
        recs = (np.asarray(_shim.waverec(packed, lens, w.rec_lo, w.rec_hi, f64=True), dtype=float)
                if len(parts) > 1 else packed)
        for row, l in enumerate(order):
            kept = list(self.coeffs)
            for c in range(1, l + 1):
                kept[-c] = np.zeros_like(kept[-c])
            out[l] = {"coeffs": kept, "signal": trim_signal(y_values, recs[row])}
        self.smoothed_signal_dict = out


def run_dwt(dwt_data: Type[DataForDWT]) -> Type[ResultsFromDWT]:
    """Multilevel decomposition (src/dwt.py:88-107).  ``levels=None`` reports the
    maximum useful level and lets ``wavedec`` pick the same value."""
    if dwt_data.levels is None:
        levels = pywt.dwt_max_level(data_len=len(dwt_data.y_values),
                                    filter_len=dwt_data.mother_wavelet.dec_len)
    else:
        levels = dwt_data.levels
    coeffs = pywt.wavedec(dwt_data.y_values, dwt_data.mother_wavelet, level=dwt_data.levels)
    return ResultsFromDWT(coeffs, levels)


def reconstruct_signal_component(signal_coeffs: list, wavelet, level: int):
    """Reconstruct from block ``level`` alone, all others zeroed (src/dwt.py:110-120)."""
    only = [c if i == level else np.zeros_like(c) for i, c in enumerate(signal_coeffs)]
    return pywt.waverec(only, wavelet)

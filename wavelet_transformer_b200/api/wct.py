"""Wavelet coherence entry point -- mirrors src/wct.py:32-158."""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple, Type

import numpy as np
import numpy.typing as npt

from .. import pycwt_compat as wavelet

DT = 1 / 12
DJ = 1 / 8
S0 = 2 * DT
MOTHER = "morlet"
MOTHER_DICT = {
    "morlet": wavelet.Morlet(6),
    "paul": wavelet.Paul(),
    "DOG": wavelet.DOG(),
    "mexicanhat": wavelet.MexicanHat(),
}
LEVELS = [0.0625, 0.125, 0.25, 0.5, 1, 2, 4, 8, 16]
WCT_LEVELS = [0.0, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 1.0]


@dataclass
class DataForWCT:
    """Inputs of a WCT run (src/wct.py:63-81)."""

    t_values: npt.NDArray = field(init=False)
    y1_values: npt.NDArray
    y2_values: npt.NDArray
    mother_wavelet: Type
    delta_t: float
    delta_j: float
    initial_scale: float
    levels: List[float]
    actual_times: npt.NDArray = None

    def __post_init__(self):
        n = self.y1_values.size
        self.t_values = self.actual_times if self.actual_times is not None else np.linspace(1, n + 1, n)


@dataclass
class ResultsFromWCT:
    """Outputs of a WCT run (src/wct.py:84-93)."""

    coherence: npt.NDArray
    period: npt.NDArray
    significance_levels: npt.NDArray
    coi: npt.NDArray
    phase_diff_u: npt.NDArray
    phase_diff_v: npt.NDArray


def calculate_phase_difference(wct_phase: npt.NDArray) -> Tuple[npt.NDArray, npt.NDArray]:
    """Arrow components, Torrence & Webster (1999) convention: in phase = north,
    y1 leading = east (src/wct.py:143-158)."""
    angle = 0.5 * np.pi - wct_phase
    return np.cos(angle), np.sin(angle)


def run_wct(wavelet_coherence_transform: Type[DataForWCT], calculate_signficance: bool = True,
            significance_level: float = 0.95) -> Type[ResultsFromWCT]:
    """Coherence, periods, coherence/significance ratio, COI and phase arrows
    (src/wct.py:96-140; keyword spelling kept)."""
    d = wavelet_coherence_transform
    coherence, phase, coi, freqs, signif = wavelet.wct(
        d.y1_values, d.y2_values, d.delta_t, dj=d.delta_j, s0=d.initial_scale, J=-1,
        sig=calculate_signficance, significance_level=significance_level, wavelet=d.mother_wavelet,
        normalize=True, cache=True)
    period = 1 / freqs
    ratio = np.abs(coherence) / (np.ones([1, d.y1_values.size]) * signif[:, None])
    u, v = calculate_phase_difference(phase)
    return ResultsFromWCT(coherence, period, ratio, coi, u, v)

"""Cross-wavelet entry point -- mirrors src/xwt.py:25-154."""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple, Type

import numpy as np
import numpy.typing as npt

from .. import pycwt_compat as wavelet
from . import wavelet_helpers

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


@dataclass
class DataForXWT:
    """Inputs of an XWT run (src/xwt.py:54-68)."""

    t_values: npt.NDArray = field(init=False)
    y1_values: npt.NDArray
    y2_values: npt.NDArray
    mother_wavelet: Type
    delta_t: float
    delta_j: float
    initial_scale: float
    levels: List[float]

    def __post_init__(self):
        n = self.y1_values.size
        self.t_values = np.linspace(1, n + 1, n)


@dataclass
class ResultsFromXWT:
    """Outputs of an XWT run (src/xwt.py:71-80)."""

    power: npt.NDArray
    period: npt.NDArray
    significance_levels: npt.NDArray
    coi: npt.NDArray
    phase_diff_u: npt.NDArray
    phase_diff_v: npt.NDArray


def calculate_phase_difference(xwt_phase: npt.NDArray) -> Tuple[npt.NDArray, npt.NDArray]:
    angle = 0.5 * np.pi - xwt_phase
    return np.cos(angle), np.sin(angle)


def run_xwt(cross_wavelet_transform: Type[DataForXWT], normalize: bool = True) -> Type[ResultsFromXWT]:
    """Cross-wavelet power, periods, power/significance ratio, COI polygon and
    phase arrows (src/xwt.py:83-139).  As in the reference the phase comes from a
    second coherence call whose ``delta_j=`` keyword is swallowed, i.e. it runs at
    the default dj=1/12."""
    d = cross_wavelet_transform
    w12, coi, freqs, signif = wavelet.xwt(y1=d.y1_values, y2=d.y2_values, dt=d.delta_t, dj=d.delta_j,
                                          s0=d.initial_scale, wavelet=d.mother_wavelet)
    n = d.y1_values.size
    if normalize:
        period, power, ratio, coi_plot = wavelet_helpers.normalize_xwt_results(
            n, w12, coi, np.log2(d.levels[2]), freqs, signif)
    else:
        period = 1 / freqs
        power = w12
        ratio = power / (np.ones([1, n]) * signif[:, None])
        coi_plot = coi
    _, phase, _, _, _ = wavelet.wct(d.y1_values, d.y2_values, d.delta_t, delta_j=d.delta_j,
                                    s0=d.initial_scale, J=-1, sig=False, wavelet=d.mother_wavelet,
                                    normalize=True, cache=True)
    u, v = calculate_phase_difference(phase)
    return ResultsFromXWT(power, period, ratio, coi_plot, u, v)

"""Continuous wavelet transform entry point -- mirrors src/cwt.py:39-135.

Same constants, dataclasses and ``run_cwt`` signature as the reference; the
transform itself runs on the GPU through ``pycwt_compat`` (pycwt-shaped facade).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Type

import numpy as np
import numpy.typing as npt

from .. import _shim
from .. import pycwt_compat as wavelet
from .wavelet_helpers import standardize_series

UNITS = "%"
NORMALIZE = True
DT = 1 / 12            # years
S0 = 2 * DT            # smallest scale
DJ = 1 / 12            # twelve sub-octaves per octave
J = 7 / DJ             # seven octaves (float, as in the reference)
MOTHER = wavelet.Morlet(f0=6)
LEVELS = [0.0625, 0.125, 0.25, 0.5, 1, 2, 4, 8, 16]


@dataclass
class DataForCWT:
    """Inputs of a CWT run (src/cwt.py:48-71).  ``t_values`` must be datetime64."""

    t_values: npt.NDArray
    y_values: npt.NDArray
    mother_wavelet: Type
    delta_t: float
    delta_j: float
    initial_scale: float
    levels: List[float]
    time_range: npt.NDArray = field(init=False)

    def __post_init__(self):
        first_year = np.min(self.t_values).astype("datetime64[Y]").astype(int) + 1970
        self.time_range = np.arange(1, self.t_values.size + 1) * self.delta_t + first_year


@dataclass
class ResultsFromCWT:
    """Outputs of a CWT run (src/cwt.py:74-81)."""

    power: npt.NDArray
    period: npt.NDArray
    significance_levels: npt.NDArray
    coi: npt.NDArray


def run_cwt(cwt_data: Type[DataForCWT], normalize: bool = True, standardize: bool = False,
            calculate_significance: bool = True, significance_level: float = 0.95,
            **kwargs) -> Type[ResultsFromCWT]:
    """Power spectrum, Fourier periods, power/significance ratio and COI.

    Behaviour kept from src/cwt.py:85-135: the module constants DT/DJ/S0/J (not
    the dataclass fields) drive the transform; ``normalize`` has no effect
    because the ``standardize`` branch decides the input; ``ar1`` always runs on
    the raw series and raises ``Warning`` when it cannot be bounded."""
    y = cwt_data.y_values
    signal = standardize_series(y, **kwargs) if standardize else y
    alpha, _, _ = wavelet.ar1(y)
    mother = wavelet._as_mother(cwt_data.mother_wavelet)
    if isinstance(mother, wavelet.Morlet):
        # |W|^2 straight from the fused kernel: no complex plane over PCIe, no host abs()**2, and
        # none of the side outputs of pycwt.cwt the reference discards (src/cwt.py:109)
        _, scales, freqs, coi = wavelet._resolve_s0_J(np.size(signal), DT, DJ, S0, J, mother)
        power, _ = _shim.cwt_morlet(np.asarray(signal, dtype=float), DT, DJ, S0, int(J), mother.f0)
        power = np.asarray(power, dtype=float)
    else:
        wave, scales, freqs, coi, _, _ = wavelet.cwt(signal, DT, DJ, S0, J, mother)
        power = np.abs(wave) ** 2
    period = 1 / freqs
    ratio = None
    if calculate_significance:
        signif, _ = wavelet.significance(1.0, DT, scales, 0, alpha, significance_level=significance_level,
                                         wavelet=cwt_data.mother_wavelet)
        ratio = power / (np.ones([1, len(cwt_data.t_values)]) * signif[:, None])
    return ResultsFromCWT(power, period, ratio, coi)


def run_cwt_batch(cwt_data_list: List[Type[DataForCWT]], normalize: bool = True, standardize: bool = False,
                  calculate_significance: bool = True, significance_level: float = 0.95,
                  **kwargs) -> List[Type[ResultsFromCWT]]:
    """``run_cwt`` of several equal-length Morlet series as ONE fused CWT+power launch
    (the loop of src/utils/transform_helpers.py:117-124 as a batch).  Per series the result
    equals ``run_cwt``'s; ``ar1`` still raises ``Warning`` for a series it cannot bound."""
    if not cwt_data_list:
        return []
    mother = wavelet._as_mother(cwt_data_list[0].mother_wavelet)
    ys = [np.asarray(d.y_values, dtype=float) for d in cwt_data_list]
    if len({y.size for y in ys}) != 1:
        raise ValueError("run_cwt_batch takes series of equal length")
    alphas = [wavelet.ar1(y)[0] for y in ys]
    signals = np.stack([standardize_series(y, **kwargs) if standardize else y for y in ys])
    n0 = signals.shape[1]
    _, scales, freqs, coi = wavelet._resolve_s0_J(n0, DT, DJ, S0, J, mother)
    power, _ = _shim.cwt_morlet(signals, DT, DJ, S0, int(J), mother.f0, f64=True)
    power = np.asarray(power, dtype=float).reshape(len(ys), scales.size, n0)
    period = 1 / freqs
    out = []
    for b, (d, alpha) in enumerate(zip(cwt_data_list, alphas)):
        ratio = None
        if calculate_significance:
            signif, _ = wavelet.significance(1.0, DT, scales, 0, alpha, significance_level=significance_level,
                                             wavelet=mother)
            ratio = power[b] / (np.ones([1, len(d.t_values)]) * signif[:, None])
        out.append(ResultsFromCWT(power[b], period.copy(), ratio, coi.copy()))
    return out

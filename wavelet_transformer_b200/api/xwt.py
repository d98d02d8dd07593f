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


def run_xwt_batch(items: List[Type[DataForXWT]], normalize: bool = True) -> List[Type[ResultsFromXWT]]:
    """``run_xwt`` of several comparisons that share one shape (series length, dt, dj, s0, Morlet
    f0) as two batched launches -- the cross spectra, then the phase at the default dj = 1/12 the
    reference's second call ends up with -- instead of two per comparison
    (src/utils/transform_helpers.py:127-140 loops ``run_xwt``).  Entry by entry the result equals
    ``run_xwt``'s; ``ar1`` still raises ``Warning`` for a series it cannot bound."""
    if not items:
        return []
    from .. import _shim
    d0 = items[0]
    mother = wavelet._as_morlet(d0.mother_wavelet)
    n = d0.y1_values.size
    for d in items:
        same = (d.y1_values.size == n and d.y2_values.size == n and d.delta_t == d0.delta_t
                and d.delta_j == d0.delta_j and d.initial_scale == d0.initial_scale
                and wavelet._as_morlet(d.mother_wavelet).f0 == mother.f0)
        if not same:
            raise ValueError("run_xwt_batch takes comparisons of one shape")
    dt, dj, s0 = d0.delta_t, d0.delta_j, d0.initial_scale
    y1 = np.stack([np.asarray(d.y1_values, dtype=float) for d in items])
    y2 = np.stack([np.asarray(d.y2_values, dtype=float) for d in items])
    a = (y1 - y1.mean(axis=1, keepdims=True)) / y1.std(axis=1, keepdims=True)     # pycwt.xwt(normalize=True)
    b = (y2 - y2.mean(axis=1, keepdims=True)) / y2.std(axis=1, keepdims=True)
    Jr, _, freqs, coi = wavelet._resolve_s0_J(n, dt, dj, s0, -1, mother)
    _, _, w12 = _shim.xwt_wct(a, b, dt, dj, s0, Jr, mother.f0, want_wct=False, want_phase=False, want_w12=True)
    w12 = np.asarray(w12, dtype=np.complex128).reshape(len(items), Jr + 1, n)
    # the phase call: pycwt.wct(..., delta_j=dj) swallows the keyword, so dj = 1/12 and J follows from it
    dj_phase = 1 / 12
    J_phase = int(np.round(np.log2(n * dt / s0) / dj_phase))
    _, phase, _ = _shim.xwt_wct(a, b, dt, dj_phase, s0, J_phase, mother.f0, want_wct=False, want_phase=True)
    phase = np.asarray(phase, dtype=float).reshape(len(items), J_phase + 1, n)
    dof = mother.dofmin
    out = []
    for i, d in enumerate(items):
        a1, a2 = wavelet.ar1(y1[i])[0], wavelet.ar1(y2[i])[0]
        pk = (wavelet.ar1_spectrum(freqs * dt, a1) * wavelet.ar1_spectrum(freqs * dt, a2)) ** 0.5
        signif = pk * wavelet._chi2_ppf(0.95, dof) / dof       # std1 = std2 = 1 after normalisation
        if normalize:
            period, power, ratio, coi_plot = wavelet_helpers.normalize_xwt_results(
                n, w12[i], coi, np.log2(d.levels[2]), freqs, signif)
        else:
            period, power, coi_plot = 1 / freqs, w12[i], coi
            ratio = power / (np.ones([1, n]) * signif[:, None])
        u, v = calculate_phase_difference(phase[i])
        out.append(ResultsFromXWT(power, period, ratio, coi_plot, u, v))
    return out

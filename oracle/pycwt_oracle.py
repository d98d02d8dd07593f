"""float64 NumPy restatement of the pycwt 0.4.0b0 routines the reference calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: pycwt is a
third-party dependency of the reference (requirements.txt:33, uv.lock:874-886)
that is neither vendored under /root/reference nor installable offline; the
algorithm below is the published one (Torrence & Compo 1998; Grinsted et al.
2004) in the exact form pycwt 0.4.0b0 implements it, following SURVEY.md
Appendix A.  Call sites it serves in the reference:

* ``pycwt.ar1``            <- src/cwt.py:106
* ``pycwt.cwt``            <- src/cwt.py:110
* ``pycwt.significance``   <- src/cwt.py:123
* ``pycwt.wct``            <- src/wct.py:106, src/xwt.py:122
* ``pycwt.xwt``            <- src/xwt.py:93
* ``pycwt.wct_significance``, ``rednoise``, ``Morlet.smooth`` <- inside ``wct``

FFT convention: the pip/uv install of the reference has no ``mkl_fft`` so pycwt
falls back to ``scipy.fftpack`` and zero-pads every transform to the next power
of two (``fft_kwargs``).  ``pad_pow2=False`` gives the conda/mkl behaviour.
"""

from __future__ import annotations

import math

import numpy as np
from scipy import fft as _fft
from scipy.signal import convolve2d, lfilter

F0_DEFAULT = 6.0
NBINS = 1000


def _pow2(n: int) -> int:
    """pycwt helpers.fft_kwargs: ``int(2 ** ceil(log2(n)))``."""
    return int(2 ** math.ceil(math.log2(n)))


class Morlet:
    """pycwt.mothers.Morlet.  ``psi_ft`` has no Heaviside step."""

    name = "morlet"

    def __init__(self, f0: float = F0_DEFAULT):
        self.f0 = f0
        self.dofmin = 2
        if f0 == 6:
            self.cdelta, self.gamma, self.deltaj0 = 0.776, 2.32, 0.60
        else:
            self.cdelta = self.gamma = self.deltaj0 = -1

    def psi_ft(self, f):
        return np.pi ** -0.25 * np.exp(-0.5 * (f - self.f0) ** 2)

    def flambda(self) -> float:
        return 4 * np.pi / (self.f0 + np.sqrt(2 + self.f0 ** 2))

    def coi(self) -> float:
        return 1.0 / np.sqrt(2)

    def smooth(self, W, dt, dj, scales, pad_pow2: bool = True):
        """Time (Gaussian, Fourier domain) then scale (boxcar) smoothing, A.3."""
        m, n = W.shape
        npad = _pow2(n) if pad_pow2 else n
        k = 2 * np.pi * _fft.fftfreq(npad)
        gauss = np.exp(-0.5 * (scales / dt)[:, None] ** 2 * k ** 2)
        T = _fft.ifft(gauss * _fft.fft(W, n=npad, axis=1), axis=1)[:, :n]
        if np.isreal(W).all():
            T = T.real
        win = rect(int(np.round(self.deltaj0 / dj * 2)), normalize=True)
        return convolve2d(T, win[:, None], "same")


class Paul:
    """pycwt.mothers.Paul [RECALLED]: psi_ft(f) = 2^m / sqrt(m (2m-1)!) f^m exp(-f) H(f)
    (Torrence & Compo 1998, table 1); no ``smooth``."""

    name = "paul"

    def __init__(self, m: int = 4):
        self.m, self.dofmin = m, 2
        self.cdelta, self.gamma, self.deltaj0 = (1.132, 1.17, 1.50) if m == 4 else (-1, -1, -1)

    def psi_ft(self, f):
        pos = np.where(f > 0, f, 0.0)
        return 2 ** self.m / np.sqrt(self.m * math.factorial(2 * self.m - 1)) * pos ** self.m * np.exp(-pos) * (f > 0)

    def psi0(self) -> float:
        """psi_0(0) = 2^m i^m m! / sqrt(pi (2m)!), real for even m."""
        return float(np.real(2 ** self.m * 1j ** self.m * math.factorial(self.m)
                             / np.sqrt(np.pi * math.factorial(2 * self.m))))

    def flambda(self) -> float:
        return 4 * np.pi / (2 * self.m + 1)

    def coi(self) -> float:
        return np.sqrt(2)  # e-folding time of T&C table 1; pycwt's own value unverified


class DOG:
    """pycwt.mothers.DOG [RECALLED]: psi_ft(f) = -i^m / sqrt(Gamma(m+1/2)) f^m exp(-f^2/2)."""

    name = "dog"

    def __init__(self, m: int = 2):
        self.m, self.dofmin = m, 1
        self.cdelta, self.gamma, self.deltaj0 = {2: (3.541, 1.43, 1.40), 6: (1.966, 1.37, 0.97)}.get(m, (-1, -1, -1))

    def psi_ft(self, f):
        return -(1j ** self.m) / np.sqrt(math.gamma(self.m + 0.5)) * f ** self.m * np.exp(-0.5 * f ** 2)

    def psi0(self) -> float:
        from numpy.polynomial.hermite_e import hermeval
        return float((-1) ** (self.m + 1) * hermeval(0.0, [0] * self.m + [1]) / np.sqrt(math.gamma(self.m + 0.5)))

    def flambda(self) -> float:
        return 2 * np.pi / np.sqrt(self.m + 0.5)

    def coi(self) -> float:
        return 1.0 / np.sqrt(2)


def icwt(W, sj, dt, dj, wavelet):
    """pycwt.icwt [RECALLED]: Torrence & Compo (1998) eq. 11."""
    psi0 = np.pi ** -0.25 if isinstance(wavelet, Morlet) else wavelet.psi0()
    return dj * np.sqrt(dt) / (wavelet.cdelta * psi0) * (np.real(W) / np.sqrt(sj)[:, None]).sum(axis=0)


def rect(n: int, normalize: bool = False):
    """pycwt.helpers.rect: boxcar with half-weight end points (A.2)."""
    w = np.ones(n)
    w[0] = w[-1] = 0.5
    if normalize:
        w /= w.sum()
    return w


def cwt_axes(n0, dt, dj, s0, J, wavelet: Morlet):
    """Scales, Fourier frequencies and COI exactly as ``pycwt.cwt`` forms them."""
    if s0 == -1:
        s0 = 2 * dt / wavelet.flambda()
    if J == -1:
        J = int(np.round(np.log2(n0 * dt / s0) / dj))
    sj = s0 * 2 ** (np.arange(0, J + 1) * dj)
    freqs = 1 / (wavelet.flambda() * sj)
    coi = n0 / 2 - np.abs(np.arange(0, n0) - (n0 - 1) / 2)
    coi = wavelet.flambda() * wavelet.coi() * dt * coi
    return sj, freqs, coi


def cwt(signal, dt, dj=1 / 12, s0=-1, J=-1, wavelet: Morlet | None = None,
        pad_pow2: bool = True):
    """pycwt.cwt (A.1).  Returns (W[S,n0] c128, sj, freqs, coi, fft, fftfreqs)."""
    wavelet = wavelet or Morlet()
    signal = np.asarray(signal, dtype=float)
    n0 = signal.size
    sj, freqs, coi = cwt_axes(n0, dt, dj, s0, J, wavelet)
    N = _pow2(n0) if pad_pow2 else n0
    signal_ft = _fft.fft(signal, n=N)
    ftfreqs = 2 * np.pi * _fft.fftfreq(N, dt)
    sj_col = sj[:, None]
    psi_ft_bar = (sj_col * ftfreqs[1] * N) ** 0.5 * np.conjugate(
        wavelet.psi_ft(sj_col * ftfreqs))
    W = _fft.ifft(signal_ft * psi_ft_bar, axis=1)
    return (W[:, :n0], sj, freqs, coi, signal_ft[1:N // 2] / N ** 0.5,
            ftfreqs[1:N // 2] / (2 * np.pi))


def ar1(x):
    """pycwt.helpers.ar1 (A.4): unbiased lag-1 autocorrelation (Allen & Smith)."""
    x = np.asarray(x, dtype=float)
    N = x.size
    x = x - x.mean()
    c0 = x.dot(x) / N
    c1 = x[:N - 1].dot(x[1:]) / (N - 1)
    B = -c1 * N - c0 * N ** 2 - 2 * c0 + 2 * c1 - c1 * N ** 2 + c0 * N
    A = c0 * N ** 2
    C = N * (c0 + c1 * N - c1)
    D = B ** 2 - 4 * A * C
    if D > 0:
        g = (-B - D ** 0.5) / (2 * A)
    else:
        raise Warning("Cannot place an upperbound on the unbiased AR(1). "
                      "Series is too short or trend is to large.")
    mu2 = -1 / N + (2 / N ** 2) * ((N - g ** N) / (1 - g)
                                   - g * (1 - g ** (N - 1)) / (1 - g) ** 2)
    c0t = c0 / (1 - mu2)
    a = ((1 - g ** 2) * c0t) ** 0.5
    return g, a, mu2


def ar1_spectrum(freqs, ar1_coef=0.0):
    """pycwt.helpers.ar1_spectrum: normalised AR(1) power spectrum."""
    freqs = np.asarray(freqs)
    return (1 - ar1_coef ** 2) / np.abs(1 - ar1_coef * np.exp(-2 * np.pi * 1j * freqs)) ** 2


def chi2_ppf_dof2(level: float) -> float:
    """``scipy.stats.chi2.ppf(level, 2)``; closed form for two degrees of freedom."""
    return -2.0 * math.log1p(-level)


def significance(signal, dt, scales, sigma_test=0, alpha=None,
                 significance_level=0.95, dof=-1, wavelet: Morlet | None = None):
    """pycwt.significance, sigma_test=0 branch (A.5) -- the only one the
    reference reaches (src/cwt.py:123-131)."""
    wavelet = wavelet or Morlet()
    try:
        n0 = len(signal)
    except TypeError:
        n0 = 1
    variance = signal if n0 == 1 else np.asarray(signal).std() ** 2
    if alpha is None:
        alpha, _, _ = ar1(signal)
    period = np.asarray(scales) * wavelet.flambda()
    freq = dt / period
    fft_theor = variance * (1 - alpha ** 2) / (
        1 + alpha ** 2 - 2 * alpha * np.cos(2 * np.pi * freq))
    if sigma_test != 0:
        raise NotImplementedError("oracle covers sigma_test=0 only")
    dofmin = wavelet.dofmin          # 2 for the complex mothers, 1 for DOG (real)
    if dofmin == 2:
        chisquare = chi2_ppf_dof2(significance_level) / dofmin
    else:
        from scipy.stats import chi2
        chisquare = chi2.ppf(significance_level, dofmin) / dofmin
    signif = fft_theor * chisquare
    return signif, fft_theor


def rednoise(N, g, a=1.0, rng=None, mode: str = "ar1"):
    """pycwt.helpers.rednoise (A.8).

    ``mode='ar1'`` (default): y[t] = g*y[t-1] + eps[t] with tau burn-in samples
    dropped -- what the name, Grinsted et al. 2004 and BASELINE.json describe.
    ``mode='white'``: what pycwt 0.4.0b0 literally computes if its ``lfilter``
    call runs over the length-1 last axis of an (N+tau, 1) column (SURVEY A.8
    low-confidence caveat).
    """
    rng = rng or np.random.default_rng()
    if g == 0:
        return rng.standard_normal(N) * a
    tau = int(np.ceil(-2 / np.log(np.abs(g))))
    eps = rng.standard_normal(N + tau) * a
    if mode == "white":
        return eps[tau:]
    return lfilter([1, 0], [1, -g], eps)[tau:]


def rednoise_from_eps(eps, g, tau):
    """AR(1) filter + burn-in drop on caller-supplied innovations (for injected
    surrogate parity)."""
    return lfilter([1, 0], [1, -g], np.asarray(eps, dtype=float))[tau:]


def burn_in(g: float) -> int:
    return 0 if g == 0 else int(np.ceil(-2 / np.log(np.abs(g))))


def _normalise(y):
    return (y - y.mean()) / y.std()


def smoothed_spectra(W1, W2, sj, dt, dj, wavelet, pad_pow2=True):
    """S1, S2, S12 of pycwt.wct / wct_significance (A.6/A.7)."""
    scales = np.ones([1, W1.shape[1]]) * sj[:, None]
    S1 = wavelet.smooth(np.abs(W1) ** 2 / scales, dt, dj, sj, pad_pow2)
    S2 = wavelet.smooth(np.abs(W2) ** 2 / scales, dt, dj, sj, pad_pow2)
    W12 = W1 * W2.conj()
    S12 = wavelet.smooth(W12 / scales, dt, dj, sj, pad_pow2)
    return S1, S2, S12, W12


def wct(y1, y2, dt, dj=1 / 12, s0=-1, J=-1, sig=True, significance_level=0.95,
        wavelet: Morlet | None = None, normalize=True, pad_pow2=True, **kwargs):
    """pycwt.wct (A.6).  Returns (WCT, aWCT, coi, freq, sig)."""
    wavelet = wavelet or Morlet()
    y1 = np.asarray(y1, dtype=float)
    y2 = np.asarray(y2, dtype=float)
    if s0 == -1:
        s0 = 2 * dt / wavelet.flambda()
    if J == -1:
        J = int(np.round(np.log2(y1.size * dt / s0) / dj))
    y1n = _normalise(y1) if normalize else y1
    y2n = _normalise(y2) if normalize else y2
    W1, sj, freq, coi, _, _ = cwt(y1n, dt, dj, s0, J, wavelet, pad_pow2)
    W2, sj, freq, coi, _, _ = cwt(y2n, dt, dj, s0, J, wavelet, pad_pow2)
    S1, S2, S12, W12 = smoothed_spectra(W1, W2, sj, dt, dj, wavelet, pad_pow2)
    WCT = np.abs(S12) ** 2 / (S1 * S2)
    aWCT = np.angle(W12)
    if sig:
        a1 = ar1(y1)[0]
        a2 = ar1(y2)[0]
        sig = wct_significance(a1, a2, dt=dt, dj=dj, s0=s0, J=J,
                               significance_level=significance_level,
                               wavelet=wavelet, **kwargs)
    else:
        sig = np.asarray([0])
    return WCT, aWCT, coi, freq, sig


def xwt(y1, y2, dt, dj=1 / 12, s0=-1, J=-1, significance_level=0.95,
        wavelet: Morlet | None = None, normalize=True, pad_pow2=True):
    """pycwt.xwt (A.9).  Returns (W12, coi, freq, signif)."""
    wavelet = wavelet or Morlet()
    y1 = np.asarray(y1, dtype=float)
    y2 = np.asarray(y2, dtype=float)
    std1, std2 = y1.std(), y2.std()
    y1n = _normalise(y1) if normalize else y1
    y2n = _normalise(y2) if normalize else y2
    W1, sj, freq, coi, _, _ = cwt(y1n, dt, dj, s0, J, wavelet, pad_pow2)
    W2, sj, freq, coi, _, _ = cwt(y2n, dt, dj, s0, J, wavelet, pad_pow2)
    W12 = W1 * W2.conj()
    if normalize:
        std1 = std2 = 1.0
    a1 = ar1(y1)[0]
    a2 = ar1(y2)[0]
    Pk1 = ar1_spectrum(freq * dt, a1)
    Pk2 = ar1_spectrum(freq * dt, a2)
    dof = wavelet.dofmin
    if dof == 2:
        chisquare = chi2_ppf_dof2(significance_level) / dof
    else:
        from scipy.stats import chi2
        chisquare = chi2.ppf(significance_level, dof) / dof
    signif = std1 * std2 * (Pk1 * Pk2) ** 0.5 * chisquare
    return W12, coi, freq, signif


def mc_geometry(dt, dj, s0, J, wavelet: Morlet):
    """Surrogate length, reliable-region mask and maxscale of wct_significance."""
    ms = s0 * (2 ** (J * dj)) / dt
    N = int(np.ceil(ms * 6))
    sj, freq, coi = cwt_axes(N, dt, dj, s0, J, wavelet)
    period = 1.0 / freq
    outsidecoi = period[:, None] <= coi[None, :]
    rows = np.nonzero(outsidecoi.any(axis=1))[0]
    maxscale = int(rows[-1]) if rows.size else 0
    return N, sj, freq, outsidecoi, maxscale


def coherence_histogram(R2, outsidecoi, maxscale, nbins=NBINS, faithful_loop=False):
    """Per-scale histogram of R2 over the reliable region, scales [0, maxscale).

    ``faithful_loop=True`` is pycwt's per-sample Python double loop (what the
    reference actually pays for); the default is the equivalent ``np.bincount``.
    Bin index is ``int(floor(R2*nbins))``; a value that rounds to exactly
    ``nbins`` (R2 == 1.0) would raise IndexError in pycwt and is clamped here.
    """
    S = R2.shape[0]
    wlc = np.zeros((S, nbins), dtype=np.int64)
    for s in range(maxscale):
        vals = R2[s, outsidecoi[s]]
        idx = np.floor(vals * nbins).astype(np.int64)
        idx = np.minimum(idx, nbins - 1)
        if faithful_loop:
            for t in idx:
                wlc[s, int(t)] += 1
        else:
            wlc[s] += np.bincount(idx, minlength=nbins)[:nbins]
    return wlc


def sig_from_histogram(wlc, maxscale, significance_level=0.95, rows_any=None):
    """Percentile step of wct_significance (A.7): cumsum over non-empty bins,
    ``(P-0.5)/P[-1]``, ``np.interp``.  Rows with any reliable point are first
    set to NaN (sic) and rows < maxscale then overwritten."""
    S, nbins = wlc.shape
    sig95 = np.zeros(S)
    if rows_any is not None:
        sig95[rows_any] = np.nan
    R2y = (np.arange(nbins) + 0.5) / nbins
    for s in range(maxscale):
        sel = wlc[s] != 0
        P = wlc[s, sel].astype(float).cumsum()
        P = (P - 0.5) / P[-1]
        sig95[s] = np.interp(significance_level, P, R2y[sel])
    return sig95


def wct_significance(al1, al2, dt, dj, s0, J, significance_level=0.95,
                     wavelet: Morlet | None = None, mc_count=300, rng=None,
                     noise_mode="ar1", surrogates=None, faithful_loop=False,
                     return_hist=False, pad_pow2=True, progress=False, cache=False):
    """pycwt.wct_significance (A.7) without the on-disk cache.

    ``surrogates``: optional array [mc_count, 2, N] of ready-made noise series
    (host-injected parity mode); otherwise AR(1) red noise from ``rng``.
    """
    wavelet = wavelet or Morlet()
    N, sj, freq, outsidecoi, maxscale = mc_geometry(dt, dj, s0, J, wavelet)
    rng = rng or np.random.default_rng()
    wlc = np.zeros((J + 1, NBINS), dtype=np.int64)
    for m in range(mc_count):
        if surrogates is not None:
            n1, n2 = surrogates[m, 0], surrogates[m, 1]
        else:
            n1 = rednoise(N, al1, 1, rng, noise_mode)
            n2 = rednoise(N, al2, 1, rng, noise_mode)
        W1 = cwt(n1, dt, dj, s0, J, wavelet, pad_pow2)[0]
        W2 = cwt(n2, dt, dj, s0, J, wavelet, pad_pow2)[0]
        S1, S2, S12, _ = smoothed_spectra(W1, W2, sj, dt, dj, wavelet, pad_pow2)
        R2 = np.abs(S12) ** 2 / (S1 * S2)
        wlc += coherence_histogram(R2, outsidecoi, maxscale, NBINS, faithful_loop)
    sig95 = sig_from_histogram(wlc, maxscale, significance_level,
                               outsidecoi.any(axis=1))
    return (sig95, wlc) if return_hist else sig95

"""Known-answer and algebraic pins for the pycwt restatement (parity unpinned by
the reference's own tests -- SURVEY.md section 8c -- so the oracle is anchored on
analytic results and on the AR(1) values measured on the reference's sample data)."""

import numpy as np
import pytest
from scipy.signal import convolve2d
from scipy.stats import chi2

from oracle import pycwt_oracle as po

DT = 1 / 12


def test_morlet_constants():
    m = po.Morlet(6)
    assert m.flambda() == pytest.approx(1.03304364775, abs=1e-10)
    assert m.coi() == pytest.approx(2 ** -0.5)
    assert (m.dofmin, m.cdelta, m.gamma, m.deltaj0) == (2, 0.776, 2.32, 0.60)
    # no Heaviside step: negative frequencies contribute pi^-1/4 e^-18
    assert m.psi_ft(0.0) == pytest.approx(np.pi ** -0.25 * np.exp(-18))


def test_cwt_shapes_and_axes_cfg1(series):
    x = series["cpi_value"]
    W, sj, freqs, coi, fft_, fftfreqs = po.cwt(x, DT, 1 / 12, 2 * DT, 7 / (1 / 12))
    assert W.shape == (85, 1346) and sj.size == 85 and coi.size == 1346
    assert fft_.size == fftfreqs.size == 2048 // 2 - 1
    assert sj[0] == 2 * DT and sj[-1] == pytest.approx(2 * DT * 2 ** 7)
    assert coi[0] == pytest.approx(coi[-1]) and coi.argmax() in (672, 673)


def test_cwt_sinusoid_peaks_at_fourier_period():
    t = np.arange(2048) * DT
    for P in (0.5, 2.0, 8.0):
        W, sj, freqs, *_ = po.cwt(np.cos(2 * np.pi * t / P), DT, 1 / 12, 2 * DT, -1)
        peak = 1 / freqs[(np.abs(W) ** 2)[:, 1024].argmax()]
        assert abs(np.log2(peak / P)) <= 1 / 12


def test_cwt_white_noise_mean_power_is_variance():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4096)
    W = po.cwt(x, DT, 1 / 4, 2 * DT, 24)[0]
    assert np.mean(np.abs(W[:16]) ** 2) == pytest.approx(1.0, rel=0.1)


def test_ar1_on_sample_data(series):
    assert po.ar1(series["inflation_value"])[0] == pytest.approx(0.9888356230, abs=1e-9)
    assert po.ar1(series["expectation_value"])[0] == pytest.approx(0.9698623532, abs=1e-9)
    d = 100 * np.diff(np.log(series["cpi_value"]))
    assert po.ar1(d)[0] == pytest.approx(0.4705602642, abs=1e-9)
    with pytest.raises(Warning):
        po.ar1(series["cpi_value"])          # raw CPI: no upper bound (src/wavelet_plots.py:684 fallback)
    with pytest.raises(Warning):
        po.ar1(series["pair_inflation"])


def test_ar1_recovers_known_coefficient():
    rng = np.random.default_rng(1)
    y = po.rednoise(20000, 0.7, 1, rng)
    assert po.ar1(y)[0] == pytest.approx(0.7, abs=0.02)
    assert po.burn_in(0.989) == 181 and po.burn_in(0.966) == 58 and po.burn_in(0.7) == 6


def test_significance_closed_form():
    sj = 2 * DT * 2 ** (np.arange(10) / 2)
    signif, theor = po.significance(1.0, DT, sj, 0, 0.5, significance_level=0.95)
    assert po.chi2_ppf_dof2(0.95) == pytest.approx(chi2.ppf(0.95, 2), rel=1e-13)
    assert np.allclose(signif, theor * chi2.ppf(0.95, 2) / 2)
    f = DT / (sj * po.Morlet().flambda())
    assert np.allclose(theor, po.ar1_spectrum(f, 0.5))


def test_rect_and_scale_window_semantics():
    assert np.allclose(po.rect(10, True), np.r_[0.5, np.ones(8), 0.5] / 9)
    # convolve2d 'same' with an even window: rows i-5 .. i+4, zero fill, no renormalisation
    T = np.zeros((30, 1)); T[12, 0] = 1
    out = convolve2d(T, po.rect(10, True)[:, None], "same")[:, 0]
    assert np.nonzero(out)[0].tolist() == list(range(8, 18))
    assert out[8] == pytest.approx(0.5 / 9) and out[17] == pytest.approx(0.5 / 9)


def test_wct_bounds_and_identities(series):
    y1, y2 = series["pair_inflation"], series["pair_expectation"]
    WCT, aWCT, coi, freq, sig = po.wct(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1, sig=False)
    assert WCT.shape == (66, 565) and sig.tolist() == [0]
    assert WCT.min() >= 0 and WCT.max() <= 1 + 1e-12
    assert np.abs(po.wct(y1, y1, DT, dj=1 / 8, s0=2 * DT, J=-1, sig=False)[0] - 1).max() < 1e-12
    W12 = po.xwt(y2, y2, DT, dj=1 / 8, s0=2 * DT, J=-1)[0]
    W = po.cwt((y2 - y2.mean()) / y2.std(), DT, 1 / 8, 2 * DT, -1)[0]
    assert np.allclose(W12, np.abs(W) ** 2)
    # unknown kwargs are swallowed when sig=False (src/xwt.py:126 relies on it)
    po.wct(y1, y2, DT, delta_j=1 / 8, s0=2 * DT, J=-1, sig=False, cache=True)


def test_mc_geometry_cfg3():
    N, sj, freq, outside, maxscale = po.mc_geometry(DT, 1 / 8, 2 * DT, 65, po.Morlet())
    assert N == 3351 and maxscale == 65 and outside.all(axis=1).sum() == 0
    assert outside.mean() == pytest.approx(0.914, abs=1e-3)


def test_wct_significance_small_and_histogram_variants():
    rng = np.random.default_rng(4)
    sig, hist = po.wct_significance(0.8, 0.6, DT, 1 / 4, 2 * DT, 16, mc_count=3, rng=rng, return_hist=True)
    assert np.isfinite(sig[:-1]).all() and np.isnan(sig[-1])
    assert ((sig[:-1] > 0.3) & (sig[:-1] < 1)).all()
    N, _, _, outside, maxscale = po.mc_geometry(DT, 1 / 4, 2 * DT, 16, po.Morlet())
    assert hist.sum() == 3 * outside[:maxscale].sum()
    R2 = np.random.default_rng(0).uniform(0, 1, outside.shape)
    a = po.coherence_histogram(R2, outside, maxscale, faithful_loop=True)
    b = po.coherence_histogram(R2, outside, maxscale, faithful_loop=False)
    assert np.array_equal(a, b)


def test_other_mothers_known_answers():
    """Torrence & Compo (1998) table 1 / table 2 constants for Paul (m=4) and DOG (m=2, 6):
    psi_0(0), Fourier factors, unit energy of psi_ft, and the reconstruction identity
    icwt(cwt(x)) ~ x that defines C_delta."""
    assert po.Paul(4).psi0() == pytest.approx(1.079, abs=5e-4)
    assert po.DOG(2).psi0() == pytest.approx(0.867, abs=5e-4)
    assert po.DOG(6).psi0() == pytest.approx(0.884, abs=5e-4)
    assert po.Paul(4).flambda() == pytest.approx(1.3963, abs=1e-4)
    assert po.DOG(2).flambda() == pytest.approx(3.9738, abs=1e-4)
    assert po.DOG(6).flambda() == pytest.approx(2.4645, abs=1e-4)
    w = np.linspace(-40, 40, 400001)
    for mother in (po.Paul(4), po.DOG(2), po.DOG(6), po.Morlet(6)):
        assert np.trapezoid(np.abs(mother.psi_ft(w)) ** 2, w) == pytest.approx(1.0, rel=1e-6)
    rng = np.random.default_rng(4)
    t = np.arange(1024)
    x = np.sin(2 * np.pi * t / 37.0) + 0.5 * np.sin(2 * np.pi * t / 90.0) + 0.1 * rng.standard_normal(1024)
    x -= x.mean()
    for mother, dj in ((po.Morlet(6), 1 / 16), (po.Paul(4), 1 / 16), (po.DOG(2), 1 / 16), (po.DOG(6), 1 / 16)):
        W, sj, *_ = po.cwt(x, 1.0, dj, -1, -1, mother)
        rec = po.icwt(W, sj, 1.0, dj, mother)
        mid = slice(200, 824)
        assert np.abs(rec[mid] - x[mid]).std() < 0.1 * x.std()


def test_published_chi_square_and_red_noise_values():
    """Torrence & Compo (1998) sec. 4: the 95 % level of chi-square with two degrees of freedom is
    5.99 (99 %: 9.21), and the normalised red-noise spectrum (their eq. 16) of alpha = 0.72 (their
    Nino3 example) is (1 - a^2) / (1 + a^2 - 2 a cos(2 pi k / N))."""
    assert po.chi2_ppf_dof2(0.95) == pytest.approx(5.991, abs=1e-3)
    assert po.chi2_ppf_dof2(0.99) == pytest.approx(9.210, abs=1e-3)
    a = 0.72
    signif, theor = po.significance(1.0, 0.25, np.array([1.0, 4.0]), 0, a, significance_level=0.95)
    period = np.array([1.0, 4.0]) * po.Morlet().flambda()
    want = (1 - a * a) / (1 + a * a - 2 * a * np.cos(2 * np.pi * 0.25 / period))
    assert np.allclose(theor, want, rtol=1e-14) and np.allclose(signif, want * 5.991464547107979 / 2, rtol=1e-12)
    # zero lag-1 autocorrelation: white noise, flat spectrum of the signal's variance
    assert np.allclose(po.significance(2.0, 0.25, np.array([1.0, 4.0]), 0, 0.0)[1], 2.0)


def test_cwt_matches_the_time_domain_definition():
    """Torrence & Compo (1998) eq. 2, evaluated directly: W_n(s) = sum_n' x_n' conj(psi((n' - n) dt / s))
    with psi(eta) = sqrt(dt / s) pi^(-1/4) exp(i w0 eta) exp(-eta^2 / 2).  The Fourier-domain route the
    oracle (and pycwt) takes is the same quantity up to the sampling of psi^ and the periodic wrap, so
    away from the edges and for scales resolved by the grid the two agree to round-off (psi^ is
    effectively band-limited): an anchor for the
    sign conventions, the sqrt(2 pi s / dt) normalisation and the scale / frequency axes that does not
    pass through any FFT."""
    rng = np.random.default_rng(11)
    n, dt = 512, 0.25
    x = rng.standard_normal(n).cumsum()
    x = (x - x.mean()) / x.std()
    W, sj, *_ = po.cwt(x, dt, 1 / 4, 4 * dt, 16)
    t = np.arange(n)
    for j in (2, 6, 10):                                  # s / dt = 5.7, 11.3, 22.6: well resolved, support << n
        s = sj[j]
        for m in (200, 256, 300):                         # interior samples
            eta = (t - m) * dt / s
            psi = np.sqrt(dt / s) * np.pi ** -0.25 * np.exp(1j * 6.0 * eta) * np.exp(-0.5 * eta ** 2)
            direct = np.sum(x * np.conj(psi))
            assert abs(W[j, m] - direct) <= 1e-10 * np.abs(W[j]).max(), (j, m, W[j, m], direct)   # measured 3e-16


def test_time_smoothing_matches_the_gaussian_window_definition():
    """Grinsted et al. (2004) smooth in time with a Gaussian exp(-t^2 / (2 s^2)) normalised to unit
    weight; pycwt applies it as exp(-0.5 (s/dt)^2 w^2) in the Fourier domain.  Evaluated directly as a
    time-domain convolution (sigma = s/dt samples, zero beyond the series), the interior of the
    time-smoothed field must agree to round-off: an FFT-free anchor for Morlet.smooth's first half."""
    rng = np.random.default_rng(3)
    n, dt, dj = 400, 1 / 12, 1 / 4
    field = rng.standard_normal((6, n)) ** 2
    sj = 2 * dt * 2.0 ** (np.arange(6) * 1.0 + 1)          # s/dt = 4 .. 128
    m = po.Morlet()
    # undo the scale boxcar by giving smooth one row at a time with dj large enough for a 1-tap window
    for i, s in enumerate(sj[:4]):
        T = m.smooth(field[i:i + 1], dt, 1.2, sj[i:i + 1])     # rect(round(0.6 / 1.2 * 2)) = rect(1): identity
        sigma = s / dt
        t = np.arange(n)
        for c in (n // 2 - 7, n // 2, n // 2 + 31):
            if c - 8 * sigma < 0 or c + 8 * sigma >= n:
                continue
            g = np.exp(-0.5 * ((t - c) / sigma) ** 2) / (sigma * np.sqrt(2 * np.pi))
            assert abs(T[0, c] - np.sum(field[i] * g)) <= 1e-9 * field[i].max(), (i, c)

"""GPU: degenerate and boundary inputs through every C entry point -- empty batches, the
shortest series each transform accepts, single scales, one-element coefficient blocks,
ragged batches split by length -- checked against the oracle where there is an answer and
for a clean error where there is none."""

import numpy as np
import pytest

from conftest import normwise_close
from oracle import modwt_oracle as mo
from oracle import pycwt_oracle as po
from oracle import pywt_oracle as pw

pytestmark = pytest.mark.gpu

DT = 1 / 12


def test_empty_batches_are_noops(shim):
    lo, hi = np.array(pw.Wavelet("db4").dec_lo), np.array(pw.Wavelet("db4").dec_hi)
    assert shim.cwt_morlet(np.zeros((0, 64)), DT, 1 / 4, 2 * DT, 8, f64=True)[0].shape == (0, 9, 64)
    assert shim.modwt(np.zeros((0, 64)), lo, hi, 3, f64=True).shape == (0, 4, 64)
    assert shim.imodwt(np.zeros((0, 4, 64)), lo, hi, f64=True).shape == (0, 64)
    assert shim.modwtmra_taps(np.zeros((0, 4, 64)), lo, hi, f64=True).shape == (0, 4, 64)
    packed, lens = shim.wavedec(np.zeros((0, 64)), lo, hi, 2, f64=True)
    assert packed.shape == (0, int(lens.sum()))
    assert shim.waverec(packed, lens, lo[::-1].copy(), hi[::-1].copy(), f64=True).shape[0] == 0
    assert shim.series_prep(np.zeros((0, 16)), f64=True)[0].shape == (0, 16)
    assert shim.xwt_wct(np.zeros((0, 64)), np.zeros((0, 64)), DT, 1 / 4, 2 * DT, 8, f64=True)[0].shape == (0, 9, 64)
    h = shim.wct_mc_hist(0.5, 0.5, DT, 1 / 4, 2 * DT, 8, mc_count=0, f64=False)
    assert h.sum() == 0


@pytest.mark.parametrize("f64", [True, False])
def test_cwt_shortest_series_and_single_scale(shim, f64):
    rng = np.random.default_rng(1)
    for n0, J in ((3, 2), (4, 0), (5, 0), (33, 11)):     # n0 = 2 is NaN in pycwt itself (sqrt of a negative frequency)
        x = rng.standard_normal(n0)
        ref = np.abs(po.cwt(x, DT, 1 / 4, 2 * DT, J)[0]) ** 2
        power, _ = shim.cwt_morlet(x, DT, 1 / 4, 2 * DT, J, f64=f64)
        assert power.shape == (J + 1, n0)
        assert np.abs(power - ref).max() <= (1e-10 if f64 else 1e-4) * max(ref.max(), 1e-30)
    y1, y2 = rng.standard_normal(9), rng.standard_normal(9)
    wct, ph, _ = shim.xwt_wct(y1, y2, DT, 1 / 2, 2 * DT, 3, f64=True)
    ref = po.wct(y1, y2, DT, dj=1 / 2, s0=2 * DT, J=3, sig=False, normalize=False)[0]
    assert np.abs(wct - ref).max() <= 1e-9


def test_filterbanks_on_tiny_series(shim):
    rng = np.random.default_rng(2)
    for name in ("haar", "db2", "sym4"):
        w = pw.Wavelet(name)
        lo, hi = np.array(w.dec_lo), np.array(w.dec_hi)
        for n in (1, 2, 3, 7, 9):
            x = rng.standard_normal(n)
            for J in (1, 4):
                ref = mo.modwt(x, name, J)
                got = shim.modwt(x, lo, hi, J, f64=True)
                assert got.shape == (J + 1, n) and np.abs(got - ref).max() <= 1e-12, (name, n, J)
                assert np.abs(shim.imodwt(got, lo, hi, f64=True) - x).max() <= 1e-10
                assert np.abs(shim.modwtmra_taps(got, lo, hi, f64=True) - mo.modwtmra(ref, name)).max() <= 1e-10
        for n in (len(lo), len(lo) + 1, 2 * len(lo) + 1):       # level 1 with a one-or-two sample overhang
            x = rng.standard_normal(n)
            packed, lens = shim.wavedec(x, lo, hi, 1, f64=True)
            ref = pw.wavedec(x, name, level=1)
            assert list(lens) == [c.size for c in ref] and np.abs(packed - np.concatenate(ref)).max() <= 1e-12
            rec = shim.waverec(packed, lens, np.array(w.rec_lo), np.array(w.rec_hi), f64=True)
            assert np.abs(rec - pw.waverec(ref, name)).max() <= 1e-12
        packed, lens = shim.wavedec(rng.standard_normal(5), lo, hi, 0, f64=True)     # level 0: identity
        assert list(lens) == [5]


def test_rowwise_ols_minimum_and_constant_rows(shim):
    x = np.array([[0.0, 1.0, 2.0], [1.0, 1.0, 1.0]])
    y = np.array([[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]])
    st = shim.rowwise_ols(x, y, f64=True)
    assert st[0, 1] == pytest.approx(1.0) and st[0, 2] == pytest.approx(2.0) and st[0, 3] == pytest.approx(0.0, abs=1e-24)
    assert not np.isfinite(st[1, 2])                    # a constant regressor has no slope (statsmodels: nan / inf)
    with pytest.raises(ValueError):
        shim.rowwise_ols(x[:, :2], y[:, :2])
    with pytest.raises(ValueError):
        shim.rowwise_ols(x, y[:, :2])


def test_icwt_single_scale_and_bad_shapes(shim):
    W = (np.arange(6) + 1j).reshape(1, 6)
    out = shim.icwt(W, np.array([4.0]), 2.0, f64=True)
    assert np.allclose(out, 2.0 * np.arange(6) / 2.0)
    with pytest.raises(ValueError):
        shim.icwt(W, np.array([1.0, 2.0]), 1.0)


def test_ragged_batches_by_length(shim):
    """The engine's batches are rectangular; ragged collections go through one call per length
    (api/regression._rowwise does the same).  Results must equal per-series calls."""
    rng = np.random.default_rng(3)
    series = [rng.standard_normal(n) for n in (64, 100, 64, 37, 100, 64)]
    lo, hi = np.array(pw.Wavelet("sym4").dec_lo), np.array(pw.Wavelet("sym4").dec_hi)
    by_len = {}
    for i, s in enumerate(series):
        by_len.setdefault(s.size, []).append(i)
    out = [None] * len(series)
    for n, idx in by_len.items():
        w = shim.modwt(np.stack([series[i] for i in idx]), lo, hi, 2, f64=True)
        for k, i in enumerate(idx):
            out[i] = w[k]
    for s, w in zip(series, out):
        assert np.abs(w - mo.modwt(s, "sym4", 2)).max() <= 1e-12


def test_host_batches_larger_than_one_staging_chunk(shim):
    """Host-buffer calls stream the batch through <= 1 GiB staging chunks (csrc/cwt.cu, wct.cu):
    series on either side of a chunk boundary, and the short last chunk, must come out exactly
    as they do alone."""
    rng = np.random.default_rng(1 << 20)
    dt = 1 / 12
    x = rng.standard_normal((2300, 1024)).astype(np.float32)      # 2169 series fit one chunk at 120 scales
    power, _ = shim.cwt_morlet(x, dt, 1 / 12, 2 * dt, 119, f64=False)
    pick = [0, 2167, 2168, 2169, 2170, 2299]
    alone, _ = shim.cwt_morlet(x[pick], dt, 1 / 12, 2 * dt, 119, f64=False)
    assert np.array_equal(power[pick], alone)
    del power
    y1 = rng.standard_normal((1500, 400))
    y2 = 0.4 * y1 + rng.standard_normal((1500, 400))
    wct, phase, _ = shim.xwt_wct(y1, y2, dt, 1 / 4, 2 * dt, -1, f64=True)
    pick = [0, 700, 1000, 1200, 1499]
    wa, pa, _ = shim.xwt_wct(y1[pick], y2[pick], dt, 1 / 4, 2 * dt, -1, f64=True)
    assert np.array_equal(wct[pick], wa) and np.array_equal(phase[pick], pa)


def test_shutdown_releases_and_next_call_rebuilds(shim):
    """wtb_shutdown frees every per-thread arena, the parameter buffers and the twiddle tables;
    the next call must rebuild them and give the same numbers (a stale device pointer would not)."""
    rng = np.random.default_rng(3)
    dt = 1 / 12
    x2 = rng.standard_normal((130, 1346))                     # 2048 two-pass kernel (parameter buffer)
    x4 = rng.standard_normal((2, 3000))                       # 4096 register rows (radix-16 tables)

    def run():
        a, _ = shim.cwt_morlet(x2, dt, 1 / 12, 2 * dt, 84, f64=False)
        b, _ = shim.cwt_morlet(x4, dt, 1 / 8, 2 * dt, 60, f64=False)
        c = shim.wct_mc_hist(0.9, 0.8, dt, 1 / 8, 2 * dt, 65, mc_count=4, seed=9, f64=False)
        d = shim.modwt(x4, *[np.array(v) for v in (_la8().dec_lo, _la8().dec_hi)], 5, f64=True)
        return a, b, c, d

    first = run()
    shim.shutdown()
    second = run()
    for u, v in zip(first, second):
        assert np.array_equal(u, v)


def _la8():
    from wavelet_transformer_b200 import pywt_compat as pywt
    return pywt.Wavelet("sym4")


# ---- un-padded transforms: pycwt with mkl_fft (the reference's conda install, environment.yml:126)
@pytest.fixture()
def no_padding(shim):
    shim.set_fft_padding("none")
    yield shim
    shim.set_fft_padding("pow2")


@pytest.mark.parametrize("n0", [565, 1346, 97, 1000])
def test_cwt_unpadded_lengths_fp64(no_padding, series, n0):
    """nfft = n0 (any length): Bluestein's chirp-z transform in the generic kernels against the
    oracle's pad_pow2=False path, FP64 at 1e-10 and FP32 at the 1e-4 gate."""
    shim = no_padding
    x = np.random.default_rng(n0).standard_normal(n0).cumsum()
    x = (x - x.mean()) / x.std()
    W_ref = po.cwt(x, DT, 1 / 12, 2 * DT, -1, pad_pow2=False)[0]
    W_pad = po.cwt(x, DT, 1 / 12, 2 * DT, -1, pad_pow2=True)[0]
    _, W = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, -1, f64=True, want_power=False, want_coef=True)
    assert np.abs(W - W_ref).max() <= 1e-10 * np.abs(W_ref).max()
    if n0 & (n0 - 1):
        assert np.abs(W_pad - W_ref).max() > 1e-6 * np.abs(W_ref).max()       # the two conventions do differ
    p32, _ = shim.cwt_morlet(np.stack([x, x[::-1]]), DT, 1 / 12, 2 * DT, -1, f64=False)
    ok, worst = normwise_close(p32[0], np.abs(W_ref) ** 2, 1e-4)
    assert ok, worst
    # explicit nfft overrides the global rule
    _, W2 = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, -1, f64=True, want_power=False, want_coef=True,
                            nfft=shim.next_pow2(n0))
    assert np.abs(W2 - W_pad).max() <= 1e-10 * np.abs(W_pad).max()


def test_wct_xwt_and_significance_unpadded(no_padding, series):
    shim = no_padding
    from wavelet_transformer_b200 import pycwt_compat as wavelet
    y1 = 100 * np.diff(np.log(series["cpi_value"]))[-565:]
    y2 = series["pair_expectation"]
    WCT, aWCT, coi, freq, _ = po.wct(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1, sig=False, pad_pow2=False)
    got, phase, c2, f2, _ = wavelet.wct(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1, sig=False)
    assert np.abs(got - WCT).max() <= 1e-10
    assert np.abs(np.angle(np.exp(1j * (phase - aWCT)))).max() <= 1e-8
    W12, *_ = po.xwt(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1, pad_pow2=False)
    X12, *_ = wavelet.xwt(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1)
    assert np.abs(X12 - W12).max() <= 1e-10 * np.abs(W12).max()
    # Monte Carlo with injected surrogates at their own length (N = 601, not a power of two)
    dj, s0, J = 1 / 4, 2 * DT, 22
    N, maxscale = shim.wct_mc_geometry(DT, dj, s0, J)
    assert N & (N - 1)
    rng = np.random.default_rng(4)
    sur = np.stack([np.stack([po.rednoise(N, 0.8, 1, rng), po.rednoise(N, 0.6, 1, rng)]) for _ in range(4)])
    sig_ref, hist_ref = po.wct_significance(0.8, 0.6, DT, dj, s0, J, mc_count=4, surrogates=sur, return_hist=True,
                                            pad_pow2=False)
    hist = shim.wct_mc_hist(0.8, 0.6, DT, dj, s0, J, mc_count=4, surrogates=sur, f64=True)
    assert hist.sum() == hist_ref.sum() and np.abs(hist.astype(np.int64) - hist_ref).sum() <= 4
    h32 = shim.wct_mc_hist(0.8, 0.6, DT, dj, s0, J, mc_count=4, surrogates=sur, f64=False)
    cdf = lambda h: h.cumsum(axis=1) / np.maximum(h.sum(axis=1, keepdims=True), 1)
    assert h32.sum() == hist_ref.sum() and np.abs(cdf(h32) - cdf(hist_ref)).max() <= 2e-3
    shim.set_fft_padding("pow2")
    padded = shim.wct_mc_hist(0.8, 0.6, DT, dj, s0, J, mc_count=4, surrogates=sur, f64=True)
    assert not np.array_equal(padded, hist)

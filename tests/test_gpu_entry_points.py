"""GPU: the re-authored reference entry points (src/cwt.py, src/wct.py, src/xwt.py,
src/dwt.py, src/modwt.py) end to end, against the same pipeline built from the oracle.
These read like the reference's own tests (tests/test_cwt.py, test_xwt.py, test_dwt.py)
with the network fetches replaced by the sample_data fixtures and values -- not only
shapes -- asserted."""

import numpy as np
import pytest

from oracle import modwt_oracle as mo
from oracle import pycwt_oracle as po
from oracle import pywt_oracle as pw

pytestmark = pytest.mark.gpu

DT = 1 / 12


@pytest.fixture(autouse=True)
def _engine(shim, tmp_path, monkeypatch):
    monkeypatch.setenv("WTB_CACHE_DIR", str(tmp_path))
    shim.set_precision("fp64")
    yield


def _dates(days):
    return days.astype("datetime64[D]")


def test_run_cwt_matches_reference_pipeline(series):
    from src import cwt
    from src.utils.wavelet_helpers import standardize_series
    y = 100 * np.diff(np.log(series["cpi_value"]))          # the app's AR(1)-bounded fallback series
    t = _dates(series["cpi_days"])[1:]
    data = cwt.DataForCWT(t, y, cwt.MOTHER, cwt.DT, cwt.DJ, cwt.S0, cwt.LEVELS)
    res = cwt.run_cwt(data, standardize=True, detrend=True)
    assert len(data.time_range) == len(t)                    # reference tests/test_cwt.py:30
    assert len(res.power) == len(res.period) == 85           # reference tests/test_cwt.py:34
    dat = standardize_series(y, detrend=True)
    alpha = po.ar1(y)[0]
    W, sj, freqs, coi, _, _ = po.cwt(dat, DT, 1 / 12, 2 * DT, 84.0)
    power = np.abs(W) ** 2
    signif, _ = po.significance(1.0, DT, sj, 0, alpha, significance_level=0.95)
    assert np.abs(res.power - power).max() <= 1e-10 * power.max()
    assert np.allclose(res.period, 1 / freqs, rtol=1e-13)
    assert np.allclose(res.coi, coi, rtol=1e-13)
    assert np.abs(res.significance_levels - power / signif[:, None]).max() <= 1e-9 * (power / signif[:, None]).max()
    assert cwt.run_cwt(data, calculate_significance=False).significance_levels is None


def test_run_cwt_raises_warning_for_unbounded_ar1(series):
    from src import cwt
    data = cwt.DataForCWT(_dates(series["cpi_days"]), series["cpi_value"], cwt.MOTHER, cwt.DT, cwt.DJ,
                          cwt.S0, cwt.LEVELS)
    with pytest.raises(Warning):                             # caught by the app at wavelet_plots.py:684
        cwt.run_cwt(data)


def test_run_cwt_fp32_mode(series, shim):
    from src import cwt
    shim.set_precision("fp32")
    y = 100 * np.diff(np.log(series["cpi_value"]))
    data = cwt.DataForCWT(_dates(series["cpi_days"])[1:], y, cwt.MOTHER, cwt.DT, cwt.DJ, cwt.S0, cwt.LEVELS)
    res = cwt.run_cwt(data)
    power = np.abs(po.cwt(y, DT, 1 / 12, 2 * DT, 84.0)[0]) ** 2
    assert res.power.dtype == np.float64
    assert np.abs(res.power - power).max() <= 1e-4 * power.max()


def test_run_wct_without_significance(series):
    from src import wct
    from src.utils.wavelet_helpers import standardize_series
    y1 = standardize_series(series["pair_inflation"], detrend=False, remove_mean=True)
    y2 = standardize_series(series["pair_expectation"], detrend=True, remove_mean=False)
    data = wct.DataForWCT(y1, y2, wct.MOTHER_DICT[wct.MOTHER], wct.DT, wct.DJ, wct.S0, wct.LEVELS)
    res = wct.run_wct(data, calculate_signficance=False)
    WCT, aWCT, coi, freq, sig = po.wct(y1, y2, DT, dj=1 / 8, s0=2 * DT, J=-1, sig=False)
    assert res.coherence.shape == (66, 565)
    assert np.abs(res.coherence - WCT).max() <= 1e-10
    assert np.allclose(res.period, 1 / freq) and np.allclose(res.coi, coi)
    u, v = np.cos(0.5 * np.pi - aWCT), np.sin(0.5 * np.pi - aWCT)
    assert np.abs(res.phase_diff_u - u).max() <= 1e-8 and np.abs(res.phase_diff_v - v).max() <= 1e-8
    with np.errstate(divide="ignore"):
        assert np.isinf(res.significance_levels).all()      # coherence / [0], as in the reference


def test_run_wct_with_monte_carlo_significance(series):
    """cfg3 end to end: 300-realisation AR(1) Monte Carlo on device (FP64 engine)."""
    from src import wct
    y1 = (100 * np.diff(np.log(series["cpi_value"])))[-565:]   # AR(1)-bounded stand-in (the app's fallback)
    y2 = series["expectation_value"]
    data = wct.DataForWCT(y1, y2, wct.MOTHER_DICT[wct.MOTHER], wct.DT, wct.DJ, wct.S0, wct.LEVELS)
    res = wct.run_wct(data, calculate_signficance=True, significance_level=0.95)
    sig_ratio = res.significance_levels
    assert sig_ratio.shape == (66, 565)
    signif = res.coherence[:, 0] / sig_ratio[:, 0]
    assert np.isnan(signif[-1]) and np.isfinite(signif[:-1]).all()
    assert ((signif[:-1] > 0.5) & (signif[:-1] < 1.0)).all()
    ref = po.wct_significance(po.ar1(y1)[0], po.ar1(y2)[0], DT, 1 / 8, 2 * DT, 65, mc_count=40,
                              rng=np.random.default_rng(3))
    assert np.abs(signif[:60] - ref[:60]).max() < 0.08       # Monte Carlo error of the 40-run reference
    # second call is served from the on-disk cache and is identical
    res2 = wct.run_wct(data, calculate_signficance=True, significance_level=0.95)
    assert np.array_equal(res2.significance_levels, res.significance_levels, equal_nan=True)


def test_run_xwt(series):
    from src import xwt
    y1, y2 = series["expectation_value"], (100 * np.diff(np.log(series["cpi_value"])))[-565:]
    data = xwt.DataForXWT(y1, y2, xwt.MOTHER_DICT[xwt.MOTHER], xwt.DT, xwt.DJ, xwt.S0, xwt.LEVELS)
    res = xwt.run_xwt(data)
    assert len(data.t_values) == 565                         # reference tests/test_xwt.py:48
    assert len(res.power) < len(y1)                          # reference tests/test_xwt.py:53
    W12, coi, freq, signif = po.xwt(y1, y2, DT, dj=1 / 8, s0=2 * DT)
    power = np.abs(W12) ** 2
    assert np.abs(res.power - power).max() <= 1e-10 * power.max()
    assert np.abs(res.significance_levels - power / signif[:, None]).max() <= 1e-9 * (power / signif[:, None]).max()
    assert res.coi.size == 565 + 4 and res.coi.min() >= np.log2(xwt.LEVELS[2])
    # the phase plane comes from the inner wct at pycwt's default dj = 1/12 (delta_j= is swallowed)
    a = po.wct(y1, y2, DT, s0=2 * DT, J=-1, sig=False)[1]
    assert res.phase_diff_u.shape == a.shape and a.shape[0] != res.power.shape[0]
    assert np.abs(res.phase_diff_u - np.cos(0.5 * np.pi - a)).max() <= 1e-8


def test_run_dwt_and_smoothing(series):
    from src import dwt
    y = series["expectation_value"]                          # odd length (565), like tests/test_regression.py:75
    data = dwt.DataForDWT(y, dwt.MOTHER)
    res = dwt.run_dwt(data)
    assert data.levels is None and res.levels == 6           # reference tests/test_dwt.py:36-42
    ref = pw.wavedec(y, "db4")
    assert [len(c) for c in res.coeffs] == [15, 15, 24, 41, 76, 146, 286]
    for a, b in zip(res.coeffs, ref):
        assert np.abs(a - b).max() <= 1e-10
    res.smooth_signal(y, dwt.MOTHER)
    assert len(res.smoothed_signal_dict[res.levels]["signal"]) == len(y)    # tests/test_dwt.py:48-50
    for l in (1, 3, 6):
        kept = [c.copy() for c in ref]
        for c in range(1, l + 1):
            kept[-c][:] = 0
        assert np.abs(res.smoothed_signal_dict[l]["signal"] - pw.waverec(kept, "db4")[1:]).max() <= 1e-10
    comp = dwt.reconstruct_signal_component(res.coeffs, dwt.MOTHER, 2)
    only = [c if i == 2 else np.zeros_like(c) for i, c in enumerate(ref)]
    assert np.abs(comp - pw.waverec(only, "db4")).max() <= 1e-10
    full = sum(dwt.reconstruct_signal_component(res.coeffs, dwt.MOTHER, i) for i in range(7))
    assert np.abs(full[:-1] - y).max() <= 1e-9


def test_pywt_facade_functions(series):
    from wavelet_transformer_b200 import pywt_compat as pywt
    y = series["inflation_value"]
    assert pywt.dwt_max_level(len(y), pywt.Wavelet("db4").dec_len) == 7
    c = pywt.wavedec(y, "sym4", level=3)
    r = pw.wavedec(y, "sym4", level=3)
    assert all(np.abs(a - b).max() <= 1e-10 for a, b in zip(c, r))
    cA, cD = pywt.dwt(y, "db2")
    assert np.abs(pywt.idwt(cA, cD, "db2")[: len(y)] - y).max() <= 1e-10
    with pytest.raises(NotImplementedError):
        pywt.wavedec(y, "db4", mode="periodization")


def test_modwt_module(series, modwt_golden):
    from src import modwt
    for name, filt in (("inflation", "sym4"), ("expectation", "db4")):
        x = series[f"{name}_value"]
        w = modwt.modwt(x, filt, 6)
        assert np.abs(w - modwt_golden[f"{name}|{filt}|modwt"]).max() <= 1e-10 * np.abs(x).max()
        assert np.abs(modwt.imodwt(w, filt) - x).max() <= 1e-9
        mra = modwt.modwtmra(w, filt)
        assert np.abs(mra - modwt_golden[f"{name}|{filt}|mra"]).max() <= 1e-10 * np.abs(x).max()
        sm = modwt.smooth_signal(w, filt, 6)
        assert np.abs(sm[6]["signal"] - modwt_golden[f"{name}|{filt}|smooth6"]).max() <= 1e-10 * np.abs(x).max()
        assert np.abs(sm[1]["signal"] - modwt_golden[f"{name}|{filt}|smooth1"]).max() <= 1e-10 * np.abs(x).max()
        assert np.array_equal(sm[3]["coeffs"][:3], np.zeros((3, x.size)))


def test_pycwt_facade_cwt_side_outputs(series):
    from wavelet_transformer_b200 import pycwt_compat as wavelet
    x = series["expectation_value"]
    W, sj, fr, coi, fft_, fftfreqs = wavelet.cwt(x, DT, 1 / 12, 2 * DT, 7 / (1 / 12), wavelet.Morlet(6))
    Wo, sjo, fro, coio, ffto, fqo = po.cwt(x, DT, 1 / 12, 2 * DT, 7 / (1 / 12))
    assert W.dtype == np.complex128 and np.abs(W - Wo).max() <= 1e-10 * np.abs(Wo).max()
    assert np.allclose(sj, sjo) and np.allclose(fr, fro) and np.allclose(coi, coio)
    assert np.allclose(fft_, ffto) and np.allclose(fftfreqs, fqo)
    Wp = wavelet.cwt(x, DT, 1 / 12, 2 * DT, 40, wavelet="paul")[0]      # other mothers: tests/test_gpu_cwt.py
    assert np.abs(Wp - po.cwt(x, DT, 1 / 12, 2 * DT, 40, po.Paul(4))[0]).max() <= 1e-10 * np.abs(Wp).max()
    with pytest.raises(NotImplementedError):                              # pycwt smooths with Morlet only
        wavelet.wct(x, x[::-1].copy(), DT, sig=False, wavelet="paul")


def test_series_prep_matches_reference_helpers(shim, series, helpers_golden):
    """Batched standardize_series + ar1 (wtb_series_prep) against the reference's own
    standardize_series outputs (golden) and the ar1 oracle."""
    g = helpers_golden
    y, a = shim.series_prep(g["y"], detrend=True, f64=True)
    assert np.abs(y - g["std_detrend"]).max() <= 1e-10 * np.abs(g["std_detrend"]).max()
    y, _ = shim.series_prep(g["y"], detrend=False, remove_mean=True, f64=True)
    assert np.abs(y - g["std_mean"]).max() <= 1e-12
    y, _ = shim.series_prep(g["y"], detrend=False, standardize=False, f64=True)
    assert np.array_equal(y, g["std_raw"])
    assert np.isnan(a)                                        # pair_inflation: pycwt.ar1 raises
    batch = np.stack([series["expectation_value"], (100 * np.diff(np.log(series["cpi_value"])))[-565:],
                      series["pair_inflation"]])
    _, ar = shim.series_prep(batch, f64=True, want_y=False)
    assert ar[0] == pytest.approx(po.ar1(batch[0])[0], rel=1e-11)
    assert ar[1] == pytest.approx(po.ar1(batch[1])[0], rel=1e-11)
    assert np.isnan(ar[2])
    y32, ar32 = shim.series_prep(batch, f64=False)
    assert y32.dtype == np.float32 and ar32[0] == pytest.approx(po.ar1(batch[0])[0], rel=1e-5)
    with pytest.raises(ValueError):
        shim.series_prep(batch, detrend=True, remove_mean=True)


def test_cwt_batch_resident(shim, series):
    """Device-resident standardize -> ar1 -> CWT -> significance ratio vs the host pipeline."""
    import torch
    from src import cwt
    from wavelet_transformer_b200 import engine
    d = 100 * np.diff(np.log(series["cpi_value"]))
    x = torch.tensor(np.stack([d[:1024], d[-1024:], d[100:1124]]), dtype=torch.float64, device="cuda")
    out = engine.cwt_batch_resident(x, cwt.DT, cwt.DJ, cwt.S0, 84, significance_level=0.95)
    torch.cuda.synchronize()
    for b in range(3):
        xb = x[b].cpu().numpy()
        data = cwt.DataForCWT(_dates(series["cpi_days"])[:1024], xb, cwt.MOTHER, cwt.DT, cwt.DJ, cwt.S0, cwt.LEVELS)
        ref = cwt.run_cwt(data, standardize=True, detrend=True)
        got = out["power"][b].cpu().numpy()
        assert np.abs(got - ref.power).max() <= 1e-9 * ref.power.max()
        ratio = got / out["signif"][b].cpu().numpy()[:, None]
        assert np.abs(ratio - ref.significance_levels).max() <= 1e-8 * ref.significance_levels.max()
        assert float(out["ar1"][b]) == pytest.approx(po.ar1(xb)[0], rel=1e-10)


def test_transform_helpers_batch_measures(series):
    """create_*_results_dict (transform_helpers.py:89-140): measures of equal shape share one
    launch; every entry must equal the reference's per-measure loop (run_dwt / wavedec / run_cwt /
    run_xwt one at a time), mixed lengths and levels included."""
    import pandas as pd
    from src import cwt, dwt, xwt
    from src.utils import transform_helpers as th
    rng = np.random.default_rng(12)
    infl, expn = series["inflation_value"], series["expectation_value"]
    n = expn.size
    growth = 100 * np.diff(np.log(series["cpi_value"]))      # AR(1)-bounded, unlike the raw inflation series
    cols = {"infl": growth[-n:], "expn": expn, "noise": rng.standard_normal(n).cumsum()}
    frame = pd.DataFrame({"date": np.arange(n).astype("datetime64[M]"), **cols})
    # DWT: three equal-length measures + one shorter one + one with levels=None
    ddict = th.create_dwt_dict(frame, list(cols))
    ddict["short"] = dwt.DataForDWT(infl[:301], dwt.MOTHER, 4)
    ddict["auto"] = dwt.DataForDWT(expn, dwt.MOTHER, None)
    names = list(ddict)
    plain = th.create_dwt_results_dict(ddict, names)
    regr = th.create_dwt_regression_dict(ddict, names)
    for m in names:
        ref = pw.wavedec(ddict[m].y_values, "db4", level=ddict[m].levels)
        one = dwt.run_dwt(ddict[m])
        assert plain[m].levels == ddict[m].levels and regr[m].levels == one.levels
        for got in (plain[m].coeffs, regr[m].coeffs):
            assert [c.size for c in got] == [c.size for c in ref]
            assert max(np.abs(g - r).max() for g, r in zip(got, ref)) <= 1e-12
            assert all(np.array_equal(g, o) for g, o in zip(got, one.coeffs))
    # CWT: batched launch vs run_cwt one by one (the bounded-AR(1) series of the app)
    kw = dict(mother_wavelet=cwt.MOTHER, delta_t=cwt.DT, delta_j=cwt.DJ, initial_scale=cwt.S0, levels=cwt.LEVELS)
    cdict = th.create_cwt_dict(frame, ["infl", "expn"], **kw)
    cdict["short"] = cwt.DataForCWT(frame["date"].to_numpy()[:300], cols["expn"][:300], **kw)
    cres = th.create_cwt_results_dict(cdict, ["short", "infl", "expn"], standardize=True, detrend=True)
    assert list(cres) == ["short", "infl", "expn"]
    for m, got in cres.items():
        one = cwt.run_cwt(cdict[m], standardize=True, detrend=True)
        assert got.power.shape == one.power.shape == (85, cdict[m].y_values.size)
        assert np.abs(got.power - one.power).max() <= 1e-10 * one.power.max()
        assert np.abs(got.significance_levels - one.significance_levels).max() <= 1e-9 * one.significance_levels.max()
        assert np.array_equal(got.period, one.period) and np.array_equal(got.coi, one.coi)
    assert th.create_cwt_results_dict(cdict, ["infl"], calculate_significance=False)["infl"].significance_levels is None
    # XWT: comparisons of one shape share two launches (cross spectra, phase); a lone one loops
    frame["expn_lag"] = np.roll(cols["expn"], 7)
    pairs = [("infl", "expn"), ("expn", "infl"), ("infl", "expn_lag")]
    xdict = th.create_xwt_dict(frame, pairs)
    xres = th.create_xwt_results_dict(xdict, pairs)
    assert list(xres) == pairs
    for c in pairs:
        one = xwt.run_xwt(xdict[c])
        got = xres[c]
        assert got.power.shape == one.power.shape and got.phase_diff_u.shape == one.phase_diff_u.shape
        assert np.abs(got.power - one.power).max() <= 1e-12 * np.abs(one.power).max()
        assert np.allclose(got.significance_levels, one.significance_levels, rtol=1e-10, atol=0)
        assert np.array_equal(got.period, one.period) and np.allclose(got.coi, one.coi, rtol=1e-13)
        assert np.abs(got.phase_diff_u - one.phase_diff_u).max() <= 1e-9
        assert np.abs(got.phase_diff_v - one.phase_diff_v).max() <= 1e-9
    lone = th.create_xwt_results_dict(xdict, pairs[:1])
    assert np.array_equal(lone[pairs[0]].power, xwt.run_xwt(xdict[pairs[0]]).power)


def test_entry_points_are_reentrant_across_host_threads(shim, series):
    """SURVEY 8b threading: several Streamlit sessions are several host threads in one process.
    ctypes releases the GIL, so the calls below really overlap; every thread has its own scratch
    arena and error slot, and each result must equal the one computed alone."""
    import threading
    from wavelet_transformer_b200 import pywt_compat as pywt
    rng = np.random.default_rng(77)
    la8 = pywt.Wavelet("sym4")
    x32 = rng.standard_normal((64, 1024))
    x64 = rng.standard_normal((3, 700))
    y1, y2 = series["pair_inflation"], series["pair_expectation"]
    n1, n2 = (y1 - y1.mean()) / y1.std(), (y2 - y2.mean()) / y2.std()
    jobs = {
        "cwt32": lambda: shim.cwt_morlet(x32, DT, 1 / 12, 2 * DT, 119, f64=False)[0],
        "cwt64": lambda: shim.cwt_morlet(x64, DT, 1 / 12, 2 * DT, 60, f64=True)[0],
        "wct": lambda: shim.xwt_wct(n1, n2, DT, 1 / 8, 2 * DT, -1, f64=True)[0],
        "modwt": lambda: shim.modwtmra_taps(shim.modwt(x64, la8.dec_lo, la8.dec_hi, 6, f64=True),
                                            la8.dec_lo, la8.dec_hi, f64=True),
        "dwt": lambda: shim.wavedec(series["inflation_value"], la8.dec_lo, la8.dec_hi, 7, f64=True)[0],
        "mc": lambda: shim.wct_mc_hist(0.8, 0.6, DT, 1 / 4, 2 * DT, 20, mc_count=16, seed=5, f64=False),
    }
    alone = {k: np.array(fn()) for k, fn in jobs.items()}
    failures = []

    def worker(name, fn):
        try:
            for _ in range(6):
                if not np.array_equal(np.array(fn()), alone[name]):
                    failures.append(f"{name}: result differs under concurrency")
                    return
        except Exception as exc:   # noqa: BLE001 - reported below
            failures.append(f"{name}: {exc!r}")

    threads = [threading.Thread(target=worker, args=item) for item in jobs.items()]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a worker thread hung"
    assert not failures, failures
    # an error raised in one thread carries that thread's own message
    with pytest.raises(Exception, match="must be >= n0"):
        shim.cwt_morlet(x64, DT, 1 / 12, 2 * DT, 10, nfft=8, f64=True)

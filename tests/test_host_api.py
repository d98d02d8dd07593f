"""Host-side logic that needs no GPU: C-ABI export list, host-only entry points,
facade closed forms, dataclasses, helpers against reference-generated goldens."""

import re
from pathlib import Path

import numpy as np
import pytest

from oracle import pycwt_oracle as po
from oracle import pywt_oracle as pw

ROOT = Path(__file__).resolve().parents[1]
DT = 1 / 12


@pytest.fixture(scope="module")
def shim_nogpu():
    from wavelet_transformer_b200 import _shim
    _shim.lib()
    return _shim


def test_library_exports_every_declared_symbol(shim_nogpu):
    header = (ROOT / "include" / "wtb.h").read_text()
    declared = set(re.findall(r"\b(wtb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(shim_nogpu.EXPORTED_SYMBOLS)
    handle = shim_nogpu.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.wtb_version() == 100


def test_no_gpu_fails_loudly(shim_nogpu):
    if shim_nogpu.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        shim_nogpu.cwt_morlet(np.zeros(64), DT, 1 / 4, 2 * DT, -1)
    with pytest.raises(RuntimeError):
        shim_nogpu.modwt(np.zeros(64), [0.7, 0.7], [-0.7, 0.7], 2)


def test_product_never_imports_oracle():
    for path in (ROOT / "wavelet_transformer_b200").rglob("*.py"):
        assert "oracle" not in path.read_text(), path
    for path in (ROOT / "src").rglob("*.py"):
        assert "oracle" not in path.read_text(), path


@pytest.mark.parametrize("n0,dj,s0,J", [(1346, 1 / 12, 2 * DT, 84), (565, 1 / 8, 2 * DT, -1), (100, 1 / 4, -1, -1)])
def test_cwt_axes_match_oracle(shim_nogpu, n0, dj, s0, J):
    sj, freqs, coi = po.cwt_axes(n0, DT, dj, s0, J, po.Morlet())
    Jr, scales, fr, c = shim_nogpu.cwt_axes(n0, DT, dj, s0, J)
    assert Jr + 1 == sj.size
    assert np.allclose(scales, sj, rtol=1e-14) and np.allclose(fr, freqs, rtol=1e-14)
    assert np.allclose(c, coi, rtol=1e-14)


def test_mc_geometry_and_percentile_match_oracle(shim_nogpu):
    for dj, J in [(1 / 8, 65), (1 / 12, 60), (1 / 4, 20)]:
        N, _, _, outside, maxscale = po.mc_geometry(DT, dj, 2 * DT, J, po.Morlet())
        assert shim_nogpu.wct_mc_geometry(DT, dj, 2 * DT, J) == (N, maxscale)
        assert np.array_equal(shim_nogpu.row_has_points(DT, dj, 2 * DT, J), outside.any(axis=1))
    rng = np.random.default_rng(0)
    hist = rng.integers(0, 50, (12, 1000)).astype(np.int64)
    hist[:, :300] = 0
    hist[3, 500:510] = 0
    ref = po.sig_from_histogram(hist, 10, 0.95, np.ones(12, bool))
    got = shim_nogpu.wct_sig_from_hist(hist.astype(np.uint64), 10, 0.95, np.ones(12, np.uint8))
    assert np.allclose(got[:10], ref[:10], rtol=1e-13) and np.isnan(got[10:]).all()
    for level in (0.5, 0.9, 0.99, 1e-9, 1 - 1e-12):
        ref = po.sig_from_histogram(hist, 12, level)
        assert np.allclose(shim_nogpu.wct_sig_from_hist(hist.astype(np.uint64), 12, level), ref, rtol=1e-13)


def test_dwt_length_helpers(shim_nogpu):
    assert list(shim_nogpu.dwt_coeff_lens(565, 8, 6)) == [15, 15, 24, 41, 76, 146, 286]
    for n in (6, 7, 8, 100, 565, 1333, 4096):
        for L in (2, 4, 8):
            assert shim_nogpu.dwt_max_level(n, L) == pw.dwt_max_level(n, L)


def test_facade_closed_forms(series):
    from wavelet_transformer_b200 import pycwt_compat as wavelet
    from wavelet_transformer_b200 import pywt_compat as pywt
    for key in ("inflation_value", "expectation_value"):
        assert np.allclose(wavelet.ar1(series[key]), po.ar1(series[key]), rtol=1e-12)
    with pytest.raises(Warning):
        wavelet.ar1(series["cpi_value"])
    sj = 2 * DT * 2 ** (np.arange(20) / 12)
    a = wavelet.significance(1.0, DT, sj, 0, 0.72, significance_level=0.9, wavelet=wavelet.Morlet(6))
    b = po.significance(1.0, DT, sj, 0, 0.72, significance_level=0.9)
    assert np.allclose(a[0], b[0], rtol=1e-13) and np.allclose(a[1], b[1], rtol=1e-13)
    m = wavelet.Morlet(6)
    assert m.flambda() == po.Morlet(6).flambda() and m.deltaj0 == 0.6 and m.name == "morlet"
    for cls in (wavelet.Paul, wavelet.DOG, wavelet.MexicanHat):
        assert cls().flambda() > 0          # constructible at import time, as the reference needs
    for name in ("db4", "sym4", "haar", "db2", "db3", "db5"):
        w, o = pywt.Wavelet(name), pw.Wavelet(name)
        assert w.dec_lo == o.dec_lo and w.dec_hi == o.dec_hi and w.rec_lo == o.rec_lo and w.dec_len == o.dec_len
    with pytest.raises(ValueError):
        pywt.Wavelet("nope")


def test_helpers_match_reference_goldens(helpers_golden):
    from src.utils import wavelet_helpers as wh
    from src import wct
    g = helpers_golden
    assert np.array_equal(wh.standardize_series(g["y"], detrend=True), g["std_detrend"])
    assert np.array_equal(wh.standardize_series(g["y"], detrend=False, remove_mean=True), g["std_mean"])
    assert np.array_equal(wh.standardize_series(g["y"], detrend=False, standardize=False), g["std_raw"])
    with pytest.raises(ValueError):
        wh.standardize_series(g["y"], detrend=True, remove_mean=True)
    period, power, sig95, coi_plot = wh.normalize_xwt_results(
        g["xw"].shape[1], g["xw"], g["coi"], float(g["coi_min"]), g["freqs"], g["signif"])
    assert np.array_equal(period, g["nx_period"]) and np.array_equal(power, g["nx_power"])
    assert np.array_equal(sig95, g["nx_sig95"]) and np.array_equal(coi_plot, g["nx_coi_plot"])
    u, v = wct.calculate_phase_difference(g["phase"])
    assert np.array_equal(u, g["phase_u"]) and np.array_equal(v, g["phase_v"])
    assert np.array_equal(wh.align_series(np.arange(10), np.arange(12.0)), g["align"])


def test_dataclasses_and_small_helpers():
    from src import cwt, dwt, modwt, wct, xwt
    t = np.arange("1978-01", "1980-01", dtype="datetime64[M]").astype("datetime64[D]")
    d = cwt.DataForCWT(t, np.zeros(t.size), cwt.MOTHER, cwt.DT, cwt.DJ, cwt.S0, cwt.LEVELS)
    assert len(d.time_range) == len(t)                       # reference tests/test_cwt.py:30
    assert d.time_range[0] == pytest.approx(1978 + 1 / 12)
    assert (cwt.DT, cwt.DJ, cwt.S0, cwt.J) == (1 / 12, 1 / 12, 2 / 12, 84.0)
    w = wct.DataForWCT(np.zeros(5), np.ones(5), wct.MOTHER_DICT[wct.MOTHER], wct.DT, wct.DJ, wct.S0, wct.LEVELS)
    assert np.array_equal(w.t_values, np.linspace(1, 6, 5))
    w2 = wct.DataForWCT(np.zeros(5), np.ones(5), None, 1, 1, 1, [], actual_times=np.arange(5))
    assert np.array_equal(w2.t_values, np.arange(5))
    x = xwt.DataForXWT(np.zeros(7), np.ones(7), None, xwt.DT, xwt.DJ, xwt.S0, xwt.LEVELS)
    assert x.t_values.size == 7
    # reference tests/test_dwt.py:18-27
    sig = list(range(1000))
    assert len(dwt.trim_signal(sig, sig)) == 1000
    sig = list(range(1001))
    assert len(dwt.trim_signal(sig, sig)) == 1000
    assert dwt.DataForDWT(np.zeros(4), dwt.MOTHER).levels is None and dwt.MOTHER.dec_len == 8
    assert list(modwt.upArrow_op([1, 2, 3], 2)) == [1, 0, 2, 0, 3] and modwt.upArrow_op([1, 2], 0) == [1]
    assert modwt.period_list([1, 2, 3, 4, 5], 3).tolist() == [5, 7, 3]
    assert modwt.period_list([1, 2, 3], 3).tolist() == [1, 2, 3]     # whole extra period, folded away


def test_mra_filters_match_oracle():
    from oracle import modwt_oracle as mo
    from src import modwt
    for filt in ("db4", "sym4", "haar"):
        for N in (37, 565):
            a = modwt.mra_filters(filt, 5, N)
            b = np.vstack(mo.mra_filters(filt, 5, N))
            assert np.array_equal(a, b)


def test_significance_cache_files(tmp_path, monkeypatch):
    """The on-disk significance cache: our exact-key file and, on request,
    the file name pycwt itself would use (np.savetxt text, one float per line, gzip)."""
    from wavelet_transformer_b200 import pycwt_compat as wavelet
    p = wavelet.pycwt_cache_file(0.1, -0.2, 1 / 12, 1 / 8, 1 / 6, 65, "morlet")
    assert p.name == "wct_sig_0.00000_1.50000_0.12500_2.00000_65_Morlet.gz"   # arctanh(.4)->0, arctanh(-.8)->-1
    q = wavelet.pycwt_cache_file(0.989, 0.966, 1 / 12, 1 / 8, 1 / 6, 65)
    assert q.name == "wct_sig_nan_nan_0.12500_2.00000_65_Morlet.gz"           # pycwt's collision for |a| > 0.25
    monkeypatch.setenv("WTB_PYCWT_CACHE_DIR", str(tmp_path / "pycwt"))
    monkeypatch.setenv("WTB_CACHE_DIR", str(tmp_path / "own"))
    monkeypatch.setenv("WTB_PYCWT_CACHE", "read")
    sig = np.linspace(0.5, 0.9, 66)
    target = wavelet.pycwt_cache_file(0.1, 0.2, 1 / 12, 1 / 8, 1 / 6, 65)
    target.parent.mkdir(parents=True)
    np.savetxt(target, sig)
    got = wavelet.wct_significance(0.1, 0.2, dt=1 / 12, dj=1 / 8, s0=1 / 6, J=65)   # served from the file: no GPU call
    assert np.allclose(got, sig)


def _build_c_demo(tmp_path):
    import shutil
    import subprocess
    from wavelet_transformer_b200 import _build
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    lib_dir = _build.LIB.parent
    exe = tmp_path / "c_abi_demo"
    subprocess.run([gcc, "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "c_abi_demo.c"),
                    "-o", str(exe), f"-L{lib_dir}", "-lwavelet_sm100a", f"-Wl,-rpath,{lib_dir}", "-lm"], check=True)
    return exe


def test_c_abi_links_from_plain_c(tmp_path):
    """include/wtb.h is valid C and every entry point the demo uses links from gcc; without a
    GPU the program fails loudly through the status code / wtb_last_error path."""
    import subprocess
    import torch
    exe = _build_c_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run is covered by the gpu-marked test")
    proc = subprocess.run([str(exe)], capture_output=True, text=True)
    assert proc.returncode == 1 and "no CPU fallback" in proc.stderr


@pytest.mark.gpu
def test_c_abi_demo_matches_python(tmp_path):
    import subprocess
    from oracle import modwt_oracle as mo
    from oracle import pycwt_oracle as po
    exe = _build_c_demo(tmp_path)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, check=True)
    fields = dict(kv.split("=") for kv in proc.stdout.split())
    t = np.arange(600)
    x = np.sin(2 * np.pi * t / 37.0) + 0.5 * np.cos(2 * np.pi * t / 90.0)
    W, sj, *_ = po.cwt(x, 1 / 12, 1 / 12, 2 / 12, -1)
    power = np.abs(W) ** 2
    assert int(fields["S"]) == sj.size
    assert float(fields["power_sum"]) == pytest.approx(power.sum(), rel=1e-9)
    assert float(fields["peak_scale"]) == pytest.approx(sj[np.unravel_index(power.argmax(), power.shape)[0]], rel=1e-6)
    w = mo.modwt(x, "sym4", 4)
    assert float(fields["modwt_energy_ratio"]) == pytest.approx((w ** 2).sum() / (x ** 2).sum(), abs=1e-10)
    assert float(fields["imodwt_err"]) < 1e-10 and int(fields["launches"]) >= 4
    # the one-call Monte-Carlo significance from C equals the Python binding's (same seed, any GPU count)
    from wavelet_transformer_b200 import _shim
    sig, hist = _shim.wct_significance(0.8, 0.6, 1 / 12, 0.25, 2 / 12, 24, mc_count=24, seed=7, f64=False, return_hist=True)
    assert int(fields["hist_total"]) == int(hist.sum()) and float(fields["sig95_0"]) == pytest.approx(sig[0], abs=1e-9)
    assert int(fields["gpus"]) >= 1


def test_transform_helper_builders(series, shim_nogpu):
    """create_dwt_dict / create_cwt_dict / create_xwt_dict (transform_helpers.py:21-86): host glue
    from DataFrame columns to the dataclasses, NaN handling and the XWT constants included."""
    import pandas as pd
    from src import cwt, dwt, xwt
    from src.utils import transform_helpers as th
    from src.utils.wavelet_helpers import standardize_series
    n = 120
    rng = np.random.default_rng(4)
    a, b = rng.standard_normal(n).cumsum(), rng.standard_normal(n)
    b_nan = b.copy()
    b_nan[:7] = np.nan
    frame = pd.DataFrame({"date": np.arange(n).astype("datetime64[M]"), "a": a, "b": b_nan})
    d = th.create_dwt_dict(frame.dropna(), ["a", "b"])
    assert set(d) == {"a", "b"} and d["a"].mother_wavelet is dwt.MOTHER
    assert d["a"].levels == shim_nogpu.dwt_max_level(n - 7, 8) and d["a"].y_values.size == n - 7
    c = th.create_cwt_dict(frame, ["a", "b"], mother_wavelet=cwt.MOTHER, delta_t=cwt.DT, delta_j=cwt.DJ,
                           initial_scale=cwt.S0, levels=cwt.LEVELS)
    assert c["a"].y_values.size == n and c["b"].y_values.size == n - 7 == c["b"].t_values.size
    assert np.array_equal(c["b"].y_values, standardize_series(b[7:]))
    x = th.create_xwt_dict(frame, [("a", "b")], detrend=False, remove_mean=True)
    item = x[("a", "b")]
    assert item.y1_values.size == n - 7 and np.array_equal(item.y2_values, standardize_series(b[7:], detrend=False, remove_mean=True))
    assert (item.delta_t, item.delta_j, item.initial_scale) == (xwt.DT, xwt.DJ, xwt.S0) and item.levels == xwt.LEVELS

"""GPU parity: MODWT / DWT kernels vs the oracle and the reference-generated goldens."""

import numpy as np
import pytest

from oracle import modwt_oracle as mo
from oracle import pywt_oracle as pw

pytestmark = pytest.mark.gpu


def _bank(name):
    w = pw.Wavelet(name)
    return np.array(w.dec_lo), np.array(w.dec_hi), np.array(w.rec_lo), np.array(w.rec_hi)


def test_modwt_matches_reference_goldens_fp64(shim, modwt_golden):
    g = modwt_golden
    cases = sorted({k.rsplit("|", 1)[0] for k in g})
    assert len(cases) == 12
    for case in cases:
        filt = case.split("|")[1]
        lo, hi, _, _ = _bank(filt)
        x, J = g[f"{case}|x"], int(g[f"{case}|J"])
        w = shim.modwt(x, lo, hi, J, f64=True)
        assert np.abs(w - g[f"{case}|modwt"]).max() <= 1e-10 * max(1.0, np.abs(x).max()), case
        assert np.abs(shim.imodwt(g[f"{case}|modwt"], lo, hi, f64=True) - g[f"{case}|imodwt"]).max() <= 1e-10 * max(1.0, np.abs(x).max())
        mra = shim.modwtmra(g[f"{case}|modwt"], np.vstack(mo.mra_filters(filt, J, x.size)), f64=True)
        assert np.abs(mra - g[f"{case}|mra"]).max() <= 1e-10 * max(1.0, np.abs(x).max()), case


def test_modwt_properties_large_batch(shim):
    """Perfect reconstruction, MRA additivity and energy conservation on a batch."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((257, 1000))
    lo, hi, _, _ = _bank("sym4")
    w = shim.modwt(x, lo, hi, 6, f64=True)
    assert np.abs(shim.imodwt(w, lo, hi, f64=True) - x).max() < 1e-10
    assert np.allclose((w ** 2).sum(axis=(1, 2)), (x ** 2).sum(axis=1), rtol=1e-10)
    mra = shim.modwtmra(w, np.vstack(mo.mra_filters("sym4", 6, 1000)), f64=True)
    assert np.abs(mra.sum(axis=1) - x).max() < 1e-10
    w32 = shim.modwt(x, lo, hi, 6, f64=False)
    assert np.abs(w32 - w).max() < 1e-5


def test_modwt_impulse_support(shim):
    x = np.zeros(256)
    x[10] = 1
    lo, hi, _, _ = _bank("db4")
    w = shim.modwt(x, lo, hi, 3, f64=True)
    for j, last in ((0, 17), (1, 31), (2, 59)):
        nz = np.nonzero(np.abs(w[j]) > 1e-14)[0]
        assert nz.min() == 10 and nz.max() == last


@pytest.mark.parametrize("n", [565, 564, 1333, 64, 23])
@pytest.mark.parametrize("name", ["db4", "sym4", "haar", "db2"])
def test_wavedec_waverec_fp64(shim, n, name):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    lo, hi, rlo, rhi = _bank(name)
    level = pw.dwt_max_level(n, lo.size)
    assert shim.dwt_max_level(n, lo.size) == level
    if level == 0:
        pytest.skip("series too short for this filter")
    ref = pw.wavedec(x, name, level=level)
    packed, lens = shim.wavedec(x, lo, hi, level, f64=True)
    assert list(lens) == [c.size for c in ref]
    assert np.abs(packed - np.concatenate(ref)).max() <= 1e-10
    rec = shim.waverec(packed, lens, rlo, rhi, f64=True)
    rec_ref = pw.waverec(ref, name)
    assert rec.shape == rec_ref.shape and np.abs(rec - rec_ref).max() <= 1e-10
    assert np.abs(rec[: n] - x).max() <= 1e-9 if n % 2 == 0 else rec.size == n + 1


def test_wavedec_cfg_lengths_and_haar_kat(shim, series):
    lo, hi, _, _ = _bank("db4")
    _, lens = shim.wavedec(series["expectation_value"], lo, hi, 6, f64=True)
    assert list(lens) == [15, 15, 24, 41, 76, 146, 286]
    hlo, hhi, _, _ = _bank("haar")
    packed, lens = shim.wavedec(np.arange(1.0, 7.0), hlo, hhi, 1, f64=True)
    assert np.allclose(packed[:3] * np.sqrt(2), [3, 7, 11], atol=1e-14)


def test_wavedec_batch_fp32(shim):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((100, 800))
    lo, hi, rlo, rhi = _bank("db4")
    packed, lens = shim.wavedec(x, lo, hi, 5, f64=False)
    ref = np.stack([np.concatenate(pw.wavedec(r, "db4", level=5)) for r in x])
    assert np.abs(packed - ref).max() < 1e-5
    rec = shim.waverec(packed, lens, rlo, rhi, f64=False)
    assert np.abs(rec - x).max() < 1e-5


@pytest.mark.timeout(120)
@pytest.mark.parametrize("f64", [True, False])
def test_imodwt_tma_path_many_ctas(shim, f64):
    """Regression: 16-byte-aligned rows take the TMA (cp.async.bulk + mbarrier) path, and
    imodwt re-arms the barrier once per level.  With thousands of CTAs a thread lagging a
    whole barrier phase used to deadlock."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((20000, 1024))
    lo, hi, _, _ = _bank("sym4")
    w = shim.modwt(x, lo, hi, 6, f64=f64)
    rec = shim.imodwt(w, lo, hi, f64=f64)
    assert np.abs(rec - x).max() < (1e-10 if f64 else 1e-4)

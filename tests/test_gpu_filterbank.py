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


def test_wavedec_pywt_documentation_examples(shim):
    """The examples printed in the PyWavelets documentation (values in tests/test_oracle_pywt.py)
    through wtb_wavedec / wtb_waverec and through the pywt façade: an anchor outside this repo for
    the symmetric-extension and down-sampling phase of the GPU kernels (blocked and generic)."""
    from test_oracle_pywt import PYWT_DOC_DWT, PYWT_DOC_WAVEDEC
    from wavelet_transformer_b200 import pywt_compat as pywt
    for x, name, cA, cD in PYWT_DOC_DWT:
        lo, hi, rlo, rhi = _bank(name)
        for generic in (False, True):
            packed, lens = shim.wavedec(np.asarray(x, dtype=float), lo, hi, 1, f64=True, generic_only=generic)
            assert list(lens) == [len(cA), len(cD)]
            assert np.allclose(packed, np.concatenate([cA, cD]), atol=5e-9)
            assert np.allclose(shim.waverec(packed, lens, rlo, rhi, f64=True, generic_only=generic), x, atol=1e-12)
    x, name, level, want = PYWT_DOC_WAVEDEC
    got = pywt.wavedec(np.asarray(x, dtype=float), name, level=level)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.allclose(g, w, atol=5e-9)
    assert np.allclose(pywt.waverec(got, name), x, atol=1e-12)
    assert pywt.dwt_max_level(1000, 10) == 6


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


# ---- register-blocked kernels (csrc/filterbank_fast.cu) vs the generic ones and the oracle ----
@pytest.mark.parametrize("name", ["haar", "db2", "db3", "db4", "sym4"])
@pytest.mark.parametrize("n,J", [(23, 6), (64, 4), (565, 6), (1000, 7), (1333, 6), (2048, 6), (3001, 7), (4096, 9)])
def test_modwt_blocked_kernels_fp64(shim, name, n, J):
    """Chains of 9 outputs at stride 2^(j-1): aligned (TMA) and odd (cooperative) rows, tails that
    do not fill a chain, and dilated filters longer than the series (n=23, J=6)."""
    rng = np.random.default_rng(n + J)
    x = rng.standard_normal((3, n))
    lo, hi, _, _ = _bank(name)
    ref = np.stack([mo.modwt(r, name, J) for r in x])
    w = shim.modwt(x, lo, hi, J, f64=True)
    assert np.abs(w - ref).max() <= 1e-12
    assert np.abs(w - shim.modwt(x, lo, hi, J, f64=True, generic_only=True)).max() <= 1e-13
    rec = shim.imodwt(w, lo, hi, f64=True)
    assert np.abs(rec - x).max() <= 1e-10
    assert np.abs(rec - shim.imodwt(w, lo, hi, f64=True, generic_only=True)).max() <= 1e-12
    mra = shim.modwtmra_taps(w, lo, hi, f64=True)
    mra_ref = np.stack([mo.modwtmra(r, name) for r in ref])
    assert np.abs(mra - mra_ref).max() <= 1e-11
    assert np.abs(mra.sum(axis=1) - x).max() <= 1e-10


def test_modwt_unblocked_tap_count_falls_back(shim):
    """L = 10 has no blocked instantiation: the generic kernels serve it, and the taps-only
    MRA entry point builds the periodised equivalent filters itself (modwt.py:56-83)."""
    rng = np.random.default_rng(10)
    x = rng.standard_normal((2, 300))
    lo, hi, _, _ = _bank("db5")
    w = shim.modwt(x, lo, hi, 4, f64=True)
    assert np.abs(w - np.stack([mo.modwt(r, "db5", 4) for r in x])).max() <= 1e-12
    assert np.abs(shim.imodwt(w, lo, hi, f64=True) - x).max() <= 1e-10
    mra = shim.modwtmra_taps(w, lo, hi, f64=True)
    assert np.abs(mra[0] - mo.modwtmra(w[0], "db5")).max() <= 1e-11
    from wavelet_transformer_b200.api import modwt as api
    assert np.abs(api.modwtmra(w[1], "db5") - mo.modwtmra(w[1], "db5")).max() <= 1e-11
    # the same host-built filters serve a covered tap count when the cascade is switched off
    lo4, hi4, _, _ = _bank("sym4")
    w4 = shim.modwt(x, lo4, hi4, 5, f64=True)
    a = shim.modwtmra_taps(w4, lo4, hi4, f64=True)
    b = shim.modwtmra_taps(w4, lo4, hi4, f64=True, generic_only=True)
    c = shim.modwtmra(w4, np.vstack(mo.mra_filters("sym4", 5, 300)), f64=True)
    assert np.abs(a - b).max() <= 1e-12 and np.abs(b - c).max() <= 1e-13


@pytest.mark.parametrize("f64", [True, False])
def test_modwt_blocked_large_batch(shim, f64):
    """TMA bulk stores of every coefficient row, double buffered, thousands of CTAs."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6000, 1024))
    lo, hi, _, _ = _bank("sym4")
    w = shim.modwt(x, lo, hi, 6, f64=f64)
    tol = 1e-12 if f64 else 2e-5
    for b in (0, 17, 5999):
        assert np.abs(w[b] - mo.modwt(x[b], "sym4", 6)).max() <= tol
    assert np.abs(w - shim.modwt(x, lo, hi, 6, f64=f64, generic_only=True)).max() <= (1e-13 if f64 else 1e-5)
    mra = shim.modwtmra_taps(w, lo, hi, f64=f64)
    assert np.abs(mra.sum(axis=1) - x).max() <= (1e-10 if f64 else 1e-4)
    for b in (3, 5998):
        assert np.abs(mra[b] - mo.modwtmra(np.asarray(w[b], dtype=float), "sym4")).max() <= (1e-11 if f64 else 1e-4)


@pytest.mark.parametrize("name", ["haar", "db2", "db3", "db4"])
@pytest.mark.parametrize("n", [23, 64, 565, 1024, 1333])
def test_dwt_blocked_vs_generic_fp64(shim, name, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((5, n))
    lo, hi, rlo, rhi = _bank(name)
    level = pw.dwt_max_level(n, lo.size)
    if level == 0:
        pytest.skip("series too short for this filter")
    packed, lens = shim.wavedec(x, lo, hi, level, f64=True)
    packed_g, _ = shim.wavedec(x, lo, hi, level, f64=True, generic_only=True)
    ref = np.stack([np.concatenate(pw.wavedec(r, name, level=level)) for r in x])
    assert np.abs(packed - ref).max() <= 1e-12 and np.abs(packed - packed_g).max() <= 1e-13
    rec = shim.waverec(packed, lens, rlo, rhi, f64=True)
    rec_g = shim.waverec(packed, lens, rlo, rhi, f64=True, generic_only=True)
    assert rec.shape == rec_g.shape and np.abs(rec - rec_g).max() <= 1e-12
    assert np.abs(rec[:, :n] - x).max() <= 1e-9 if n % 2 == 0 else rec.shape[1] == n + 1


def test_dwt_blocked_large_batch_fp32(shim):
    rng = np.random.default_rng(12)
    x = rng.standard_normal((4000, 1024))
    lo, hi, rlo, rhi = _bank("db4")
    packed, lens = shim.wavedec(x, lo, hi, 7, f64=False)
    for b in (0, 3999):
        assert np.abs(packed[b] - np.concatenate(pw.wavedec(x[b], "db4", level=7))).max() < 2e-5
    rec = shim.waverec(packed, lens, rlo, rhi, f64=False)
    assert np.abs(rec - x).max() < 1e-4


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("n", [1333, 1000, 565])
def test_blocked_kernels_any_row_alignment(shim, f64, n):
    """Rows that start 4, 8 or 12 bytes off a 16-byte boundary (odd lengths, offset device
    pointers): the TMA copies cover the enclosing aligned span and the kernels address the row
    at its shift; results must not depend on where the buffers start."""
    import torch
    dev = torch.device("cuda", 0)
    dt = torch.float64 if f64 else torch.float32
    lo, hi, rlo, rhi = _bank("sym4")
    J, B = 6, 37
    rng = np.random.default_rng(n)
    x = rng.standard_normal((B, n))
    ref_w = shim.modwt(x, lo, hi, J, f64=f64)                      # host path: staging arena, aligned base
    ref_mra = shim.modwtmra_taps(ref_w, lo, hi, f64=f64)
    level = pw.dwt_max_level(n, 8)
    ref_pk, lens = shim.wavedec(x, lo, hi, level, f64=f64)
    ref_rec = shim.waverec(ref_pk, lens, rlo, rhi, f64=f64)
    tol = 1e-12 if f64 else 1e-5
    st = torch.cuda.current_stream().cuda_stream
    for off_in, off_out in ((0, 0), (1, 0), (1, 3), (2, 1), (3, 2)):
        def buf(shape, off):
            count = int(np.prod(shape))
            flat = torch.zeros(count + 8, dtype=dt, device=dev)
            return flat[off:off + count].view(*shape)
        xd = buf((B, n), off_in)
        xd.copy_(torch.from_numpy(x).to(dt))
        wd = buf((B, J + 1, n), off_out)
        shim.modwt_device(xd.data_ptr(), B, n, lo, hi, J, wd.data_ptr(), f64=f64, stream=st)
        assert np.abs(wd.cpu().numpy() - ref_w).max() <= tol
        rec = buf((B, n), off_in)
        shim.imodwt_device(wd.data_ptr(), B, n, lo, hi, J, rec.data_ptr(), f64=f64, stream=st)
        assert np.abs(rec.cpu().numpy() - x).max() <= (1e-10 if f64 else 1e-4)
        mra = buf((B, J + 1, n), off_out)
        shim.modwtmra_taps_device(wd.data_ptr(), B, n, lo, hi, J, mra.data_ptr(), f64=f64, stream=st)
        assert np.abs(mra.cpu().numpy() - ref_mra).max() <= tol
        pk = buf(ref_pk.shape, off_out)
        shim.wavedec_device(xd.data_ptr(), B, n, lo, hi, level, pk.data_ptr(), f64=f64, stream=st)
        assert np.abs(pk.cpu().numpy() - ref_pk).max() <= tol
        xr = buf(ref_rec.shape, off_in)
        shim.waverec_device(pk.data_ptr(), B, lens, rlo, rhi, xr.data_ptr(), f64=f64, stream=st)
        assert np.abs(xr.cpu().numpy() - ref_rec).max() <= tol


def test_modwt_single_step_helpers(shim, series):
    """circular_convolve_d / _s / _mra (modwt.py:81-123): one analysis, synthesis and MRA step
    with the reference's argument conventions, against the oracle's gather formulas."""
    from src import modwt
    x = series["expectation_value"]
    N = x.size
    g, h = np.array(pw.Wavelet("sym4").dec_lo) / np.sqrt(2), np.array(pw.Wavelet("sym4").dec_hi) / np.sqrt(2)
    for j in (1, 3, 6, 8):                                       # 2^7 * 7 > 565: the kernel folds more than once
        w_ref, v_ref = mo._circ_gather(x, h, 2 ** (j - 1), -1), mo._circ_gather(x, g, 2 ** (j - 1), -1)
        w = modwt.circular_convolve_d(h, x, j)
        v = modwt.circular_convolve_d(g, x, j)
        assert np.abs(w - w_ref).max() <= 1e-12 and np.abs(v - v_ref).max() <= 1e-12
        back = modwt.circular_convolve_s(h, g, w, v, j)
        ref_back = mo._circ_gather(w_ref, h, 2 ** (j - 1), 1) + mo._circ_gather(v_ref, g, 2 ** (j - 1), 1)
        assert np.abs(back - ref_back).max() <= 1e-12
        assert np.abs(back - x).max() <= 1e-10                   # one level inverts exactly
    filt = mo.mra_filters("sym4", 4, N)
    wfull = mo.modwt(x, "sym4", 4)
    for r in (0, 3, 4):
        assert np.abs(modwt.circular_convolve_mra(filt[r], wfull[r]) - mo.modwtmra(wfull, "sym4")[r]).max() <= 1e-12


def test_filterbank_fuzz_blocked_vs_generic_vs_oracle(shim):
    """Random lengths, level counts, filters and batch sizes through every filterbank entry
    point, both precisions: the register-blocked kernels (chains, halos, TMA at any alignment,
    long-row MRA chains) against the one-output-per-thread generic kernels on every case and
    against the oracle in FP64."""
    rng = np.random.default_rng(909)
    names = ["haar", "db2", "db3", "db4", "sym4", "db5"]
    for case in range(48):
        name = names[int(rng.integers(len(names)))]
        lo, hi, rlo, rhi = _bank(name)
        n = int(rng.choice([rng.integers(8, 64), rng.integers(64, 700), rng.integers(700, 2100), rng.integers(2100, 5000)]))
        J = int(rng.integers(1, 12))
        batch = int(rng.choice([1, 2, 7, 33]))
        f64 = bool(rng.integers(2))
        tol = 1e-11 if f64 else 3e-5
        x = rng.standard_normal((batch, n))
        tag = f"case {case}: {name} n={n} J={J} batch={batch} f64={f64}"
        w = shim.modwt(x, lo, hi, J, f64=f64)
        wg = shim.modwt(x, lo, hi, J, f64=f64, generic_only=True)
        assert np.abs(w - wg).max() <= tol, tag
        rec = shim.imodwt(w, lo, hi, f64=f64)
        assert np.abs(rec - shim.imodwt(w, lo, hi, f64=f64, generic_only=True)).max() <= tol, tag
        assert np.abs(rec - x).max() <= (1e-9 if f64 else 2e-4), tag
        mra = shim.modwtmra_taps(w, lo, hi, f64=f64)
        assert np.abs(mra - shim.modwtmra_taps(w, lo, hi, f64=f64, generic_only=True)).max() <= 10 * tol, tag
        assert np.abs(mra.sum(axis=1) - x).max() <= (1e-9 if f64 else 2e-4), tag
        if f64:
            assert np.abs(w[0] - mo.modwt(x[0], name, J)).max() <= 1e-11, tag
            assert np.abs(mra[0] - mo.modwtmra(np.asarray(w[0]), name)).max() <= 1e-10, tag
        level = int(rng.integers(0, pw.dwt_max_level(n, lo.size) + 1))
        packed, lens = shim.wavedec(x, lo, hi, level, f64=f64)
        packed_g, _ = shim.wavedec(x, lo, hi, level, f64=f64, generic_only=True)
        assert np.abs(packed - packed_g).max() <= tol, tag + f" level={level}"
        if f64:
            ref = np.concatenate(pw.wavedec(x[0], name, level=level))
            assert np.abs(np.atleast_2d(packed)[0] - ref).max() <= 1e-11, tag + f" level={level}"
        if level > 0:
            back = shim.waverec(packed, lens, rlo, rhi, f64=f64)
            back_g = shim.waverec(packed, lens, rlo, rhi, f64=f64, generic_only=True)
            assert back.shape == back_g.shape and np.abs(back - back_g).max() <= tol, tag + f" level={level}"
            assert np.abs(np.atleast_2d(back)[:, :n] - x).max() <= (1e-9 if f64 else 2e-4) or n % 2, tag

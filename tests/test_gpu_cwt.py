"""GPU parity: batched Morlet CWT (wtb_cwt_morlet) vs the pycwt oracle.

Tolerances (BASELINE.json north_star): FP64 mode rtol 1e-10 (relative to the
plane maximum -- FFT round-off is norm-wise), FP32 mode 1e-4 norm-wise."""

import numpy as np
import pytest

from conftest import normwise_close
from oracle import pycwt_oracle as po

pytestmark = pytest.mark.gpu

DT = 1 / 12


def _oracle_plane(x, dt, dj, s0, J):
    W = po.cwt(x, dt, dj, s0, J, po.Morlet(6))[0]
    return W


@pytest.mark.parametrize("name,dj,J", [("cpi", 1 / 12, 84), ("inflation", 1 / 12, 84), ("expectation", 1 / 8, -1)])
def test_cwt_fp64_sample_series(shim, series, name, dj, J):
    x = series[f"{name}_value"]
    x = (x - x.mean()) / x.std()
    W = _oracle_plane(x, DT, dj, 2 * DT, J)
    power, coef = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=True, want_power=True, want_coef=True)
    assert coef.shape == W.shape and power.shape == W.shape
    scale = np.abs(W).max()
    assert np.abs(coef - W).max() <= 1e-10 * scale
    assert np.abs(power - np.abs(W) ** 2).max() <= 1e-10 * scale ** 2


def test_cwt_fp64_difflog_cpi_cfg1(shim, series):
    """BASELINE cfg1: the app's fallback series 100*diff(log(cpi)), dj=1/12, s0=2dt, J=84."""
    x = 100 * np.diff(np.log(series["cpi_value"]))
    W = _oracle_plane(x, DT, 1 / 12, 2 * DT, 84)
    power, _ = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, 84, f64=True)
    assert power.shape == (85, 1345)
    ref = np.abs(W) ** 2
    assert np.abs(power - ref).max() <= 1e-10 * ref.max()


@pytest.mark.parametrize("n0", [64, 100, 565, 1024, 1346, 2048, 3351, 4096])
@pytest.mark.parametrize("generic", [False, True])
def test_cwt_fp32_lengths(shim, n0, generic):
    rng = np.random.default_rng(n0)
    x = rng.standard_normal((3, n0))
    dj, s0 = 1 / 8, 2 * DT
    power, _ = shim.cwt_morlet(x, DT, dj, s0, -1, f64=False, generic_only=generic)
    for b in range(3):
        ref = np.abs(_oracle_plane(x[b], DT, dj, s0, -1)) ** 2
        ok, err = normwise_close(power[b], ref, 1e-4)
        assert ok, f"n0={n0} row {b}: normwise err {err:.3e}"


def test_cwt_fp32_cfg4_shape(shim):
    """BASELINE cfg4 shape: N=1024, 120 scales (dj=1/12, s0=2dt, J=119), AR(1) g=0.7."""
    rng = np.random.default_rng(1234)
    from scipy.signal import lfilter
    x = lfilter([1.0], [1.0, -0.7], rng.standard_normal((64, 1024 + 64)), axis=1)[:, 64:]
    x *= np.sqrt(1 - 0.49)
    power, _ = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, 119, f64=False)
    power_g, _ = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, 119, f64=False, generic_only=True)
    assert power.shape == (64, 120, 1024)
    for b in range(0, 64, 7):
        ref = np.abs(_oracle_plane(x[b], DT, 1 / 12, 2 * DT, 119)) ** 2
        for got in (power[b], power_g[b]):
            ok, err = normwise_close(got, ref, 1e-4)
            assert ok, f"row {b}: normwise err {err:.3e}"


def test_cwt_batch_equals_single(shim):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((9, 300))
    pb, _ = shim.cwt_morlet(x, DT, 1 / 4, 2 * DT, -1, f64=True)
    for b in range(9):
        p1, _ = shim.cwt_morlet(x[b], DT, 1 / 4, 2 * DT, -1, f64=True)
        assert np.array_equal(p1, pb[b])


def test_cwt_linearity_and_sinusoid_peak(shim):
    """Size-independent properties: linearity of W and the Fourier-period peak."""
    n = 4096
    t = np.arange(n) * DT
    a, b = np.cos(2 * np.pi * t / 2.0), np.sin(2 * np.pi * t / 7.0)
    _, Wa = shim.cwt_morlet(a, DT, 1 / 12, 2 * DT, -1, f64=True, want_power=False, want_coef=True)
    _, Wb = shim.cwt_morlet(b, DT, 1 / 12, 2 * DT, -1, f64=True, want_power=False, want_coef=True)
    _, Wab = shim.cwt_morlet(2 * a - 3 * b, DT, 1 / 12, 2 * DT, -1, f64=True, want_power=False, want_coef=True)
    assert np.abs(Wab - (2 * Wa - 3 * Wb)).max() <= 1e-10 * np.abs(Wab).max()
    _, _, freqs, _ = shim.cwt_axes(n, DT, 1 / 12, 2 * DT, -1)
    peak = 1 / freqs[(np.abs(Wa) ** 2)[:, n // 2].argmax()]
    assert abs(np.log2(peak / 2.0)) < 1 / 12


def test_cwt_coi_mask(shim):
    x = np.random.default_rng(3).standard_normal(500)
    p_mask, _ = shim.cwt_morlet(x, DT, 1 / 8, 2 * DT, -1, f64=True, coi_mask=True)
    p, _ = shim.cwt_morlet(x, DT, 1 / 8, 2 * DT, -1, f64=True)
    _, _, freqs, coi = shim.cwt_axes(500, DT, 1 / 8, 2 * DT, -1)
    outside = (1 / freqs)[:, None] > coi[None, :]
    assert np.isnan(p_mask[outside]).all()
    assert np.array_equal(p_mask[~outside], p[~outside])


@pytest.mark.parametrize("n0,batch", [(400, 5), (512, 4), (700, 3), (1024, 3), (1346, 13), (2048, 12), (3351, 3), (4096, 2),
                                      (3351, 13), (4096, 12)])
def test_cwt_coi_mask_fused_fast_kernels(shim, n0, batch):
    """north_star (1): the cone-of-influence mask is fused into the FP32 fast kernels' store loops
    (nfft 512 two-series, 1024, 2048 warp pairs, 4096 register rows and warp quads).  Inside the cone the power is bit-equal to
    the unmasked fast kernel, outside it is NaN, and the mask itself equals the oracle's
    `period > coi` (the generic kernel's, too)."""
    x = np.random.default_rng(n0).standard_normal((batch, n0))
    launches = shim.kernel_launches()
    p_mask, _ = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, -1, f64=False, coi_mask=True)
    per_call = shim.kernel_launches() - launches
    p, _ = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, -1, f64=False)
    assert shim.kernel_launches() - launches == 2 * per_call          # same kernels with and without the mask
    g_mask, _ = shim.cwt_morlet(x, DT, 1 / 12, 2 * DT, -1, f64=False, coi_mask=True, generic_only=True)
    _, freqs, coi = po.cwt(x[0], DT, 1 / 12, 2 * DT, -1)[1:4]
    outside = (1 / freqs)[:, None] > coi[None, :]
    assert outside.any() and (~outside).any()
    for b in range(batch):
        assert np.isnan(p_mask[b][outside]).all()
        assert np.array_equal(p_mask[b][~outside], p[b][~outside])
        assert np.array_equal(np.isnan(g_mask[b]), outside)


def test_cwt_against_the_time_domain_definition(shim):
    """The GPU transform against Torrence & Compo (1998) eq. 2 evaluated directly (no FFT, no
    oracle): W_n(s) = sum_n' x_n' conj(psi((n' - n) dt / s)), psi(eta) = sqrt(dt/s) pi^(-1/4)
    exp(i 6 eta) exp(-eta^2 / 2).  Interior samples of resolved scales agree to round-off."""
    rng = np.random.default_rng(11)
    n, dt = 512, 0.25
    x = rng.standard_normal(n).cumsum()
    x = (x - x.mean()) / x.std()
    _, W = shim.cwt_morlet(x, dt, 1 / 4, 4 * dt, 16, f64=True, want_power=False, want_coef=True)
    p32, _ = shim.cwt_morlet(np.stack([x] * 3), dt, 1 / 4, 4 * dt, 16, f64=False)
    t = np.arange(n)
    for j in (2, 6, 10):
        s = 4 * dt * 2.0 ** (j / 4)
        for m in (200, 256, 300):
            eta = (t - m) * dt / s
            psi = np.sqrt(dt / s) * np.pi ** -0.25 * np.exp(1j * 6.0 * eta) * np.exp(-0.5 * eta ** 2)
            direct = np.sum(x * np.conj(psi))
            assert abs(W[j, m] - direct) <= 1e-10 * np.abs(W[j]).max()
            assert abs(p32[0, j, m] - abs(direct) ** 2) <= 1e-4 * (abs(direct) ** 2 + p32[0].max())


def test_cwt_rejects_bad_arguments(shim):
    x = np.zeros(100)
    with pytest.raises((ValueError, RuntimeError)):
        shim.cwt_morlet(x, DT, 1 / 8, 2 * DT, -1, nfft=64)       # shorter than the series
    with pytest.raises((ValueError, RuntimeError)):
        shim.cwt_morlet(x, DT, -1.0, 2 * DT, -1)                 # dj <= 0
    with pytest.raises((ValueError, RuntimeError)):
        shim.cwt_morlet(np.zeros(20000), DT, 1 / 8, 2 * DT, -1, f64=True)  # beyond the smem FFT


# ---- other pycwt mothers and the inverse transform -----------------------
@pytest.mark.parametrize("mother", ["paul4", "paul3", "dog2", "dog3", "dog6"])
def test_cwt_other_mothers_fp64_and_fp32(shim, series, mother):
    """Paul / DOG daughters (real and imaginary prefactors -i^m) through the generic kernel."""
    from wavelet_transformer_b200 import pycwt_compat as wavelet
    kind, m = mother[:-1], int(mother[-1])
    ref_m = po.Paul(m) if kind == "paul" else po.DOG(m)
    eng_m = wavelet.Paul(m) if kind == "paul" else wavelet.DOG(m)
    x = series["expectation_value"]
    x = (x - x.mean()) / x.std()
    W_ref, sj_ref, fr_ref, coi_ref, _, _ = po.cwt(x, DT, 1 / 8, -1, -1, ref_m)
    W, sj, fr, coi, _, _ = wavelet.cwt(x, DT, 1 / 8, -1, -1, eng_m)
    assert W.shape == W_ref.shape and W.dtype == np.complex128
    assert np.allclose(sj, sj_ref, rtol=1e-13) and np.allclose(fr, fr_ref, rtol=1e-13)
    assert np.allclose(coi, coi_ref, rtol=1e-13)
    assert np.abs(W - W_ref).max() <= 1e-10 * np.abs(W_ref).max()
    code = shim.PAUL if kind == "paul" else shim.DOG
    p32, _ = shim.cwt(x, DT, 1 / 8, -1, -1, code, m, f64=False)
    ok, err = normwise_close(p32, np.abs(W_ref) ** 2, 1e-4)
    assert ok, err


def test_other_mothers_significance_xwt_and_large_orders(shim, series):
    """pycwt.significance and pycwt.xwt work for every mother (only flambda, dofmin,
    gamma enter; xwt needs no smoothing), so run_cwt's non-Morlet branch must survive its default
    calculate_significance=True; and z^m of a large-order Paul / DOG daughter must not overflow FP32."""
    from src import cwt as cwt_mod
    from wavelet_transformer_b200 import pycwt_compat as wavelet
    y = 100 * np.diff(np.log(series["cpi_value"]))
    t = series["cpi_days"].astype("datetime64[D]")[1:]
    alpha = po.ar1(y)[0]
    for eng_m, ref_m in ((wavelet.Paul(4), po.Paul(4)), (wavelet.DOG(2), po.DOG(2)), (wavelet.MexicanHat(), po.DOG(2))):
        data = cwt_mod.DataForCWT(t, y, eng_m, cwt_mod.DT, cwt_mod.DJ, cwt_mod.S0, cwt_mod.LEVELS)
        res = cwt_mod.run_cwt(data)
        W, sj, *_ = po.cwt(y, DT, 1 / 12, 2 * DT, 84.0, ref_m)
        signif, _ = po.significance(1.0, DT, sj, 0, alpha, significance_level=0.95, wavelet=ref_m)
        want = np.abs(W) ** 2 / signif[:, None]
        assert np.abs(res.significance_levels - want).max() <= 1e-9 * want.max()
    # xwt with a non-Morlet mother: W1 conj(W2) of the two transforms, red-noise significance from dofmin
    a = series["pair_expectation"]          # (the merged inflation series makes pycwt.ar1 raise, as in the reference)
    b = a[::-1].copy()
    W12, coi, freq, signif = wavelet.xwt(a, b, DT, dj=1 / 8, s0=2 * DT, J=-1, wavelet=wavelet.Paul(4))
    R12, rcoi, rfreq, rsig = po.xwt(a, b, DT, dj=1 / 8, s0=2 * DT, J=-1, wavelet=po.Paul(4))
    assert np.abs(W12 - R12).max() <= 1e-10 * np.abs(R12).max()
    assert np.allclose(signif, rsig, rtol=1e-12) and np.allclose(freq, rfreq, rtol=1e-13)
    # large orders in FP32: finite everywhere and at the FP32 gate against the FP64 kernel
    x = (a - a.mean()) / a.std()
    for code, m in ((shim.PAUL, 20), (shim.DOG, 30)):
        p32, _ = shim.cwt(x, DT, 1 / 8, -1, -1, code, m, f64=False)
        p64, _ = shim.cwt(x, DT, 1 / 8, -1, -1, code, m, f64=True)
        assert np.isfinite(p32).all() and np.isfinite(p64).all()
        ok, err = normwise_close(p32, p64, 1e-4)
        assert ok, (code, m, err)


def test_icwt_and_mother_names(shim, series):
    from wavelet_transformer_b200 import pycwt_compat as wavelet
    x = series["inflation_value"]
    x = (x - x.mean()) / x.std()
    for name, ref_m in (("morlet", po.Morlet(6)), ("paul", po.Paul(4)), ("mexicanhat", po.DOG(2))):
        W, sj, *_ = wavelet.cwt(x, DT, 1 / 12, -1, -1, name)
        W_ref, sj_ref, *_ = po.cwt(x, DT, 1 / 12, -1, -1, ref_m)
        assert np.abs(W - W_ref).max() <= 1e-10 * np.abs(W_ref).max()
        rec = wavelet.icwt(W, sj, DT, 1 / 12, name)
        rec_ref = po.icwt(W_ref, sj_ref, DT, 1 / 12, ref_m)
        assert rec.shape == x.shape and np.abs(rec - rec_ref).max() <= 1e-10 * np.abs(rec_ref).max()
        assert np.abs(wavelet.icwt(W.T, sj, DT, 1 / 12, name) - rec).max() == 0      # pycwt accepts the transpose
    with pytest.raises(ValueError):
        wavelet.cwt(x, DT, wavelet="haar")
    with pytest.raises(NotImplementedError):
        wavelet.wct(x, x[::-1].copy(), DT, sig=False, wavelet=wavelet.Paul(4))
    batch = np.stack([np.asarray(wavelet.cwt(x * k, DT, 1 / 4, -1, -1, "dog")[0]) for k in (1.0, 2.0)])
    sj4 = wavelet.cwt(x, DT, 1 / 4, -1, -1, "dog")[1]
    out = shim.icwt(batch, sj4, 1.0, f64=True)                                       # batched entry point
    assert np.allclose(out[1], 2 * out[0], rtol=1e-12)


@pytest.mark.parametrize("n0,batch", [(4096, 2), (3351, 3), (2049, 1)])
def test_cwt_fp32_nfft4096_register_rows(shim, n0, batch):
    """Series of 2049..4096 samples take the radix-16 register-FFT rows (two series per CTA in
    the two FFMA2 lanes; an odd batch ends with a half-empty CTA): oracle and generic parity."""
    rng = np.random.default_rng(n0)
    x = rng.standard_normal((batch, n0))
    dj, J = 1 / 8, 65
    power, coef = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False, want_power=True, want_coef=True)
    gen, _ = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False, generic_only=True)
    assert power.shape == (batch, J + 1, n0) and coef.shape == power.shape
    for b in range(batch):
        W = po.cwt(x[b], DT, dj, 2 * DT, J, po.Morlet(6))[0]
        ok, err = normwise_close(power[b], np.abs(W) ** 2, 1e-4)
        assert ok, err
        assert np.abs(coef[b] - W).max() <= 1e-4 * np.abs(W).max()
        ok, err = normwise_close(power[b], gen[b], 1e-4)
        assert ok, err


@pytest.mark.parametrize("n0,dj,J", [(1346, 1 / 12, 84), (1345, 1 / 12, 84), (2048, 1 / 8, 70), (1025, 1 / 4, -1)])
def test_cwt_fp32_nfft2048_interleaved_passes(shim, monkeypatch, n0, dj, J):
    """Batches of 1025..2048-sample series (BASELINE cfg1's 1346-month CPI shape) take the
    two-warps-per-row kernel (even / odd bins, one 1024-point transform each), even and odd row
    lengths.  Oracle parity on sampled series, generic-kernel parity on all.  1100 series give
    every CTA seven or eight series: its three-slot spectrum ring is refilled twice."""
    rng = np.random.default_rng(n0)
    batch = 1100 if n0 == 1346 else 300
    x = rng.standard_normal((batch, n0)).cumsum(axis=1) * 0.05 + rng.standard_normal((batch, n0))
    power, _ = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False)
    gen, _ = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False, generic_only=True)
    assert power.shape == gen.shape and power.shape[0] == batch and power.shape[2] == n0
    for b in range(batch):
        ok, err = normwise_close(power[b], gen[b], 1e-4)
        assert ok, f"series {b}: {err:.3e} vs the generic kernel"
    for b in (0, 147, 148, batch - 1):
        ref = np.abs(_oracle_plane(x[b], DT, dj, 2 * DT, J)) ** 2
        ok, err = normwise_close(power[b], ref, 1e-4)
        assert ok, f"series {b}: {err:.3e} vs the oracle"
    # below the default threshold the generic kernel serves the call
    monkeypatch.delenv("WTB_CWT_MIN_BATCH")
    small, _ = shim.cwt_morlet(x[:5], DT, dj, 2 * DT, J, f64=False)
    assert np.array_equal(small, gen[:5])



@pytest.mark.parametrize("n0,dj,J,batch", [(3351, 1 / 8, 65, 300), (4096, 1 / 12, 100, 150), (2049, 1 / 4, -1, 12),
                                            (3000, 1 / 6, 40, 450)])
def test_cwt_fp32_nfft4096_warp_quads(shim, monkeypatch, n0, dj, J, batch):
    """Batches (at least four series per SM by default; the tests lower the threshold to 1) of 2049..4096-sample series take the four-warps-per-row kernel (bins k = 4 j + r per
    warp, radix-4 combine through the transpose buffers).  Oracle parity on sampled series, generic-kernel
    parity on all; 300 and 450 series give every CTA two to four series (its two-slot spectrum ring is
    refilled), 12 series split their rows over CTAs."""
    rng = np.random.default_rng(n0)
    x = rng.standard_normal((batch, n0)).cumsum(axis=1) * 0.05 + rng.standard_normal((batch, n0))
    power, _ = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False)
    gen, _ = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False, generic_only=True)
    assert power.shape == gen.shape and power.shape[0] == batch and power.shape[2] == n0
    for b in range(batch):
        ok, err = normwise_close(power[b], gen[b], 1e-4)
        assert ok, f"series {b}: {err:.3e} vs the generic kernel"
    for b in (0, batch // 2, batch - 1):
        ref = np.abs(_oracle_plane(x[b], DT, dj, 2 * DT, J)) ** 2
        ok, err = normwise_close(power[b], ref, 1e-4)
        assert ok, f"series {b}: {err:.3e} vs the oracle"
    # a series' numbers do not depend on the batch it came in (split over CTAs or not)
    first, _ = shim.cwt_morlet(x[:13], DT, dj, 2 * DT, J, f64=False)
    assert np.array_equal(first, power[:13])
    # below the default threshold (four series per SM) the register-row kernel (k_cwt_rows_4096) serves the call
    monkeypatch.delenv("WTB_CWT_MIN_BATCH")
    small, _ = shim.cwt_morlet(x[:5], DT, dj, 2 * DT, J, f64=False)
    assert not np.array_equal(small, power[:5])
    for b in range(5):
        ok, err = normwise_close(small[b], gen[b], 1e-4)
        assert ok, f"series {b}: {err:.3e} (register rows) vs the generic kernel"


def test_cwt_fp32_dispatch_fuzz(shim):
    """Random transform parameters through every FP32 dispatch target (1024 warp kernel, the
    2048 two-pass kernel, the 4096 register rows, generic): unusual f0, sampling steps, smallest
    scales below the Nyquist period and very coarse / fine dj move the daughter's band edges, which
    is what the pruned transforms key on.  Reference = the FP64 generic kernel (itself oracle-gated)."""
    rng = np.random.default_rng(2026)
    cases = []
    for n0 in (1024, 700, 1346, 2048, 1500, 3351, 4096, 300):
        for _ in range(3):
            dt = float(rng.choice([1 / 12, 0.25, 1.0, 3.0]))
            dj = float(rng.choice([1 / 2, 1 / 4, 1 / 7, 1 / 12, 1 / 16]))
            s0 = dt * float(rng.choice([0.5, 1.0, 2.0, 5.0]))
            f0 = float(rng.choice([4.0, 6.0, 6.0, 8.5, 12.0]))
            jmax = int(np.floor(np.log2(n0 * dt / s0) / dj))
            J = int(rng.integers(max(1, jmax // 2), min(jmax, 120) + 1))
            cases.append((n0, dt, dj, s0, f0, J))
    for n0, dt, dj, s0, f0, J in cases:
        batch = 140 if n0 > 1024 and n0 <= 2048 else 5      # the 2048 kernel starts at 128 series
        x = rng.standard_normal((batch, n0)) + 0.02 * rng.standard_normal((batch, n0)).cumsum(axis=1)
        got, _ = shim.cwt_morlet(x, dt, dj, s0, J, f0, f64=False)
        ref, _ = shim.cwt_morlet(x[:5], dt, dj, s0, J, f0, f64=True, generic_only=True)
        for b in range(5):
            ok, err = normwise_close(got[b], ref[b], 1e-4)
            assert ok, f"n0={n0} dt={dt} dj={dj:.4f} s0={s0} f0={f0} J={J} series {b}: {err:.3e}"


@pytest.mark.parametrize("n0,batch", [(512, 6), (400, 7), (257, 1), (300, 33)])
def test_cwt_fp32_nfft512_two_series_per_warp(shim, n0, batch):
    """Series of 257..512 samples share the 1024-point warp kernel two at a time (even / odd bins
    of the interleaved spectrum, sum / difference of the two output halves); an odd batch leaves
    the last warp with one series.  Oracle and generic-kernel parity, and pairing must not leak:
    a series gives the same plane whatever its partner is."""
    rng = np.random.default_rng(n0 + batch)
    x = rng.standard_normal((batch, n0)) * 10.0 ** rng.uniform(-4, 5, size=(batch, 1))   # amplitudes over 9 decades
    dj, J = 1 / 12, int(np.floor(np.log2(n0 * DT / (2 * DT)) * 12))
    power, _ = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False)
    gen, _ = shim.cwt_morlet(x, DT, dj, 2 * DT, J, f64=False, generic_only=True)
    assert power.shape == (batch, J + 1, n0)
    for b in range(batch):
        ok, err = normwise_close(power[b], gen[b], 1e-4)
        assert ok, f"series {b}: {err:.3e} vs the generic kernel"
    for b in sorted({0, batch - 1}):
        ref = np.abs(_oracle_plane(x[b], DT, dj, 2 * DT, J)) ** 2
        ok, err = normwise_close(power[b], ref, 1e-4)
        assert ok, f"series {b}: {err:.3e} vs the oracle"
    if batch >= 3:
        swapped, _ = shim.cwt_morlet(x[[0, 2, 1]], DT, dj, 2 * DT, J, f64=False)
        ok, err = normwise_close(swapped[0], power[0], 5e-6)
        assert ok, f"series 0 changed with its partner: {err:.3e}"


@pytest.mark.parametrize("n0,below,above", [(400, 19, 20), (1346, 11, 12)])
def test_cwt_fp32_small_batches_take_the_generic_kernel(shim, monkeypatch, n0, below, above):
    """Small batches: the warp kernels split a series' rows over up to 16 warps, which puts the
    1024 kernel ahead of the generic one from a single series on; the two-series-per-warp (512)
    and two-pass (2048) variants have measured break-evens of 20 and 12 series, below which the
    dispatch must pick the generic kernel."""
    monkeypatch.delenv("WTB_CWT_MIN_BATCH")
    rng = np.random.default_rng(n0)
    x = rng.standard_normal((above, n0))
    J = 60
    gen, _ = shim.cwt_morlet(x, DT, 1 / 8, 2 * DT, J, f64=False, generic_only=True)
    lo, _ = shim.cwt_morlet(x[:below], DT, 1 / 8, 2 * DT, J, f64=False)
    hi, _ = shim.cwt_morlet(x, DT, 1 / 8, 2 * DT, J, f64=False)
    assert np.array_equal(lo, gen[:below])
    assert not np.array_equal(hi, gen)
    ok, err = normwise_close(hi[0], gen[0], 1e-4)
    assert ok, err


@pytest.mark.parametrize("n0,S,dj", [(1024, 64, 1 / 8), (565, 66, 1 / 8), (1024, 7, 1), (2048, 30, 1 / 4), (400, 92, 1 / 12)])
def test_cwt_fp32_row_split_is_invisible(shim, n0, S, dj):
    """Fewer series than warps on the machine: each series' scale rows are dealt out to up to 16
    warps (row c, c + split, ...), every one repeating the forward transform.  The split depends
    on the batch size only, so a series must come out bit-identical at any batch size (2400
    series: no split), and equal to the oracle."""
    rng = np.random.default_rng(n0 + S)
    J = S - 1
    big = rng.standard_normal((2400, n0))
    ref_big, _ = shim.cwt_morlet(big, DT, dj, 2 * DT, J, f64=False)
    # 257..512 samples: two series share a warp, so only whole pairs keep their rounding
    for batch in ((2, 4, 40, 150, 600) if n0 <= 512 else (1, 3, 40, 149, 600)):
        power, _ = shim.cwt_morlet(big[:batch], DT, dj, 2 * DT, J, f64=False)
        assert np.array_equal(power, ref_big[:batch]), f"batch {batch}"
    ref = np.abs(_oracle_plane(big[2], DT, dj, 2 * DT, J)) ** 2
    ok, err = normwise_close(ref_big[2], ref, 1e-4)
    assert ok, err

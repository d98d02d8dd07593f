"""GPU: BASELINE.json's full per-GPU sizes, checked through size-independent properties
(the oracle cannot finish these sizes): Parseval per scale row, batch-position invariance,
histogram totals, partition invariance."""

import numpy as np
import pytest

from oracle import pycwt_oracle as po

pytestmark = pytest.mark.gpu

DT = 1 / 12


@pytest.mark.timeout(300)
def test_cfg4_shard_parseval_and_batch_invariance(shim):
    """cfg4 per-GPU shard: 125 000 series x N=1024 x 120 scales, FP32, device resident."""
    import torch
    from wavelet_transformer_b200 import engine
    dev = torch.device("cuda", 0)
    B, n0, J = 125_000, 1024, 119
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    x = torch.randn((B, n0), generator=g, device=dev, dtype=torch.float32)
    power = torch.empty((B, J + 1, n0), dtype=torch.float32, device=dev)
    engine.cwt_power_resident(x, power, DT, 1 / 12, 2 * DT, J)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(power[::997]).all())
    # Parseval per (series, scale): sum_t |W|^2 = (1/N) sum_k |X^[k] psi_s[k]|^2
    _, scales, _, _ = shim.cwt_axes(n0, DT, 1 / 12, 2 * DT, J)
    idx = torch.tensor([0, 1, 4242, 62_499, 99_991, B - 1], device=dev)
    xs = x[idx].double()
    X = torch.fft.fft(xs, dim=1)
    w = 2 * np.pi * torch.fft.fftfreq(n0, DT, device=dev, dtype=torch.float64)
    s = torch.tensor(scales, device=dev)[:, None]
    psi = torch.sqrt(s * w[1] * n0) * np.pi ** -0.25 * torch.exp(-0.5 * (s * w[None, :] - 6.0) ** 2)
    expect = (X.abs()[:, None, :] ** 2 * psi[None] ** 2).sum(dim=2) / n0
    got = power[idx].double().sum(dim=2)
    assert float(((got - expect).abs() / expect).max()) < 2e-4
    # position in the batch does not matter: same rows recomputed alone are bit-identical
    alone = torch.empty((idx.numel(), J + 1, n0), dtype=torch.float32, device=dev)
    engine.cwt_power_resident(x[idx].contiguous(), alone, DT, 1 / 12, 2 * DT, J)
    torch.cuda.synchronize()
    assert torch.equal(alone, power[idx])
    # and one of them against the oracle
    ref = np.abs(po.cwt(xs[2].cpu().numpy(), DT, 1 / 12, 2 * DT, J)[0]) ** 2
    assert np.abs(power[4242].cpu().numpy() - ref).max() <= 1e-4 * ref.max()


@pytest.mark.timeout(300)
def test_cfg1_shape_at_batch_scale(shim):
    """BASELINE cfg1's shape (1346 monthly samples -> nfft 2048, 85 scales) as a 30 000-series
    device-resident batch: the two-pass-per-row kernel against the generic kernel on sampled
    series, batch-position invariance through duplicated series, and the oracle."""
    import torch
    from wavelet_transformer_b200 import engine
    dev = torch.device("cuda", 0)
    B, n0, J = 30_000, 1346, 84
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    x = torch.randn((B, n0), generator=g, device=dev, dtype=torch.float32)
    x = x + 0.1 * torch.cumsum(x, dim=1)
    x[20_001] = x[0]                                    # the same series at two batch positions
    x[B - 1] = x[4242]
    power = torch.empty((B, J + 1, n0), dtype=torch.float32, device=dev)
    engine.cwt_power_resident(x, power, DT, 1 / 12, 2 * DT, J)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(power[::499]).all())
    assert torch.equal(power[0], power[20_001]) and torch.equal(power[4242], power[B - 1])
    idx = torch.tensor([0, 1, 4242, 14_999, 29_998, B - 1], device=dev)
    alone = torch.empty((idx.numel(), J + 1, n0), dtype=torch.float32, device=dev)
    engine.cwt_power_resident(x[idx].contiguous(), alone, DT, 1 / 12, 2 * DT, J, generic_only=True)
    torch.cuda.synchronize()
    scale = alone.amax(dim=(1, 2), keepdim=True)
    assert float(((alone - power[idx]).abs() / scale).max()) <= 1e-4
    ref = np.abs(po.cwt(x[14_999].double().cpu().numpy(), DT, 1 / 12, 2 * DT, J)[0]) ** 2
    assert np.abs(power[14_999].cpu().numpy() - ref).max() <= 1e-4 * ref.max()


@pytest.mark.timeout(300)
def test_cfg5_histogram_totals_and_partition(shim):
    """cfg5 shape (N=3351 -> 4096, 66 scales): every reliable sample of every realisation is
    binned exactly once, and sharding the realisations does not change the histogram."""
    dj, s0, J = 1 / 8, 2 * DT, 65
    N, maxscale = shim.wct_mc_geometry(DT, dj, s0, J)
    _, _, freqs, coi = shim.cwt_axes(N, DT, dj, s0, J)
    per_real = int(((1 / freqs)[:maxscale, None] <= coi[None, :]).sum())
    mc = 1500
    full = shim.wct_mc_hist(0.989, 0.966, DT, dj, s0, J, mc_first=0, mc_count=mc, seed=2024, f64=False)
    assert int(full.sum()) == mc * per_real
    assert int(full[maxscale:].sum()) == 0
    parts = sum(shim.wct_mc_hist(0.989, 0.966, DT, dj, s0, J, mc_first=a, mc_count=b - a, seed=2024, f64=False)
                for a, b in ((0, 187), (187, 750), (750, 1500)))
    assert np.array_equal(full, parts)
    sig = shim.wct_sig_from_hist(full, maxscale, 0.95, shim.row_has_points(DT, dj, s0, J))
    assert np.isfinite(sig[:maxscale]).all() and ((sig[:maxscale] > 0.55) & (sig[:maxscale] < 0.999)).all()
    # thresholds are a smooth function of scale away from the largest scales
    assert np.abs(np.diff(sig[5:50])).max() < 0.05


@pytest.mark.timeout(300)
@pytest.mark.parametrize("n", [1024, 1333])
def test_cfg2_batched_modwt_properties(shim, n):
    """cfg2 shape at batch scale: 100 000 series x N (the reference's 1333-sample series and a
    power of two), LA8, J = 6, FP64, device resident.  Energy conservation, perfect
    reconstruction, MRA additivity, DWT round trip, and oracle parity on sampled rows."""
    import torch
    from oracle import modwt_oracle as mo
    from oracle import pywt_oracle as pw
    dev = torch.device("cuda", 0)
    B, J = 100_000, 6
    w8 = pw.Wavelet("sym4")
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    x = torch.randn((B, n), generator=g, device=dev, dtype=torch.float64)
    w = torch.empty((B, J + 1, n), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    shim.modwt_device(x.data_ptr(), B, n, w8.dec_lo, w8.dec_hi, J, w.data_ptr(), f64=True, stream=st)
    torch.cuda.synchronize()
    energy = (w ** 2).sum(dim=(1, 2))
    assert float(((energy - (x ** 2).sum(dim=1)).abs() / energy).max()) < 1e-10   # the tabulated taps are orthonormal to ~1e-12
    rec = torch.empty_like(x)
    shim.imodwt_device(w.data_ptr(), B, n, w8.dec_lo, w8.dec_hi, J, rec.data_ptr(), f64=True, stream=st)
    assert float((rec - x).abs().max()) < 1e-10
    mra = torch.empty_like(w)
    shim.modwtmra_taps_device(w.data_ptr(), B, n, w8.dec_lo, w8.dec_hi, J, mra.data_ptr(), f64=True, stream=st)
    assert float((mra.sum(dim=1) - x).abs().max()) < 1e-10
    for b in (0, 31_337, B - 1):
        xb = x[b].cpu().numpy()
        ref = mo.modwt(xb, "sym4", J)
        assert np.abs(w[b].cpu().numpy() - ref).max() < 1e-12
        assert np.abs(mra[b].cpu().numpy() - mo.modwtmra(ref, "sym4")).max() < 1e-11
    level = pw.dwt_max_level(n, 8)
    lens = shim.dwt_coeff_lens(n, 8, level)
    pk = torch.empty((B, int(lens.sum())), dtype=torch.float64, device=dev)
    shim.wavedec_device(x.data_ptr(), B, n, w8.dec_lo, w8.dec_hi, level, pk.data_ptr(), f64=True, stream=st)
    xr = torch.empty((B, shim.waverec_len(lens, 8)), dtype=torch.float64, device=dev)
    shim.waverec_device(pk.data_ptr(), B, lens, w8.rec_lo, w8.rec_hi, xr.data_ptr(), f64=True, stream=st)
    off = xr.shape[1] - n                     # odd lengths reconstruct one sample long (dwt.py:82-85)
    assert float((xr[:, : n] - x).abs().max()) < 1e-9 if off == 0 else xr.shape[1] == n + 1
    b = 77_777
    assert np.abs(pk[b].cpu().numpy() - np.concatenate(pw.wavedec(x[b].cpu().numpy(), "sym4", level=level))).max() < 1e-12

"""Known-answer pins for the PyWavelets restatement (parity unpinned)."""

import numpy as np
import pytest

from oracle import pywt_oracle as pw


@pytest.mark.parametrize("name", ["haar", "db2", "db3", "db4", "db5", "sym4"])
def test_qmf_identities(name):
    w = pw.Wavelet(name)
    g, h = np.array(w.dec_lo), np.array(w.dec_hi)
    L = g.size
    assert g.sum() == pytest.approx(np.sqrt(2), abs=1e-12)
    assert (g ** 2).sum() == pytest.approx(1, abs=1e-11)
    assert h.sum() == pytest.approx(0, abs=1e-11)
    for m in range(1, L // 2):
        assert np.dot(g[2 * m:], g[:L - 2 * m]) == pytest.approx(0, abs=1e-11)
    assert np.dot(g, h) == pytest.approx(0, abs=1e-12)
    # vanishing moments: L/2 of them
    k = np.arange(L)
    for p in range(L // 2):
        assert np.dot(h, k ** p) == pytest.approx(0, abs=1e-8)
    assert w.rec_lo == w.dec_lo[::-1] and w.rec_hi == w.dec_hi[::-1] and w.dec_len == L


def test_db4_table_signs():
    w = pw.Wavelet("db4")
    assert w.dec_hi[0] == pytest.approx(-0.23037781330885523)
    assert w.dec_hi[1] == pytest.approx(0.7148465705525415)


def test_haar_known_answer():
    cA, cD = pw.dwt(np.arange(1.0, 7.0), "haar")
    assert np.allclose(cA * np.sqrt(2), [3, 7, 11])
    assert np.allclose(cD * np.sqrt(2), [-1, -1, -1])
    assert np.allclose(pw.idwt(cA, cD, "haar"), np.arange(1.0, 7.0))


def test_levels_and_lengths(series):
    assert pw.dwt_max_level(565, 8) == 6 and pw.dwt_max_level(1333, 8) == 7
    assert pw.dwt_max_level(6, 8) == 0
    c = pw.wavedec(series["expectation_value"], "db4", level=6)
    assert [a.size for a in c] == [15, 15, 24, 41, 76, 146, 286]


@pytest.mark.parametrize("n", [64, 564, 1000])
@pytest.mark.parametrize("name", ["haar", "db2", "db3", "db4", "db5", "sym4"])
def test_perfect_reconstruction_even(n, name):
    x = np.random.default_rng(n).standard_normal(n)
    rec = pw.waverec(pw.wavedec(x, name), name)
    assert rec.size == n and np.abs(rec - x).max() < 1e-10


def test_odd_length_reconstructs_one_sample_long(series):
    x = series["expectation_value"]          # n = 565
    rec = pw.waverec(pw.wavedec(x, "db4"), "db4")
    assert rec.size == 566
    assert np.abs(rec[:-1] - x).max() < 1e-9  # the extra sample is at the END of the raw reconstruction


def test_symmetric_extension_longer_than_signal():
    x = np.arange(3.0)
    e = pw._sym_ext(x, 7)
    assert e.tolist() == [0, 0, 1, 2, 2, 1, 0, 0, 1, 2, 2, 1, 0, 0, 1, 2, 2]

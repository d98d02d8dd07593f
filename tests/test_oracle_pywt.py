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


# ---- published known answers: the examples printed in the PyWavelets documentation ------------
# (https://pywavelets.readthedocs.io: "Discrete Wavelet Transform (DWT)" and "Multilevel
# decomposition using wavedec"; values as printed there, 8 decimals).  They pin the phase of the
# symmetric extension and of the down-sampling (SURVEY App. B "medium confidence") from outside.
PYWT_DOC_DWT = [
    # (x, wavelet, cA, cD)
    ([1, 2, 3, 4, 5, 6], "db1", [2.12132034, 4.94974747, 7.77817459], [-0.70710678, -0.70710678, -0.70710678]),
    ([3, 7, 1, 1, -2, 5, 4, 6], "db2", [5.65685425, 7.39923721, 0.22414387, 3.33677403, 7.77817459],
     [-2.44948974, -1.60368225, -4.44140056, -0.41361256, 1.22474487]),
]
PYWT_DOC_WAVEDEC = ([1, 2, 3, 4, 5, 6, 7, 8], "db1", 2,
                    [[5.0, 13.0], [-2.0, -2.0], [-0.70710678, -0.70710678, -0.70710678, -0.70710678]])


@pytest.mark.parametrize("x,name,cA,cD", PYWT_DOC_DWT)
def test_pywt_documentation_dwt_examples(x, name, cA, cD):
    a, d = pw.dwt(np.asarray(x, dtype=float), name)
    assert np.allclose(a, cA, atol=5e-9) and np.allclose(d, cD, atol=5e-9)
    assert np.allclose(pw.idwt(a, d, name), x, atol=1e-12)


def test_pywt_documentation_wavedec_example():
    x, name, level, want = PYWT_DOC_WAVEDEC
    got = pw.wavedec(np.asarray(x, dtype=float), name, level=level)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.allclose(g, w, atol=5e-9)
    assert np.allclose(pw.waverec(got, name), x, atol=1e-12)


def test_pywt_documentation_max_level_and_db2_taps():
    # "pywt.dwt_max_level(data_len=1000, filter_len=w.dec_len)" with w = Wavelet('sym5') prints 6
    assert pw.dwt_max_level(1000, 10) == 6
    # Wavelet('db2').dec_lo as printed in the documentation's filter-bank example
    assert np.allclose(pw.Wavelet("db2").dec_lo, [-0.12940952255092145, 0.22414386804185735, 0.836516303737469,
                                                  0.48296291314469025], atol=1e-9)

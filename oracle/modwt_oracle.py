"""float64 NumPy restatement of the reference's own MODWT arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED: checked against the
reference's ``src/modwt.py`` functions themselves (run in the build container by
tests/golden/make_golden.py; outputs committed under tests/golden/).

Index forms of what the reference expresses
through ``scipy.ndimage.convolve1d(mode="wrap")``:

* analysis   (modwt.py:86-102,126-144):  w_j[t] = sum_l h~[l] v_{j-1}[(t - 2^{j-1} l) mod N]
* synthesis  (modwt.py:105-123,147-160): v_{j-1}[t] = sum_l h~[l] w_j[(t + 2^{j-1} l) mod N]
                                                   + sum_l g~[l] v_j[(t + 2^{j-1} l) mod N]
* MRA        (modwt.py:56-83,163-194):   D_j[t] = sum_l hj[l] w_j[(t + l) mod N] with the
  level-j equivalent filter hj periodised to length N.
"""

from __future__ import annotations

import numpy as np

from .pywt_oracle import Wavelet


def _taps(filters):
    w = filters if hasattr(filters, "dec_lo") else Wavelet(filters)
    h = np.asarray(w.dec_hi, dtype=float)
    g = np.asarray(w.dec_lo, dtype=float)
    return h, g


def _circ_gather(v, taps, stride, sign):
    """sum_l taps[l] * v[(t + sign*stride*l) mod N] for every t."""
    N = v.size
    t = np.arange(N)
    out = np.zeros(N)
    for l, c in enumerate(taps):
        out += c * v[np.mod(t + sign * stride * l, N)]
    return out


def modwt(x, filters, level):
    """[w_1, ..., w_J, v_J] stacked as (J+1, N)  (modwt.py:126-144)."""
    h, g = _taps(filters)
    h_t, g_t = h / np.sqrt(2), g / np.sqrt(2)
    v = np.asarray(x, dtype=float)
    rows = []
    for j in range(1, level + 1):
        stride = 2 ** (j - 1)
        w = _circ_gather(v, h_t, stride, -1)
        v = _circ_gather(v, g_t, stride, -1)
        rows.append(w)
    rows.append(v)
    return np.vstack(rows)


def imodwt(w, filters):
    """Inverse pyramid (modwt.py:147-160)."""
    h, g = _taps(filters)
    h_t, g_t = h / np.sqrt(2), g / np.sqrt(2)
    w = np.asarray(w, dtype=float)
    level = w.shape[0] - 1
    v = w[-1]
    for j in range(level, 0, -1):
        stride = 2 ** (j - 1)
        v = _circ_gather(w[j - 1], h_t, stride, +1) + _circ_gather(v, g_t, stride, +1)
    return v


def _upsample(taps, j):
    """upArrow_op (modwt.py:56-63): insert 2^(j-1)-1 zeros between taps; j==0 -> [1]."""
    if j == 0:
        return np.array([1.0])
    step = 2 ** (j - 1)
    out = np.zeros(step * (len(taps) - 1) + 1)
    out[::step] = taps
    return out


def _periodise(f, N):
    """period_list (modwt.py:66-78): zero-pad to the next multiple of N (a whole
    extra N when already a multiple) and fold onto length N."""
    f = np.asarray(f, dtype=float)
    n_app = N - (f.size % N)
    f = np.concatenate([f, np.zeros(n_app)])
    if f.size < 2 * N:
        return f
    return f.reshape(-1, N).sum(axis=0)


def mra_filters(filters, level, N):
    """Periodised equivalent filters [h_1..h_J, g_J] used by modwtmra."""
    h, g = _taps(filters)
    out = []
    g_part = np.array([1.0])
    for j in range(level):
        g_part = np.convolve(g_part, _upsample(g, j))
        h_j = np.convolve(g_part, _upsample(h, j + 1)) / (2 ** ((j + 1) / 2.0))
        if j == 0:
            h_j = h / np.sqrt(2)
        out.append(_periodise(h_j, N))
    j = level - 1
    g_j = np.convolve(g_part, _upsample(g, j + 1)) / (2 ** ((j + 1) / 2.0))
    out.append(_periodise(g_j, N))
    return out


def modwtmra(w, filters):
    """Details D_1..D_J and smooth S_J stacked as (J+1, N)  (modwt.py:163-194)."""
    w = np.asarray(w, dtype=float)
    level, N = w.shape[0] - 1, w.shape[1]
    filt = mra_filters(filters, level, N)
    rows = [_circ_gather(w[j], filt[j], 1, +1) for j in range(level)]
    rows.append(_circ_gather(w[-1], filt[-1], 1, +1))
    return np.vstack(rows)


def smooth_signal(coeffs, filters, levels):
    """modwt.py:232-251: for l=J..1 zero rows 0..l-1 then imodwt."""
    out = {}
    for l in range(levels, 0, -1):
        c = np.array(coeffs, dtype=float, copy=True)
        c[:l] = 0.0
        out[l] = {"coeffs": c, "signal": imodwt(c, filters)}
    return out

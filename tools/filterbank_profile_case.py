"""One MODWT -> MRA -> wavedec -> waverec pass of a given shape for profiling (ncu captures).

    python tools/filterbank_profile_case.py [n] [batch] [f64|f32] [repeats]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wavelet_transformer_b200 import _shim  # noqa: E402
from wavelet_transformer_b200 import pywt_compat as pywt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1333
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
f64 = (sys.argv[3] if len(sys.argv) > 3 else "f64") == "f64"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
_shim.init(0)
dev = torch.device("cuda", 0)
dtype = torch.float64 if f64 else torch.float32
w = pywt.Wavelet("sym4")
J = 6
st = torch.cuda.current_stream().cuda_stream
x = torch.randn((B, n), dtype=dtype, device=dev)
out = torch.empty((B, J + 1, n), dtype=dtype, device=dev)
mra = torch.empty_like(out)
level = pywt.dwt_max_level(n, 8)
lens = _shim.dwt_coeff_lens(n, 8, level)
packed = torch.empty((B, int(lens.sum())), dtype=dtype, device=dev)
xr = torch.empty((B, _shim.waverec_len(lens, 8)), dtype=dtype, device=dev)
for r in range(reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    _shim.modwt_device(x.data_ptr(), B, n, w.dec_lo, w.dec_hi, J, out.data_ptr(), f64=f64, stream=st)
    ev[1].record()
    _shim.modwtmra_taps_device(out.data_ptr(), B, n, w.dec_lo, w.dec_hi, J, mra.data_ptr(), f64=f64, stream=st)
    ev[2].record()
    _shim.wavedec_device(x.data_ptr(), B, n, w.dec_lo, w.dec_hi, level, packed.data_ptr(), f64=f64, stream=st)
    ev[3].record()
    _shim.waverec_device(packed.data_ptr(), B, lens, w.rec_lo, w.rec_hi, xr.data_ptr(), f64=f64, stream=st)
    ev[4].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    print(f"rep {r}: n={n} B={B} {'f64' if f64 else 'f32'}: modwt {t[0]:.3f} ms, mra {t[1]:.3f} ms, wavedec {t[2]:.3f} ms, waverec {t[3]:.3f} ms")

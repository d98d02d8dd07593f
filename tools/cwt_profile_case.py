"""One fused CWT+power call of a given shape for profiling (ncu launch list / --set full captures).

    python tools/cwt_profile_case.py [n0] [batch] [repeats] [dj] [J]

Device-resident input and output; prints the device time per call.
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from wavelet_transformer_b200 import _shim  # noqa: E402

DT = 1 / 12
n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1346
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2960
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dj = float(sys.argv[4]) if len(sys.argv) > 4 else 1 / 12
J = int(sys.argv[5]) if len(sys.argv) > 5 else 84
_shim.init(0)
dev = torch.device("cuda", 0)
x = torch.randn((batch, n0), dtype=torch.float32, device=dev)
out = torch.empty((batch, J + 1, n0), dtype=torch.float32, device=dev)
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _shim.cwt_power_device(x.data_ptr(), batch, n0, DT, dj, 2 * DT, J, 6.0, out.data_ptr(), f64=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: {batch} x {n0} x {J + 1} in {ms:.3f} ms -> {batch * (J + 1) * n0 / ms * 1e3:.3e} coeff/s")

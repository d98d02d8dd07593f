#!/usr/bin/env python
"""Write-only, read-only and copy HBM bandwidth on this GPU (torch fill_/zero_/sum/copy_ over 8 GiB).

The fused CWT+power kernel is a pure store stream (1 byte read per 120 written), so next to the
copy figure of MEASURED_PEAKS.json (read + write in flight together) the write-only figure is the
bandwidth that actually bounds it.
"""

from __future__ import annotations

import json

import torch


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    dev = torch.device("cuda", 0)
    n = 2 << 30                                  # 2 Gi float32 = 8 GiB
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)
    by = 4.0 * n
    out = {
        "fill_GBs": by / timed(lambda: a.fill_(1.5)) / 1e9,
        "memset_GBs": by / timed(lambda: a.zero_()) / 1e9,
        "read_sum_GBs": by / timed(lambda: a.sum()) / 1e9,
        "copy_GBs": 2 * by / timed(lambda: b.copy_(a)) / 1e9,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()

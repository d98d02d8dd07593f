"""The box's bare device-to-host ceiling for the CWT end-to-end arm.

`bench.py`'s cfg4 `e2e` streams a 4 GB power plane per step from each GPU into pinned host memory;
this tool measures what the same copies cost with nothing else in the way: every rank issues one
`cudaMemcpyAsync` D2H of the same size from its GPU into its own pinned buffer, all ranks at once,
and the aggregate rate is reported.  `e2e` divided by this number is the fraction of the ceiling
the library reaches; what is left is the H2D copy, the kernel and the per-chunk synchronisation.

    python tools/bench_d2h_ceiling.py                     # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_d2h_ceiling.py

Rank 0 prints one JSON line and merges {"<world>": GB/s} into profiles/r2_d2h_ceiling.json
(or the file named by --out).
"""

from __future__ import annotations

import argparse
import json
import os
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=4 * 8192 * 120 * 1024, help="bytes per copy (bench.py e2e default)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r2_d2h_ceiling.json"))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    try:   # same placement as bench.py: run on the CPUs next to the GPU before pinning memory
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.bytes // 4
    d = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    h = torch.empty(n, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    h.copy_(d, non_blocking=True)     # warm-up: first touch of the pinned pages
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.copy_(d, non_blocking=True)     # one cudaMemcpyAsync per step
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gbs = world * args.bytes * args.steps / float(t.item()) / 1e9
    if rank == 0:
        rec = {"world": world, "bytes_per_copy": args.bytes, "steps": args.steps, "aggregate_d2h_GBs": gbs,
               "per_gpu_GBs": gbs / world}
        print(json.dumps(rec))
        out = Path(args.out)
        cur = json.loads(out.read_text()) if out.exists() else {}
        cur[str(world)] = gbs
        out.parent.mkdir(parents=True, exist_ok=True)
        out.write_text(json.dumps(cur, indent=1) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

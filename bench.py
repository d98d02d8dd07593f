#!/usr/bin/env python
"""Benchmark of the wavelet hot path (BASELINE.json metric: "CWT coeffs/sec; WCT Monte Carlo
surrogates/sec at 1/2/4/8 B200 vs CPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wct_mc|cwt] [--impl reference]

Workloads (one "step" = one pass of the hot path over one batch of synthetic input):
  wct_mc  (default, the line's `value`)  BASELINE cfg5 as stated: ONE job of 100 000 AR(1)
          surrogate pairs (a1 = 0.989, a2 = 0.966, seed 2024, N = 3351 -> FFT 4096, 66 scales,
          dj = 1/8), FP32, STRONG-scaled: the N ranks split the same 100 000 realisations by global
          index.  A step runs from the first launch to sig95: Philox surrogates -> CWT -> smoothing
          -> coherence -> per-scale histograms -> the one collective (integer all-reduce of the
          [66, 1000] histograms, NCCL) -> percentile kernel -> sig95 on rank 0's host.
          metric: surrogate pairs / s.  The bench itself asserts that the all-reduced histogram is
          bit-identical to the one a single GPU computes for the whole job.
  cwt     (reported under `secondary` with its own roofline / e2e / cpu_baseline)  BASELINE cfg4:
          fused Morlet CWT + |W|^2 of AR(1) series, N = 1024, 120 scales (dj = 1/12, s0 = 2dt),
          FP32.  cfg4 is 1M series over 8 GPUs = 125 000 series per GPU; that per-GPU shard
          (61.4 GB of coefficients, device-resident) is the batch at every N (weak scaling).
          metric: CWT coefficients / s.

Under torchrun every rank drives one GPU; timing is CUDA events on the launching stream,
bracketed by barrier + synchronize, MAX over ranks; rank 0 prints one JSON line.
`e2e` of wct_mc is the job through the public API with HOST buffers: the surrogate pairs come from
pinned host memory (north_star's per-realisation parity mode; each rank its block), 2.68 GB of H2D
per job, double-buffered against the kernels inside the library, histogram back to the host,
all-reduce, percentile; it must give the histogram of the device-RNG arm (asserted).
`e2e_one_call` is the call a user of the reference makes -- pycwt's wct_significance, here
`wtb_wct_significance`: parameters in, thresholds out, no input payload -- in ONE process that
drives all N GPUs through the library's own worker pool (rank 0 makes the call, the other ranks
wait on a CPU barrier), host wall clock.
`--impl reference` times the CPU restatement of the reference's path (oracle/, NumPy float64 --
pycwt itself is not installable offline) on all host cores instead.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DT = 1 / 12
CWT = dict(n0=1024, dj=1 / 12, s0=2 * DT, J=119, f0=6.0, ar1=0.7)
MC = dict(a1=0.989, a2=0.966, dj=1 / 8, s0=2 * DT, J=65, f0=6.0, seed=2024, level=0.95)
MC_FLOP = 150e6          # algorithmic flop per realisation (N = 4096, 66 scales, 5 N log2 N per FFT)
FP32_PEAK_NOMINAL = 148 * 128 * 2 * 1.965e9


def _profile_json(name):
    p = ROOT / "profiles" / name
    return json.loads(p.read_text()) if p.exists() else None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 100 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,utilization.gpu,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def wait_ready(self, timeout_s: float = 5.0):
        """Block until the first sample is on disk: nvidia-smi's start-up (NVML init over every
        GPU of the box, 0.1-1 s) holds driver locks that stall kernel launches of a launch-heavy
        step for tens of ms -- keep that out of the timed region."""
        if self.proc is None:
            return
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < timeout_s and os.path.getsize(self.file.name) == 0:
            time.sleep(0.02)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        rows = []
        for line in Path(self.file.name).read_text().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8 and parts[0].isdigit():
                rows.append(parts)
        os.unlink(self.file.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        loaded = [r for r in rows if r[2].isdigit() and int(r[2]) >= 50] or rows
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in loaded)]
        power = [float(r[3]) for r in loaded if r[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": statistics.median(int(r[0]) for r in loaded), "sm_max_mhz": int(rows[0][1]),
                "reasons": reasons, "samples": len(rows), "samples_under_load": len(loaded),
                "power_w_max": max(power) if power else None}


# --------------------------------------------------------------------------- CPU arms
def _cpu_cwt_chunk(args):
    """Oracle CWT+power of `count` AR(1) series; returns coefficients produced."""
    seed, count = args
    from oracle import pycwt_oracle as po
    rng = np.random.default_rng(seed)
    mother = po.Morlet(CWT["f0"])
    done = 0
    for _ in range(count):
        x = po.rednoise(CWT["n0"], CWT["ar1"], 1.0, rng)
        W = po.cwt(x, DT, CWT["dj"], CWT["s0"], CWT["J"], mother)[0]
        power = np.abs(W) ** 2
        done += power.size
    return done


def _cpu_mc_chunk(args):
    seed, count = args
    from oracle import pycwt_oracle as po
    po.wct_significance(MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], mc_count=count,
                        rng=np.random.default_rng(seed), faithful_loop=True)
    return count


def cpu_baseline(workload: str, budget_s: float = 12.0):
    """Single-core oracle ('port') on a bounded sample of the same workload."""
    fn, unit_per = (_cpu_cwt_chunk, 16) if workload == "cwt" else (_cpu_mc_chunk, 2)
    fn((0, 1))  # warm caches / imports
    t0 = time.perf_counter()
    units = 0.0
    n = 0
    while time.perf_counter() - t0 < budget_s:
        units += fn((1000 + n, unit_per))
        n += unit_per
    dt = time.perf_counter() - t0
    if workload == "cwt":
        return {"value": units / dt, "unit": "coeff/s", "cores": 1, "kind": "port", "host_cpus": os.cpu_count(),
                "sample": f"{n} series x N=1024 x 120 scales, oracle/pycwt_oracle.cwt + |W|^2, float64, {dt:.1f} s"}
    return {"value": units / dt, "unit": "surrogates/s", "cores": 1, "kind": "port", "host_cpus": os.cpu_count(),
            "sample": f"{n} realisations of the 100000 (N=3351->4096, 66 scales), oracle wct_significance with "
                      f"pycwt's per-sample Python histogram loop, float64, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the CPU restatement on all host cores (pycwt/pywt are not installable
    offline, so the oracle port stands in for them)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    fn, per_core = (_cpu_cwt_chunk, 24) if args.workload == "cwt" else (_cpu_mc_chunk, 1)
    with mp.get_context("fork").Pool(cores) as pool:
        def step(i):
            return sum(pool.map(fn, [(10_000 * i + c, per_core) for c in range(cores)]))
        for w in range(args.warmup):
            step(w)
        t0 = time.perf_counter()
        units = 0
        for k in range(args.steps):
            units += step(100 + k)
        dt = time.perf_counter() - t0
    value = units / dt
    metric, unit = (("cwt_coeffs_per_sec", "coeff/s") if args.workload == "cwt"
                    else ("wct_mc_surrogates_per_sec", "surrogates/s"))
    sample = (f"{cores * per_core} series per step (N=1024, 120 scales)" if args.workload == "cwt"
              else f"{cores * per_core} of the job's 100000 realisations per step (N=3351->4096, 66 scales, "
                   "pycwt's per-sample Python histogram loop); throughput metric, per-realisation cost does not "
                   "depend on the sample size")
    config = cwt_config(args) if args.workload == "cwt" else mc_config(args)
    config["reference_sample_per_step"] = cores * per_core
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong" if args.workload == "wct_mc" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": sample + ", oracle/pycwt_oracle (NumPy/SciPy float64) in a fork pool"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cwt_config(args):
    return {"workload": "cfg4: fused Morlet CWT+|W|^2, synthetic AR(1) g=0.7 series x N=1024, 120 scales "
                        "(dj=1/12, s0=2dt, J=119); 125000 series per GPU = the 8-GPU shard of 1M series",
            "series_per_gpu": args.series, "n": 1024, "scales": 120,
            "parallelism": f"series sharded over {args.gpus} GPU(s), no collective",
            "l2_policy": "inputs (512 MB) and outputs (61 GB) exceed the 126 MB L2"}


def mc_config(args):
    return {"workload": "cfg5: WCT Monte Carlo significance, ONE job of 100000 AR(1) surrogate pairs x N=3351 "
                        "(FFT 4096), 66 scales (dj=1/8, s0=2dt), a1=0.989, a2=0.966, seed 2024, level 0.95; "
                        "first launch -> sig95 on the host",
            "realisations_per_job": args.realisations, "n": 3351, "nfft": 4096, "scales": 66,
            "parallelism": f"the job's realisations split by global index over {args.gpus} GPU(s) (strong scaling); "
                           "one integer all-reduce (NCCL) of the [66,1000] histograms per step, then the "
                           "percentile kernel",
            "l2_policy": "per-step intermediates (> 10 GB) exceed the 126 MB L2"}


# --------------------------------------------------------------------------- GPU arm
class Ctx:
    pass


def setup_gpu(args):
    import torch
    import torch.distributed as dist

    from wavelet_transformer_b200 import _shim

    c = Ctx()
    # Only the JSON line may reach stdout: route everything libraries print (NCCL's version
    # banner, warnings) to stderr and keep a private handle on the real stdout.
    sys.stdout.flush()
    c.real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    if c.world != args.gpus and c.world > 1:
        args.gpus = c.world
    torch.cuda.set_device(c.local)
    # Run this rank on the CPUs next to its GPU before any pinned host buffer is allocated: the
    # host-buffer (e2e) arm is a PCIe stream per rank, and first-touch places its staging memory
    # on the NUMA node the thread runs on.
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(c.local)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        numa = sorted(os.sched_getaffinity(0))
        c.numa = f"{numa[0]}-{numa[-1]} ({len(numa)} cpus)"
    except Exception as exc:  # containers without the affinity interface: keep the default placement
        c.numa = f"unchanged ({type(exc).__name__})"
    _shim.init(c.local)
    c.cpu_group = None
    if c.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", c.local))
        c.cpu_group = dist.new_group(backend="gloo")     # host-side waits that must not occupy a GPU
    c.dev = torch.device("cuda", c.local)
    c.torch, c.dist, c.shim = torch, dist, _shim

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        if c.world > 1:
            dist.barrier(group=c.cpu_group)

    def max_over_ranks(v: float) -> float:
        if c.world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=c.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c.barrier, c.cpu_barrier, c.max_over_ranks = barrier, cpu_barrier, max_over_ranks
    c.hbm_peak, c.peak_src = measured_peaks()
    return c


def timed_steps(c, step, steps, warmup, sampler=None):
    """W warm-up steps, then exactly K steps between CUDA events on the current stream, bracketed by
    barrier + synchronize; returns (ms total = max over ranks, per-step spread, launches)."""
    torch = c.torch
    if sampler:
        sampler.wait_ready()
    for _ in range(warmup):
        step()
    c.barrier()
    launches0 = c.shim.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps - 1)]   # per-step spread, no syncs
    ev0.record()
    for i in range(steps):
        step()
        if i < steps - 1:
            marks[i].record()
    ev1.record()
    torch.cuda.synchronize()
    ms = c.max_over_ranks(ev0.elapsed_time(ev1))
    edges = [ev0] + marks + [ev1]
    per_step = sorted(a.elapsed_time(b) for a, b in zip(edges[:-1], edges[1:]))
    spread = {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]}
    launches = c.shim.kernel_launches() - launches0
    c.barrier()
    return ms, spread, int(launches)


def bench_mc(c, args, sampler):
    """cfg5: the 100 000-realisation significance job, strong-scaled over the ranks."""
    torch, dist, shim = c.torch, c.dist, c.shim
    from wavelet_transformer_b200 import engine
    R = args.realisations
    S = MC["J"] + 1
    first, stop = engine.shard_range(R, c.rank, c.world)
    nsurr, maxscale = shim.wct_mc_geometry(DT, MC["dj"], MC["s0"], MC["J"], MC["f0"])
    has = shim.row_has_points(DT, MC["dj"], MC["s0"], MC["J"], MC["f0"])
    hist = torch.zeros((S, shim.NBINS), dtype=torch.int64, device=c.dev)
    d_sig = torch.empty(S, dtype=torch.float64, device=c.dev)
    h_sig = torch.empty(S, dtype=torch.float64).pin_memory()
    red0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + args.warmup)]
    red1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + args.warmup)]
    k = {"i": 0}

    def step():
        stream = torch.cuda.current_stream().cuda_stream
        hist.zero_()
        engine.wct_hist_resident(hist, MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"],
                                 first, stop - first, MC["seed"])
        i = k["i"]
        k["i"] += 1
        red0[i].record()
        engine.reduce_histogram(hist)                      # the one collective of the path
        red1[i].record()
        shim.wct_sig_from_hist_device(hist.data_ptr(), S, maxscale, MC["level"], has, d_sig.data_ptr(), stream=stream)
        if c.rank == 0:
            h_sig.copy_(d_sig, non_blocking=True)          # sig95 lands on the host inside the timed region

    ms, spread, launches = timed_steps(c, step, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler else None
    value = R * args.steps / (ms * 1e-3)
    per_step_s = ms * 1e-3 / args.steps
    # the collective's share: event pairs around the all-reduce (includes waiting for the slowest rank)
    red_ms = sorted(a.elapsed_time(b) for a, b in zip(red0[args.warmup:], red1[args.warmup:]))
    allreduce = {"median_ms": red_ms[len(red_ms) // 2], "max_ms": red_ms[-1],
                 "share_of_step": red_ms[len(red_ms) // 2] / (1e3 * per_step_s),
                 "note": "CUDA events around the histogram all-reduce on rank 0; includes the wait for the "
                         "slowest rank" if c.world > 1 else "single rank: no collective issued"}

    # ---- checks inside the bench: percentile kernel == host percentile; partition invariance
    total = hist.cpu().numpy().astype(np.uint64)
    sig_host = shim.wct_sig_from_hist(total, maxscale, MC["level"], has)
    check = {"hist_sha256_16": hashlib.sha256(total.tobytes()).hexdigest()[:16], "hist_total": int(total.sum())}
    if c.rank == 0:
        assert np.array_equal(h_sig.numpy(), sig_host, equal_nan=True), "device percentile differs from the host's"
        whole = torch.zeros_like(hist)
        if c.world > 1:
            engine.wct_hist_resident(whole, MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"], 0, R, MC["seed"])
            how = f"all-reduced histogram of {c.world} ranks == the whole job on rank 0 alone"
        else:
            for a, b in (engine.shard_range(R, 0, 3), engine.shard_range(R, 1, 3), engine.shard_range(R, 2, 3)):
                engine.wct_hist_resident(whole, MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"], a, b - a,
                                         MC["seed"])
            how = "one launch of the whole job == the sum of three separately computed thirds"
        torch.cuda.synchronize()
        assert torch.equal(whole, hist), "histogram depends on the partition of realisations"
        check["partition_invariance"] = "bit-identical: " + how
        check["sig95_first_rows"] = [float(v) for v in sig_host[:4]]
    c.barrier()

    # ---- end to end: the reference-facing call, one process driving all N GPUs
    e2e = None
    steps_e = args.e2e_steps
    if c.rank == 0:
        n_pool = c.world if shim.device_count() >= c.world else 1
        shim.init_multi(n_pool)
        kw = dict(level=MC["level"], mc_count=R, seed=MC["seed"], f64=False, return_hist=True)
        sig_e, hist_e = shim.wct_significance(MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"], **kw)
        assert np.array_equal(hist_e, total) and np.array_equal(sig_e, sig_host, equal_nan=True), \
            "one-call multi-GPU path differs from the sharded ranks"
        t0 = time.perf_counter()
        for _ in range(steps_e):
            shim.wct_significance(MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"], **kw)
        e2e_s = time.perf_counter() - t0
        shim.init_multi(1)
        e2e = {"value": R * steps_e / e2e_s, "unit": "surrogates/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": 8 * S * shim.NBINS + 0, "realisations_per_step": R, "steps": steps_e,
               "gpus_driven_by_the_call": n_pool, "ms_per_step": 1e3 * e2e_s / steps_e,
               "path": "wtb_wct_significance (what pycwt_compat.wct_significance / run_wct call): ONE host call in "
                       "one process, the library's worker pool drives every GPU, device 0 sums the peers' "
                       "histograms over NVLink, histogram D2H + percentile on the host; host wall clock. Device "
                       "RNG, so the only payload is the 528 KB histogram coming back",
               "same_numbers_as_value_arm": True}
    c.cpu_barrier()

    # ---- end to end with a real input payload: host-supplied surrogates (north_star's per-realisation
    # parity mode).  Every rank holds its block of the job's surrogate pairs in pinned host memory;
    # a step = H2D of the block (double-buffered against the kernels inside the library) + pipeline +
    # histogram D2H + the all-reduce + percentile on rank 0.
    e2e_inj = None
    if not args.no_injected:
        nloc = stop - first
        sur_h = torch.empty((nloc, 2, nsurr), dtype=torch.float32).pin_memory()
        blk = 4096
        tmp = torch.empty((blk, 2, nsurr), dtype=torch.float32, device=c.dev)
        for b0 in range(0, nloc, blk):          # the same series the device RNG draws: same histogram expected
            nb = min(blk, nloc - b0)
            shim.rednoise_device(MC["a1"], MC["a2"], nsurr, first + b0, nb, MC["seed"], tmp.data_ptr(),
                                 stream=torch.cuda.current_stream().cuda_stream)
            sur_h[b0:b0 + nb].copy_(tmp[:nb])
        torch.cuda.synchronize()
        del tmp
        lib = shim.lib()
        import ctypes as C
        hist_pin = torch.zeros((S, shim.NBINS), dtype=torch.int64).pin_memory()     # the call's host histogram
        hist_h = hist_pin.numpy().view(np.uint64)
        t_dev = torch.zeros((S, shim.NBINS), dtype=torch.int64, device=c.dev)

        def inj_step():
            hist_h[:] = 0
            rc = lib.wtb_wct_mc_hist(MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"], first, nloc,
                                     C.c_uint64(MC["seed"]), C.c_void_p(sur_h.data_ptr()), 0,
                                     C.c_void_p(hist_pin.data_ptr()), None)
            if rc != 0:
                raise RuntimeError(lib.wtb_last_error().decode())
            if c.world > 1:                       # the one collective: histograms meet on the devices
                t_dev.copy_(hist_pin, non_blocking=True)
                engine.reduce_histogram(t_dev)
                hist_pin.copy_(t_dev, non_blocking=True)
                torch.cuda.synchronize()
            return shim.wct_sig_from_hist(hist_h, maxscale, MC["level"], has), hist_h
        sig_i, t_i = inj_step()
        if c.rank == 0:
            assert np.array_equal(t_i, total), "injected surrogates give another histogram"
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            inj_step()
        torch.cuda.synchronize()
        inj_s = c.max_over_ranks(time.perf_counter() - t0)
        e2e_inj = {"value": R * args.e2e_steps / inj_s, "unit": "surrogates/s", "ms_per_step": 1e3 * inj_s / args.e2e_steps,
                   "h2d_bytes_per_step": 4 * 2 * nsurr * R, "h2d_bytes_per_step_per_rank": 4 * 2 * nsurr * nloc,
                   "d2h_bytes_per_step": 8 * S * shim.NBINS * c.world, "steps": args.e2e_steps,
                   "path": "wtb_wct_mc_hist with HOST-supplied surrogates in pinned memory (each rank its block), "
                           "histogram back to the host, all-reduce, wtb_wct_sig_from_hist; H2D, kernels and D2H inside "
                           "the timed region",
                   "same_histogram_as_value_arm": True}
        del sur_h
    c.barrier()

    ach = MC_FLOP * R / c.world / per_step_s / 1e12      # per GPU: every rank runs R / world realisations
    traffic = _profile_json("r2_traffic_mc.json")
    roofline = {"bound": "fp32", "achieved": ach, "peak": FP32_PEAK_NOMINAL / 1e12, "unit": "TFLOP/s",
                "frac": ach / (FP32_PEAK_NOMINAL / 1e12),
                "traffic": ((traffic or {}).get("dram_bytes_per_realisation_pipeline") or 0) * R / c.world or None,
                "traffic_source": "profiles/r2_traffic_mc.json: dram read+write bytes per realisation of the four "
                                  "pipeline kernels (ncu --set full), times the realisations one GPU runs per step; "
                                  "HBM is not the bound here (3.4 MB per realisation = 1.1 TB/s at this rate)",
                "peak_source": "nominal non-tensor FP32 peak (148 SM x 128 lanes x 2 x 1.965 GHz); this path is "
                               "FP32-FLOP bound, tensor cores are not applicable (no dense contraction); "
                               "MEASURED_PEAKS.json has no FP32 entry",
                "kernel": "the Monte-Carlo pipeline per GPU: k_wct_spec_4096 46 %, k_wct_coh_4096 36 %, "
                          "k_wct_boxcar_4096 11 %, k_wct_spec_direct 4 %, k_rednoise 1.5 %, k_fwd_fft_4096 0.8 % "
                          "(profiles/r2_launches_mc.csv)",
                "algorithmic_flop_per_realisation": MC_FLOP, "per_gpu": True,
                "frac_at_measured_clock": (ach * 1e12 / (148 * 128 * 2 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6))}
    line = {
        "metric": "wct_mc_surrogates_per_sec", "value": value, "unit": "surrogates/s", "n_gpus": c.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": mc_config(args),
        "roofline": roofline, "e2e": e2e_inj if e2e_inj else e2e, "gpu_launches": launches, "clocks": clocks,
        "ms_per_step_spread": spread, "allreduce": allreduce, "checks": check,
    }
    if e2e_inj:
        line["e2e_one_call"] = e2e      # the reference-facing call (parameters in, thresholds out: no input payload)
    return line


def bench_cwt(c, args, sampler=None, steps=None, warmup=None):
    """cfg4: fused CWT+power on the per-GPU shard (weak scaling, no collective)."""
    torch, shim = c.torch, c.shim
    from wavelet_transformer_b200 import engine
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    S, n0, B = CWT["J"] + 1, CWT["n0"], args.series

    def make_series(count, seed):
        """Unit-variance AR(1) g=0.7 series generated on the device by the library's own Philox
        generator (synthetic cfg4 input; counter = global series index, so shards differ)."""
        pairs = (count + 1) // 2
        buf = torch.empty((pairs, 2, n0), dtype=torch.float32, device=c.dev)
        shim.rednoise_device(CWT["ar1"], CWT["ar1"], n0, c.rank * pairs, pairs, seed, buf.data_ptr(),
                             stream=torch.cuda.current_stream().cuda_stream)
        y = buf.reshape(pairs * 2, n0)[:count] * (1 - CWT["ar1"] ** 2) ** 0.5
        return y.contiguous()

    x = make_series(B, 1234)
    power = torch.empty((B, S, n0), dtype=torch.float32, device=c.dev)

    def step():
        engine.cwt_power_resident(x, power, DT, CWT["dj"], CWT["s0"], CWT["J"], CWT["f0"])

    ms, spread, launches = timed_steps(c, step, steps, warmup, sampler)
    clocks = sampler.stop() if sampler else None
    value = c.world * B * S * n0 * steps / (ms * 1e-3)
    per_launch_s = ms * 1e-3 / steps
    alg_bytes = 4.0 * B * n0 * (1 + S)          # read x once, write the power plane once
    alg_flops = B * (5 * n0 * 10 + S * (5 * n0 * 10 + 5 * n0))

    # ---- end to end through the C ABI with pinned HOST buffers
    Be = args.e2e_series
    xh = torch.empty((Be, n0), dtype=torch.float32).pin_memory()
    xh.copy_(x[:Be].cpu() if Be <= x.shape[0] else make_series(Be, 99).cpu())
    ph = torch.empty((Be, S, n0), dtype=torch.float32).pin_memory()
    lib = shim.lib()

    def e2e_step():
        rc = lib.wtb_cwt_morlet(xh.data_ptr(), Be, n0, n0, DT, CWT["dj"], CWT["s0"], CWT["J"], CWT["f0"], 0,
                                ph.data_ptr(), None, None)
        if rc != 0:
            raise RuntimeError(lib.wtb_last_error().decode())
    e2e_step()
    c.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = c.max_over_ranks(time.perf_counter() - t0)
    ceiling = _profile_json("r2_d2h_ceiling.json") or {}
    e2e = {"value": c.world * Be * S * n0 * args.e2e_steps / e2e_s, "unit": "coeff/s",
           "h2d_bytes_per_step": 4 * Be * n0, "d2h_bytes_per_step": 4 * Be * S * n0,
           "series_per_step": Be, "steps": args.e2e_steps, "rank0_cpu_affinity": c.numa,
           "d2h_GBs_aggregate": c.world * 4 * Be * S * n0 * args.e2e_steps / e2e_s / 1e9,
           "bare_d2h_ceiling_GBs": ceiling.get(str(c.world)),
           "path": "wtb_cwt_morlet with pinned HOST buffers; H2D, kernels and D2H inside the timed region"}
    del xh, ph

    achieved = alg_bytes / per_launch_s / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    traffic = _profile_json("r1_traffic.json")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s",
                "frac": achieved / c.hbm_peak,
                "traffic": (args.traffic_bytes if args.traffic_bytes is not None else
                            (float(traffic["dram_bytes_per_series"]) * B if traffic else None)),
                "traffic_source": "profiles/r1_traffic.json: dram read+write bytes per series from one "
                                  "ncu --set full capture, scaled to this launch's series count",
                "peak_source": c.peak_src, "kernel": "k_cwt_fast_1024: fused CWT+power (one launch per step)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "fp32": {"algorithmic_flop_per_launch": alg_flops,
                         "achieved_tflops": alg_flops / per_launch_s / 1e12,
                         "peak_tflops_nominal": FP32_PEAK_NOMINAL / 1e12,
                         "frac_nominal": alg_flops / per_launch_s / FP32_PEAK_NOMINAL,
                         "frac_at_measured_clock": alg_flops / per_launch_s / (148 * 128 * 2 * sm_mhz * 1e6),
                         "note": "5*N*log2N per FFT convention (SURVEY 8d): 6.81 MFLOP per series; "
                                 "non-tensor FP32 peak = 148 SM x 128 lanes x 2 x clock"}}
    line = {
        "metric": "cwt_coeffs_per_sec", "value": value, "unit": "coeff/s", "n_gpus": c.world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cwt_config(args),
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "ms_per_step_spread": spread,
    }
    del x, power
    torch.cuda.empty_cache()
    return line


def bench_filterbank(c):
    """MODWT LA8 J=6 (BASELINE cfg2 shape, batched): HBM-bound filterbank kernel, rank-local."""
    torch, shim = c.torch, c.shim
    from wavelet_transformer_b200 import pywt_compat as pywt
    la8 = pywt.Wavelet("sym4")
    Bf, nf, Jf = 50_000, 1024, 6
    xf = torch.randn((Bf, nf), dtype=torch.float64, device=c.dev)
    wf = torch.empty((Bf, Jf + 1, nf), dtype=torch.float64, device=c.dev)
    cur = torch.cuda.current_stream().cuda_stream

    def fb_step():
        shim.modwt_device(xf.data_ptr(), Bf, nf, la8.dec_lo, la8.dec_hi, Jf, wf.data_ptr(), f64=True, stream=cur)
    for _ in range(3):
        fb_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fb_step()
    e1.record()
    torch.cuda.synchronize()
    fb_s = e0.elapsed_time(e1) * 1e-3 / 5
    fb_bytes = 8.0 * nf * (Jf + 2) * Bf
    return {"metric": "modwt_coeffs_per_sec", "value": Bf * (Jf + 1) * nf / fb_s, "unit": "coeff/s",
            "per_gpu": True, "achieved_GBs": fb_bytes / fb_s / 1e9, "frac_hbm": fb_bytes / fb_s / 1e9 / c.hbm_peak,
            "config": f"MODWT LA8 (sym4) J={Jf}, {Bf} series x N={nf}, FP64, device-resident; "
                      "algorithmic bytes 8 N (1 + J+1) per series"}


def bench_cwt_shapes(c):
    """Fused CWT+power at the other FFT lengths (rank-local, device-resident): BASELINE cfg1's shape in batch
    (1346 samples -> nfft 2048, 85 scales: two warps per row) and the Monte-Carlo surrogate length as a plain
    CWT (3351 -> 4096, 66 scales: four warps per row).  Roofline = the slower of FFT flops at the nominal FP32
    peak and coefficient bytes at the measured HBM bandwidth (north_star); SURVEY 8d's per-series figures."""
    torch, shim = c.torch, c.shim
    fp32_peak = 148 * 128 * 2 * 1.965e9
    out = []
    for label, n0, dj, J, batch in (("cfg1 shape", 1346, 1 / 12, 84, 10_000), ("nfft 4096", 3351, 1 / 8, 65, 4_000)):
        S = J + 1
        x = torch.randn((batch, n0), dtype=torch.float32, device=c.dev)
        pw = torch.empty((batch, S, n0), dtype=torch.float32, device=c.dev)

        def step():
            shim.cwt_power_device(x.data_ptr(), batch, n0, DT, dj, 2 * DT, J, 6.0, pw.data_ptr(), f64=False)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / 5
        nfft = 1 << (n0 - 1).bit_length()
        lg = nfft.bit_length() - 1
        flop = 5.0 * nfft * lg + S * (5.0 * nfft * lg + 5.0 * nfft)
        by = 4.0 * n0 * (1 + S)
        t_hbm, t_fp32 = by / (c.hbm_peak * 1e9), flop / fp32_peak
        out.append({"shape": label, "n0": n0, "nfft": nfft, "scales": S, "series": batch, "ms": t * 1e3,
                    "value": batch * S * n0 / t, "unit": "coeff/s", "per_gpu": True,
                    "achieved_GBs": by * batch / t / 1e9, "achieved_TFLOPs": flop * batch / t / 1e12,
                    "bound": "fp32" if t_fp32 > t_hbm else "hbm", "frac_roofline": max(t_hbm, t_fp32) * batch / t})
        del x, pw
    return out


def run_gpu(args):
    c = setup_gpu(args)
    sampler = ClockSampler(c.local) if c.rank == 0 else None
    def guarded(name, fn):
        """Secondary measurements must never cost the primary line: a failure is recorded, not raised
        (every rank takes the same path -- the secondaries contain collectives only through barriers)."""
        try:
            return fn()
        except Exception as exc:  # noqa: BLE001
            return {"error": f"{name}: {type(exc).__name__}: {exc}"}

    if args.workload == "wct_mc":
        line = bench_mc(c, args, sampler)
        if not args.no_secondary:
            sampler2 = ClockSampler(c.local) if c.rank == 0 else None
            line["secondary"] = guarded("cfg4 secondary", lambda: bench_cwt(c, args, sampler2, steps=args.secondary_steps,
                                                                           warmup=3))
            line["secondary_filterbank"] = guarded("filterbank secondary", lambda: bench_filterbank(c))
            line["secondary_cwt_shapes"] = guarded("cwt shapes secondary", lambda: bench_cwt_shapes(c))
    else:
        line = bench_cwt(c, args, sampler)
        if not args.no_secondary:
            line["secondary_filterbank"] = guarded("filterbank secondary", lambda: bench_filterbank(c))
    if c.rank == 0:
        if c.world == 1 and not args.no_cpu:
            line["cpu_baseline"] = guarded("cpu baseline", lambda: cpu_baseline(args.workload, args.cpu_budget))
            if "secondary" in line and "error" not in line["secondary"]:
                line["secondary"]["cpu_baseline"] = guarded("cpu baseline (cwt)", lambda: cpu_baseline("cwt", args.cpu_budget))
        c.real_stdout.write(json.dumps(line) + "\n")
        c.real_stdout.flush()
    if c.world > 1:
        c.cpu_barrier()
        c.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="wct_mc", choices=["wct_mc", "cwt"])
    ap.add_argument("--realisations", type=int, default=100_000, help="realisations of the one cfg5 job (all GPUs)")
    ap.add_argument("--series", type=int, default=125_000, help="series per GPU per step (cfg4 shard)")
    ap.add_argument("--secondary-steps", type=int, default=20)
    ap.add_argument("--e2e-series", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-injected", action="store_true", help="skip the host-supplied-surrogates end-to-end arm")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

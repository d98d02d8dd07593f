#!/usr/bin/env python
"""Benchmark of the wavelet hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cwt|wct_mc] [--impl reference]

Workloads (one "step" = one pass of the hot path over one batch of synthetic input):
  cwt     BASELINE cfg4: fused Morlet CWT + |W|^2 of AR(1) series, N=1024, 120 scales
          (dj=1/12, s0=2dt, J=119), FP32.  cfg4 is 1M series over 8 GPUs = 125 000 series
          per GPU; that per-GPU shard (61.4 GB of coefficients, device-resident) is the
          batch at every N (weak scaling).  metric: CWT coefficients / s.
  wct_mc  BASELINE cfg5: AR(1) Monte Carlo coherence significance, surrogates of
          N=3351 -> FFT 4096, 66 scales (dj=1/8), FP32.  metric: surrogate pairs / s.
          Ranks shard realisations; one integer all-reduce of the histograms per step.

Under torchrun every rank drives one GPU; timing is CUDA events on the launching
stream, bracketed by barrier + synchronize, MAX over ranks; rank 0 prints one JSON line.
`--impl reference` times the CPU restatement of the reference's path (oracle/, NumPy
float64 -- pycwt itself is not installable offline) on all host cores instead.
"""

from __future__ import annotations

import argparse
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
MC = dict(a1=0.989, a2=0.966, dj=1 / 8, s0=2 * DT, J=65, f0=6.0, seed=2024)


def measured_traffic_per_series():
    """DRAM bytes per series of the dominant kernel from the committed ncu capture (or None)."""
    p = ROOT / "profiles" / "r1_traffic.json"
    if p.exists():
        return float(json.loads(p.read_text())["dram_bytes_per_series"])
    return None


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
        return {"value": units / dt, "unit": "coeff/s", "cores": 1, "kind": "port",
                "sample": f"{n} series x N=1024 x 120 scales, oracle/pycwt_oracle.cwt + |W|^2, float64, {dt:.1f} s"}
    return {"value": units / dt, "unit": "surrogates/s", "cores": 1, "kind": "port",
            "sample": f"{n} realisations (N=3351->4096, 66 scales), oracle wct_significance with pycwt's "
                      f"per-sample Python histogram loop, float64, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the CPU restatement on all host cores (pycwt/pywt are not installable
    offline, so the oracle port stands in for them)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    fn, per_core = (_cpu_cwt_chunk, 24) if args.workload == "cwt" else (_cpu_mc_chunk, 2)
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
              else f"{cores * per_core} realisations per step (N=3351->4096, 66 scales, Python histogram loop)")
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cores * per_core),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": sample + ", oracle/pycwt_oracle (NumPy/SciPy float64) in a fork pool"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, per_step_units=None):
    if args.workload == "cwt":
        return {"workload": "cfg4: fused Morlet CWT+|W|^2, synthetic AR(1) g=0.7 series x N=1024, 120 scales "
                            "(dj=1/12, s0=2dt, J=119); 125000 series per GPU = the 8-GPU shard of 1M series",
                "series_per_gpu": per_step_units if per_step_units else args.series, "n": 1024, "scales": 120,
                "parallelism": f"series sharded over {args.gpus} GPU(s), no collective",
                "l2_policy": "inputs (512 MB) and outputs (61 GB) exceed the 126 MB L2"}
    return {"workload": "cfg5: AR(1) Monte Carlo WCT significance, surrogate pairs x N=3351 (FFT 4096), "
                        "66 scales (dj=1/8, s0=2dt), a1=0.989, a2=0.966, seed 2024",
            "realisations_per_gpu": per_step_units if per_step_units else args.realisations, "n": 3351,
            "nfft": 4096, "scales": 66,
            "parallelism": f"realisations sharded over {args.gpus} GPU(s); one int64 histogram all-reduce per step",
            "l2_policy": "per-step intermediates (>1 GB) exceed the 126 MB L2"}


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from wavelet_transformer_b200 import _shim, engine

    # Only the JSON line may reach stdout: route everything libraries print (NCCL's version
    # banner, warnings) to stderr and keep a private handle on the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    # Run this rank on the CPUs next to its GPU before any pinned host buffer is allocated: the
    # host-buffer (e2e) arm is a PCIe stream per rank, and first-touch places its staging memory
    # on the NUMA node the thread runs on.
    numa = None
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        numa = sorted(os.sched_getaffinity(0))
        numa = f"{numa[0]}-{numa[-1]} ({len(numa)} cpus)"
    except Exception as exc:  # containers without the affinity interface: keep the default placement
        numa = f"unchanged ({type(exc).__name__})"
    _shim.init(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, peak_src = measured_peaks()
    S = CWT["J"] + 1
    n0 = CWT["n0"]

    def make_series(count, seed):
        """Unit-variance AR(1) g=0.7 series generated on the device by the library's own Philox
        generator (synthetic cfg4 input; counter = global series index, so shards differ)."""
        pairs = (count + 1) // 2
        buf = torch.empty((pairs, 2, n0), dtype=torch.float32, device=dev)
        _shim.rednoise_device(CWT["ar1"], CWT["ar1"], n0, rank * pairs, pairs, seed, buf.data_ptr(),
                              stream=torch.cuda.current_stream().cuda_stream)
        y = buf.reshape(pairs * 2, n0)[:count] * (1 - CWT["ar1"] ** 2) ** 0.5
        return y.contiguous()

    # ------------------------------------------------------------------ main workload
    sampler = ClockSampler(local) if rank == 0 else None
    if args.workload == "cwt":
        B = args.series
        x = make_series(B, 1234)
        power = torch.empty((B, S, n0), dtype=torch.float32, device=dev)

        def step():
            engine.cwt_power_resident(x, power, DT, CWT["dj"], CWT["s0"], CWT["J"], CWT["f0"])
        units_per_step = B * S * n0
        metric, unit = "cwt_coeffs_per_sec", "coeff/s"
        alg_bytes = 4.0 * B * n0 * (1 + S)          # read x once, write the power plane once
        alg_flops = B * (5 * n0 * 10 + S * (5 * n0 * 10 + 5 * n0))
    else:
        R = args.realisations
        hist = torch.zeros((MC["J"] + 1, _shim.NBINS), dtype=torch.int64, device=dev)
        first = rank * R
        counter = {"k": 0}

        def step():
            # every step draws fresh realisations (global index advances), then the one collective
            base = (counter["k"] * world) * R + first
            counter["k"] += 1
            engine.wct_hist_resident(hist, MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"],
                                     base, R, MC["seed"])
            engine.reduce_histogram(hist)
        units_per_step = R
        metric, unit = "wct_mc_surrogates_per_sec", "surrogates/s"
        alg_bytes = 0.0
        alg_flops = R * 150e6                         # ~150 MFLOP per realisation
    if sampler:
        sampler.wait_ready()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _shim.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps - 1)]   # per-step spread, no syncs
    ev0.record()
    for i in range(args.steps):
        step()
        if i < args.steps - 1:
            marks[i].record()
    ev1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    edges = [ev0] + marks + [ev1]
    per_step = sorted(a.elapsed_time(b) for a, b in zip(edges[:-1], edges[1:]))
    step_spread = {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]}
    launches = _shim.kernel_launches() - launches0
    barrier()
    clocks = sampler.stop() if sampler else None
    value = world * units_per_step * args.steps / (ms * 1e-3)
    per_launch_s = ms * 1e-3 / args.steps

    # ------------------------------------------------------------------ end-to-end through the C ABI
    if args.workload == "cwt":
        Be = args.e2e_series
        xh = torch.empty((Be, n0), dtype=torch.float32).pin_memory()
        xh.copy_(x[:Be].cpu() if Be <= x.shape[0] else make_series(Be, 99).cpu())
        ph = torch.empty((Be, S, n0), dtype=torch.float32).pin_memory()
        lib = _shim.lib()

        def e2e_step():
            rc = lib.wtb_cwt_morlet(xh.data_ptr(), Be, n0, n0, DT, CWT["dj"], CWT["s0"], CWT["J"], CWT["f0"], 0,
                                    ph.data_ptr(), None, None)
            if rc != 0:
                raise RuntimeError(lib.wtb_last_error().decode())
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * Be * S * n0 * args.e2e_steps / e2e_s, "unit": unit,
               "h2d_bytes_per_step": 4 * Be * n0, "d2h_bytes_per_step": 4 * Be * S * n0,
               "series_per_step": Be, "steps": args.e2e_steps, "rank0_cpu_affinity": numa,
               "path": "wtb_cwt_morlet with pinned HOST buffers; H2D, kernels and D2H inside the timed region"}
        del xh, ph
    else:
        # host-facing call: histogram accumulated on device, copied back and reduced to thresholds each step
        Re = args.realisations   # same batch as the device-resident step: the call is not copy-bound

        def e2e_mc_step(k):
            h = _shim.wct_mc_hist(MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"],
                                  mc_first=10_000_000 + (k * world + rank) * Re, mc_count=Re, seed=MC["seed"],
                                  f64=False)
            return engine.significance_from_histogram(h, DT, MC["dj"], MC["s0"], MC["J"], 0.95, MC["f0"])
        e2e_mc_step(args.e2e_steps)          # untimed warm-up call (host staging buffers), like the cwt arm
        barrier()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            e2e_mc_step(k)
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * Re * args.e2e_steps / e2e_s, "unit": unit, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": 8 * (MC["J"] + 1) * _shim.NBINS, "realisations_per_step": Re,
               "steps": args.e2e_steps,
               "path": "wtb_wct_mc_hist (host histogram out) + wtb_wct_sig_from_hist; device RNG so no H2D payload"}

    # ------------------------------------------------------------------ secondary metric (bounded)
    secondary = None
    if args.workload == "cwt" and not args.no_secondary:
        R2 = args.secondary_realisations
        hist = torch.zeros((MC["J"] + 1, _shim.NBINS), dtype=torch.int64, device=dev)

        def mc_step(k):
            engine.wct_hist_resident(hist, MC["a1"], MC["a2"], DT, MC["dj"], MC["s0"], MC["J"], MC["f0"],
                                     (k * world + rank) * R2, R2, MC["seed"])
            engine.reduce_histogram(hist)
        for k in range(3):                    # warm-up: scratch arena growth, NCCL buffers, clocks
            mc_step(k)
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(3):
            mc_step(3 + k)
        e1.record()
        torch.cuda.synchronize()
        ms2 = max_over_ranks(e0.elapsed_time(e1))
        secondary = {"metric": "wct_mc_surrogates_per_sec", "value": world * R2 * 3 / (ms2 * 1e-3),
                     "unit": "surrogates/s", "realisations_per_gpu_per_step": R2, "steps": 3,
                     "config": "cfg5 shape: N=3351->4096, 66 scales, a1=0.989, a2=0.966, FP32, device Philox"}

    # MODWT LA8 J=6 (BASELINE cfg2 shape, batched): HBM-bound filterbank kernel, rank-local
    filterbank = None
    if args.workload == "cwt" and not args.no_secondary:
        from wavelet_transformer_b200 import pywt_compat as pywt
        la8 = pywt.Wavelet("sym4")
        Bf, nf, Jf = 50_000, 1024, 6
        xf = torch.randn((Bf, nf), dtype=torch.float64, device=dev)
        wf = torch.empty((Bf, Jf + 1, nf), dtype=torch.float64, device=dev)
        cur = torch.cuda.current_stream().cuda_stream

        def fb_step():
            _shim.modwt_device(xf.data_ptr(), Bf, nf, la8.dec_lo, la8.dec_hi, Jf, wf.data_ptr(), f64=True, stream=cur)
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
        filterbank = {"metric": "modwt_coeffs_per_sec", "value": Bf * (Jf + 1) * nf / fb_s, "unit": "coeff/s",
                      "per_gpu": True, "achieved_GBs": fb_bytes / fb_s / 1e9, "frac_hbm": fb_bytes / fb_s / 1e9 / hbm_peak,
                      "config": f"MODWT LA8 (sym4) J={Jf}, {Bf} series x N={nf}, FP64, device-resident; "
                                "algorithmic bytes 8 N (1 + J+1) per series"}
        del xf, wf

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak_nominal = 148 * 128 * 2 * 1.965e9
    fp32_peak_at_clock = 148 * 128 * 2 * sm_mhz * 1e6
    if args.workload == "cwt":
        achieved = alg_bytes / per_launch_s / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak,
                    "traffic": (args.traffic_bytes if args.traffic_bytes is not None else
                                (measured_traffic_per_series() or 0) * args.series or None),
                    "traffic_source": "profiles/r1_traffic.json: dram read+write bytes per series from one "
                                      "ncu --set full capture, scaled to this launch's series count",
                    "peak_source": peak_src, "kernel": "fused CWT+power (one launch per step)",
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "fp32": {"algorithmic_flop_per_launch": alg_flops,
                             "achieved_tflops": alg_flops / per_launch_s / 1e12,
                             "peak_tflops_nominal": fp32_peak_nominal / 1e12,
                             "frac_nominal": alg_flops / per_launch_s / fp32_peak_nominal,
                             "frac_at_measured_clock": alg_flops / per_launch_s / fp32_peak_at_clock,
                             "note": "5*N*log2N per FFT convention (SURVEY 8d): 6.81 MFLOP per series; "
                                     "non-tensor FP32 peak = 148 SM x 128 lanes x 2 x clock"}}
    else:
        ach = alg_flops / per_launch_s / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": fp32_peak_nominal / 1e12, "unit": "TFLOP/s",
                    "frac": ach / (fp32_peak_nominal / 1e12), "traffic": args.traffic_bytes,
                    "peak_source": "nominal non-tensor FP32 peak (148 SM x 128 x 2 x 1.965 GHz); this path is "
                                   "FP32-FLOP bound, tensor cores are not applicable (no dense contraction)",
                    "kernel": "k_wct_spec_4096 (CWT + cross spectrum + Gaussian time filter per scale row; 56% of "
                              "the step, with k_wct_coh_4096 30% and k_wct_boxcar_4096 9%: "
                              "profiles/r1_launches_mc_final.csv)",
                    "algorithmic_flop_per_launch": alg_flops}
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "ms_per_step_spread": step_spread,
    }
    if secondary:
        line["secondary"] = secondary
    if filterbank:
        line["secondary_filterbank"] = filterbank
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_budget)
        line["cpu_baseline"]["host_cpus"] = os.cpu_count()
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cwt", choices=["cwt", "wct_mc"])
    ap.add_argument("--series", type=int, default=125_000, help="series per GPU per step (cfg4 shard)")
    ap.add_argument("--realisations", type=int, default=2048, help="MC realisations per GPU per step")
    ap.add_argument("--secondary-realisations", type=int, default=2048)
    ap.add_argument("--e2e-series", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

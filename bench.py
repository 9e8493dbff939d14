#!/usr/bin/env python
"""bench.py -- force+energy evaluations per second of the SlicedNonbondedForce hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--impl reference]

One "step" is one complete evaluation (neighbour list, direct space, exceptions, PME, all slice
energies delivered to the host) of one synthetic system:
  * N = 1 : C3, the DHFR-size system BASELINE.json's target is quoted on (override with --workload);
  * N > 1 : C5, the STMV-size system, strong scaling -- direct-space i-blocks and PME subset grids are
            split across ranks, forces/energies combined with an NCCL all-reduce.
Prints ONE JSON line (contract in the task statement): `value` is device-resident throughput timed
with CUDA events (L2 flushed between steps, outside the events), `e2e` the same metric through the
host-buffer API (host->device copy of positions and device->host copy of forces + energies inside
the timed region), `roofline` the FP32-pipe fraction of the pair kernel, `cpu_baseline` the
reference's own Reference-platform arithmetic (oracle/_ref) on this box's host cores.

`--impl reference` times that CPU implementation instead (rank 0 only).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 72            # FP32-equivalent flops per interacting pair
FS_PER_STEP = 2.0             # ns/day figure assumes one evaluation per 2 fs step
# DRAM traffic of one k_pair launch from the committed ncu captures (profiles/README.md): the kernel's working
# set (positions, parameters, lists) is L2-resident, so this is far below any bandwidth limit
PAIR_TRAFFIC_BYTES = {"C3": 8676608}
# the same for the reciprocal chain (k_spread + the FFT / convolution kernels + k_gather, summed over the launches of
# one evaluation): profiles/r01_ncu_pme_summary.txt
PME_TRAFFIC_BYTES = {"C3": 3370752 + 6334976 + 7616256 + 6531072 + 2742272, "C5": 1084900000}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--workload", default=None, help="C1..C5 (default: C3 at 1 GPU, C5 at N > 1)")
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    return p.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs: NVML every millisecond when the
    binding is importable (the timed region of a small workload lasts only milliseconds), else nvidia-smi."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.samples, self.stop, self.index = [], threading.Event(), index
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(bits & getattr(n, name, 0)) for name in
                 ("nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                  "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
        return [str(mhz), str(self.max_mhz), ""] + ["Active" if f else "Not Active" for f in flags]

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    self.stop.wait(0.001)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def sample_now(self):
        """One sample from the calling thread (the background thread can be starved by a busy timed loop)."""
        try:
            if self.nvml is not None:
                self.samples.append(self._sample_nvml())
        except Exception:
            pass

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = [n for k, n in enumerate(self.NAMES) if any(len(s) > 3+k and s[3+k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def pair_roofline(pair_count, pair_ms, tile_efficiency, traffic=None, note=None):
    """FP32-pipe roofline of the pair kernel: 72 flop per interacting pair over the kernel's
    CUDA-event duration, against 148 SMs x 128 lanes x 2 flop x the measured maximum SM clock."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fp32_peak = 148*128*2*float(peaks.get("sm_max_mhz", 1965.0))*1e6/1e12
    achieved = FLOP_PER_PAIR*pair_count/(pair_ms*1e-3)/1e12
    out = {"bound": "fp32", "kernel": "k_pair", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
           "frac": achieved/fp32_peak, "traffic": traffic,
           "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json (that file has no FP32 figure)"
                          if "sm_max_mhz" in peaks else "fallback: 148 SM x 128 lanes x 2 x 1965 MHz",
           "algorithmic": f"{FLOP_PER_PAIR} flop x {pair_count} interacting pairs", "kernel_ms": pair_ms}
    if tile_efficiency is not None:
        out["tile_efficiency"] = tile_efficiency
    if note:
        out["note"] = note
    return out


def ns_per_day(evals_per_s, fs=FS_PER_STEP):
    return evals_per_s*86400*fs*1e-6


CONFIG_ATOMS = {"C1": 648, "C2": 7530, "C3": 23558, "C4": 92224, "C5": 1066628}


def load_description(name):
    systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
    return systems.CONFIGS[name]["description"]


def load_workload(name):
    systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
    return systems.make_system(name)


REFERENCE_SECONDS_PER_EVAL = {"C1": 0.03, "C2": 0.4, "C3": 1.3, "C4": 9.0, "C5": 420.0}   # 16 host cores, measured


def run_reference(args, workload_name):
    """The reference's own CPU implementation of the path (oracle/_ref: the plugin's unmodified Reference-platform
    translation units), all host threads it can use.  One step = one full evaluation of a BOUNDED SAMPLE of the
    workload: the largest BASELINE configuration whose (warmup + steps) evaluations finish within ~3 minutes;
    the reported value is scaled to the workload by the atom ratio (cost per atom is constant at equal density
    and cutoff: neighbour list and direct space are O(N), the PME grid grows with N)."""
    from oracle import oracle
    nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
    kind = "reference" if oracle.available("reference") else "port"
    order = ["C5", "C4", "C3", "C2", "C1"]
    candidates = order[order.index(workload_name):] if workload_name in order else [workload_name]
    sample_name = candidates[-1]
    for name in candidates:
        if REFERENCE_SECONDS_PER_EVAL.get(name, 1e9)*(args.warmup + args.steps) <= 180.0:
            sample_name = name
            break
    full = load_workload(workload_name) if sample_name == workload_name else None
    s = full if full is not None else load_workload(sample_name)
    n_full = CONFIG_ATOMS.get(workload_name, s.force.getNumParticles())
    n_sample = s.force.getNumParticles()
    desc = nbs.build_desc(s.system, s.force)
    lam = np.ones((s.force.getNumSlices(), 2))
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = oracle.evaluate(desc, s.positions, s.box, lam, kind=kind)
        dt = time.perf_counter()-t0
        if it >= args.warmup:
            times.append(dt)
    ms_sample = 1e3*float(np.mean(times))
    ms = ms_sample*n_full/n_sample
    value = 1e3/ms
    description = load_description(workload_name)
    sample = "full evaluation per step (neighbour list + direct + PME); single-threaded except pocketfft"
    if sample_name != workload_name:
        sample = (f"each step = one full evaluation of {sample_name} ({n_sample} atoms, same density, cutoff and PME accuracy), "
                  f"{ms_sample:.1f} ms; value scaled by the atom ratio {n_full}/{n_sample} to {workload_name}")
    line = {
        "impl": "reference", "metric": "force+energy evals/s", "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{workload_name}: {description}", "ns_per_day_2fs": ns_per_day(value)},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": os.cpu_count(), "kind": kind, "sample": sample,
                         "breakdown_s_of_sample": {k: float(v) for k, v in res.timings.items()}},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_baseline(workload_name, s, desc, seconds):
    from oracle import oracle
    kind = "reference" if oracle.available("reference") else "port"
    lam = np.ones((s.force.getNumSlices(), 2))
    t_start, times, last = time.perf_counter(), [], None
    while True:
        t0 = time.perf_counter()
        last = oracle.evaluate(desc, s.positions, s.box, lam, kind=kind)
        times.append(time.perf_counter()-t0)
        if time.perf_counter()-t_start > seconds or len(times) >= 20:
            break
    value = 1.0/float(np.mean(times))
    return {"value": value, "unit": "evals/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"{len(times)} full evaluation(s) of {workload_name} (neighbour list + direct + PME)",
            "threads": "1 (pocketfft may use all cores)",
            "breakdown_s": {k: float(v) for k, v in last.timings.items()}}, last


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload_name = args.workload or ("C3" if args.gpus == 1 else "C5")
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, workload_name)
        return
    import torch
    nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
    abi = nbs.abi
    if world > 1:
        from importlib import import_module
        multi = import_module("openmm-nonbonded-slicing_b200.multigpu")
        return multi.bench_main(args, workload_name)

    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    s = load_workload(workload_name)
    n = s.force.getNumParticles()
    nsl = s.force.getNumSlices()
    lam = np.ones((nsl, 2))

    # ---- device-resident throughput ------------------------------------------------------------
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
    kernel.initialize(s.system, s.force)
    pos_dev = torch.tensor(s.positions, dtype=torch.float64, device=dev).contiguous()
    frc_dev = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        return kernel.execute_device(pos_dev.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    launches_before = kernel.getLaunchCount()
    per_step = []
    with ClockSampler() as clocks:
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            start.record()
            energies = step_device()
            end.record()
            clocks.sample_now()                      # right after the step (the call is synchronous; boost state outlives it)
            torch.cuda.synchronize()
            per_step.append(start.elapsed_time(end))
        t_wall = time.perf_counter()-t_wall0
    launches = kernel.getLaunchCount()-launches_before
    ms = float(np.mean(per_step))
    value = 1e3/ms

    # ---- the same evaluation without slice energies (what an MD step between reports asks for) --------
    forces_only = []
    for it in range(args.warmup + args.steps):
        flush.fill_(1)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        kernel.execute_device(pos_dev.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream, want_energies=False)
        end.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            forces_only.append(start.elapsed_time(end))
    forces_ms = float(np.mean(forces_only))
    # leave the force buffer as the full evaluation produced it (the parity check below reads it)
    step_device()
    torch.cuda.synchronize()

    # ---- end to end through the host-buffer API (what a plugin user calls) ---------------------
    pos_host = torch.tensor(s.positions, dtype=torch.float64).pin_memory()
    frc_host = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
    pos_np, frc_np = pos_host.numpy(), frc_host.numpy()
    e2e_times = []
    for it in range(args.warmup + args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        frc_np[:] = 0
        e_host = kernel._evaluate(pos_np, s.box, lam, np.zeros(0), True, True, frc_np)
        dt = time.perf_counter()-t0
        if it >= args.warmup:
            e2e_times.append(dt)
    e2e_value = 1.0/float(np.mean(e2e_times))

    # ---- per-kernel durations (CUDA events around every kernel, separate profiled context) ------
    prof = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=abi.NBS_FLAG_PROFILE))
    prof.initialize(s.system, s.force)
    acc = {}
    reps = 10
    for it in range(3 + reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        prof.execute_device(pos_dev.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream)
        if it >= 3:
            for name, t in prof.getKernelTimes():
                acc[name] = acc.get(name, 0.0) + t/reps
    pair_count, _, _ = prof.getPairSet(with_pairs=False)
    stats = prof.getNlistStats()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    pair_ms = acc.get("pair", float("nan"))
    grid = s.force.getPMEParameters()[1]
    G = grid**3
    pme_bytes = 32*s.force.getNumSubsets()*G + 52*n
    pme_ms = sum(acc.get(k, 0.0) for k in ("spread", "fft_fwd", "fft_conv_inv", "gather"))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # dram__bytes_read.sum + dram__bytes_write.sum of k_pair from profiles/ (ncu --set full, per launch); C3 only
    traffic = PAIR_TRAFFIC_BYTES.get(workload_name)
    roofline = pair_roofline(pair_count, pair_ms, pair_count/max(stats[3], 1), traffic)
    roofline_pme = {"bound": "hbm", "kernels": "k_spread + 3 plane-fused FFT/convolution kernels + k_gather", "achieved": pme_bytes/(pme_ms*1e-3)/1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": pme_bytes/(pme_ms*1e-3)/1e9/hbm_peak,
                    "peak_source": "hbm_gbs of MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "algorithmic": f"32*nS*G + 52*N = {pme_bytes} bytes", "kernel_ms": pme_ms,
                    "traffic": PME_TRAFFIC_BYTES.get(workload_name)}

    line = {
        "metric": "force+energy evals/s", "value": value, "unit": "evals/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload_name}: {s.description}", "atoms": n, "subsets": s.force.getNumSubsets(),
                   "pme_grid": grid, "cutoff_nm": 1.0, "ns_per_day_2fs": ns_per_day(value), "ns_per_day_4fs": ns_per_day(value, 4.0),
                   "l2": "256 MiB buffer written between steps, outside the per-step CUDA events",
                   "neighbour_list": "rebuilt from scratch every step", "interacting_pairs": pair_count,
                   "wall_s_timed_loop": t_wall},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(pos_np.nbytes),
                "d2h_bytes_per_step": int(frc_np.nbytes + 8*2*36 + 64), "ns_per_day_2fs": ns_per_day(e2e_value)},
        "gpu_launches": int(launches),
        "forces_only": {"ms_per_step": forces_ms, "value": 1e3/forces_ms, "unit": "evals/s", "ns_per_day_2fs": ns_per_day(1e3/forces_ms),
                        "note": "same evaluation without slice energies / dE/dlambda (single-precision PME grids, no double-precision pair energies)"},
        "roofline": roofline,
        "roofline_pme": roofline_pme,
        "kernel_ms": {k: round(v, 5) for k, v in acc.items()},
        "slice_energy_checksum": float(np.abs(energies).sum()),
    }
    if not args.no_cpu_baseline:
        desc = kernel.desc
        base, ref = cpu_baseline(workload_name, s, desc, args.cpu_baseline_seconds)
        line["cpu_baseline"] = base
        scale = np.maximum(np.abs(ref.slice_energies), 1.0)
        frc_check = frc_dev.cpu().numpy()
        line["parity_vs_cpu_baseline"] = {
            "force_rel_rms": float(np.sqrt(((frc_check-ref.forces)**2).sum()/(ref.forces**2).sum())),
            "max_energy_err_over_max_absE_1": float(np.max(np.abs(e_host-ref.slice_energies)/scale)),
        }
    print(json.dumps(line))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- force+energy evaluations per second of the SlicedNonbondedForce hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--impl reference]

One "step" is one complete evaluation (neighbour list upkeep, direct space, exceptions, PME, all slice
energies delivered to the host) of one synthetic system WHOSE ATOMS MOVE between steps (every atom has its
own thermal velocity, config.motion), so the neighbour list is re-used and rebuilt as it would be in a run:
  * plain `python bench.py` (N = 1): C3, the DHFR-size system BASELINE.json's target is quoted on
    (override with --workload); the line also carries C5 on this one GPU as the anchor of the scaling curve;
  * under torch.distributed.run (WORLD_SIZE set, any N including 1): C5, the STMV-size system, strong
    scaling -- direct-space i-blocks and x-slabs of the PME grids split over all ranks, x pass and force
    reduction over NVLink peer memory.
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
THERMAL_SIGMA_NM = 0.0009     # displacement per step and component: ~300 K, 12 amu, 2 fs (an oxygen; hydrogens move 3x further
                              # in reality, rigid water constraints slow them again -- one figure for every atom keeps it simple)
MOTION = (f"ballistic: every atom has its own velocity, N(0, {THERMAL_SIGMA_NM} nm) per component and step (about 300 K at 2 fs), "
          "reversed every 48 steps; positions are advanced between steps, outside the timed region")
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
    p.add_argument("--no-scale-anchor", action="store_true", help="skip the C5-on-one-GPU measurement of the plain N = 1 run")
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


_MEASURED_FP32 = {}


def measured_fp32_peak(device=0):
    """Dense FP32 FMA rate of this GPU from the library's dependent-free FFMA micro-benchmark (nbs_measure_peaks,
    csrc/k_peak.cu), in TFLOP/s; MEASURED_PEAKS.json has no FP32 figure."""
    if device not in _MEASURED_FP32:
        import ctypes as C
        nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
        out = (C.c_double*4)()
        best = 0.0
        for _ in range(2):
            nbs.abi.check(nbs.abi.load_library().nbs_measure_peaks(device, out))
            best = max(best, float(out[0]))
        _MEASURED_FP32[device] = best
    return _MEASURED_FP32[device]


def pair_roofline(pair_count, pair_ms, tile_efficiency, traffic=None, note=None, device=0):
    """FP32-pipe roofline of the pair kernel: 72 flop per interacting pair over the kernel's
    CUDA-event duration, against the FP32 FMA rate measured on this GPU (dependent-free FFMA micro-benchmark);
    the nominal 148 SMs x 128 lanes x 2 flop x the maximum SM clock is carried beside it."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    nominal = 148*128*2*float(peaks.get("sm_max_mhz", 1965.0))*1e6/1e12
    try:
        fp32_peak, source = measured_fp32_peak(device), "FFMA micro-benchmark on this GPU in this run (nbs_measure_peaks); MEASURED_PEAKS.json has no FP32 figure"
    except Exception as exc:                      # noqa: BLE001 -- the roofline then says which fallback it used
        fp32_peak, source = nominal, f"fallback: 148 SM x 128 lanes x 2 x sm_max_mhz ({exc})"
    achieved = FLOP_PER_PAIR*pair_count/(pair_ms*1e-3)/1e12
    out = {"bound": "fp32", "kernel": "k_pair", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
           "frac": achieved/fp32_peak, "traffic": traffic, "peak_source": source, "peak_nominal": nominal,
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


REFERENCE_SECONDS_PER_EVAL = {"C1": 0.03, "C2": 0.4, "C3": 1.3, "C4": 9.0, "C5": 120.0}   # 16 host cores, measured
REFERENCE_CACHE = os.path.join(ROOT, "baseline", "_ref", "reference_arm_cache.json")      # git-ignored, like baseline/_ref itself


def run_reference(args, workload_name):
    """The reference's own CPU implementation of the path (oracle/_ref: the plugin's unmodified Reference-platform
    translation units), all host threads it can use.  One step = one REAL full evaluation of the workload itself --
    never a scaled-down sample: when (warmup + steps) evaluations would not finish within minutes (C5: about two
    minutes each), ONE evaluation is timed, the line says steps = 1 / warmup = 0, and the measurement is cached for
    the other invocations of the same round (labelled `cached`)."""
    from oracle import oracle
    nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
    kind = "reference" if oracle.available("reference") else "port"
    description = load_description(workload_name)
    cfg = importlib.import_module("openmm-nonbonded-slicing_b200.systems").CONFIGS.get(workload_name, {})
    warmup, steps = args.warmup, args.steps
    if REFERENCE_SECONDS_PER_EVAL.get(workload_name, 1.0)*(warmup + steps) > 240.0:
        warmup, steps = 0, 1
    cached = None
    try:
        cached = json.load(open(REFERENCE_CACHE)).get(f"{workload_name}:{kind}")
    except Exception:
        pass
    t_run0 = time.perf_counter()
    if cached is not None and steps == 1:
        ms, timings, n_evals = cached["ms_per_step"], cached["breakdown_s"], 0
    else:
        s = load_workload(workload_name)
        desc = nbs.build_desc(s.system, s.force)
        lam = np.ones((s.force.getNumSlices(), 2))
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            res = oracle.evaluate(desc, s.positions, s.box, lam, kind=kind)
            dt = time.perf_counter()-t0
            if it >= warmup:
                times.append(dt)
        ms = 1e3*float(np.mean(times))
        timings = {k: float(v) for k, v in res.timings.items()}
        n_evals = warmup + steps
        if steps == 1:
            try:
                os.makedirs(os.path.dirname(REFERENCE_CACHE), exist_ok=True)
                old = {}
                try:
                    old = json.load(open(REFERENCE_CACHE))
                except Exception:
                    pass
                old[f"{workload_name}:{kind}"] = {"ms_per_step": ms, "breakdown_s": timings, "cores": os.cpu_count()}
                json.dump(old, open(REFERENCE_CACHE, "w"))
            except Exception:
                pass
    value = 1e3/ms
    sample = (f"{steps} full evaluation(s) of {workload_name} itself after {warmup} warm-up (neighbour list + direct + PME); "
              "single-threaded except pocketfft; about two thirds of the time is the oracle's own restated neighbour list, "
              "the rest the reference's arithmetic (breakdown_s)")
    line = {
        "impl": "reference", "metric": "force+energy evals/s", "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "requested": {"steps": args.steps, "warmup": args.warmup},
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{workload_name}: {description}", "atoms": CONFIG_ATOMS.get(workload_name),
                   "subsets": len(cfg.get("solute", [])) + 1 + (1 if workload_name == "C1" else 0), "pme_grid": cfg.get("grid"),
                   "cutoff_nm": 1.0, "ns_per_day_2fs": ns_per_day(value),
                   "motion": "none: the CPU arm evaluates the un-moved configuration (its cost does not depend on the positions; "
                             "it rebuilds its neighbour list on every evaluation, like the Reference platform)"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": os.cpu_count(), "kind": kind, "sample": sample,
                         "breakdown_s": timings},
        "cached": n_evals == 0, "evaluations_run_now": n_evals, "run_s": time.perf_counter()-t_run0,
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


class MovingSystem:
    """Device-resident positions that advance between steps (MOTION): pos(t) = pos0 + t * velocity, identical on every
    rank (seeded).  `advance` runs outside the timed region -- it is the integrator's work, not the force's."""

    def __init__(self, positions, device, seed=1):
        import torch
        rng = np.random.default_rng(seed)
        self.vel_host = rng.normal(scale=THERMAL_SIGMA_NM, size=positions.shape)
        self.pos0_host = np.ascontiguousarray(positions, dtype=np.float64)
        self.pos0 = torch.tensor(self.pos0_host, dtype=torch.float64, device=device)
        self.vel = torch.tensor(self.vel_host, dtype=torch.float64, device=device)
        self.pos = self.pos0.clone()

    TURN = 48          # steps after which every velocity is reversed

    @classmethod
    def phase(cls, t):
        """Steps' worth of displacement at step t: 0, 1, .., TURN, TURN-1, .., 0, 1, ..  A long run (--steps 500) then
        never carries atoms further than TURN steps from the generated configuration -- far enough for the list to be
        rebuilt every few steps, not so far that atoms overlap and forces leave the 64-bit fixed-point range."""
        k = t % (2*cls.TURN)
        return k if k <= cls.TURN else 2*cls.TURN - k

    def advance(self, t):
        import torch
        torch.add(self.pos0, self.vel, alpha=float(self.phase(t)), out=self.pos)

    def host_positions(self, t, lo=0, hi=None):
        return self.pos0_host[lo:hi] + self.phase(t)*self.vel_host[lo:hi]


MIN_WARMUP = 24


def timed_steps(step, moving, flush, warmup, steps, barrier=None, t0=0):
    """`warmup` untimed and `steps` timed evaluations along the trajectory (from its step t0); per-step CUDA events on the
    current stream, L2 flushed and positions advanced outside the events.  Returns the per-step milliseconds and the
    last result."""
    import torch
    per_step, result = [], None
    for t in range(warmup + steps):
        moving.advance(t0 + t)
        flush.fill_(1)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
            torch.cuda.synchronize()
        start.record()
        result = step()
        end.record()
        torch.cuda.synchronize()
        if t >= warmup:
            per_step.append(start.elapsed_time(end))
    return per_step, result


def list_policy(stats_before, stats_after):
    evals = stats_after["evaluations"] - stats_before["evaluations"]
    builds = stats_after["builds"] - stats_before["builds"]
    return {"skin_nm": stats_after["skin"], "evaluations": int(evals), "list_builds": int(builds),
            "redone_evaluations": int(stats_after["redone"] - stats_before["redone"]),
            "rebuild_fraction": builds/max(evals, 1)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    distributed = "WORLD_SIZE" in os.environ
    workload_name = args.workload or ("C5" if (distributed or args.gpus > 1) else "C3")
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, workload_name)
        return
    import torch
    nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
    abi = nbs.abi
    if distributed:
        from importlib import import_module
        multi = import_module("openmm-nonbonded-slicing_b200.multigpu")
        return multi.bench_main(args, workload_name)
    if args.gpus > 1:
        raise SystemExit("bench.py --gpus N with N > 1 must be launched with torch.distributed.run (one rank per GPU)")

    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    s = load_workload(workload_name)
    n = s.force.getNumParticles()
    nsl = s.force.getNumSlices()
    lam = np.ones((nsl, 2))
    # (at least 24 untimed steps: along the trajectory the evaluations that build a list and those that re-use one are two
    # different CUDA graphs, each captured on its second occurrence -- the timed region should contain replays only)
    warmup = max(args.warmup, MIN_WARMUP)

    # ---- device-resident throughput, atoms moving, neighbour list re-used with the default skin ----------------
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
    kernel.initialize(s.system, s.force)
    moving = MovingSystem(s.positions, dev)
    frc_dev = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device(k=kernel, want=True):
        return k.execute_device(moving.pos.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream, want_energies=want)

    timed_steps(step_device, moving, flush, warmup, 0)
    launches_before, stats_before = kernel.getLaunchCount(), kernel.getListStats()
    with ClockSampler() as clocks:
        t_wall0 = time.perf_counter()
        per_step, energies = timed_steps(step_device, moving, flush, 0, args.steps, t0=warmup)
        clocks.sample_now()
        t_wall = time.perf_counter()-t_wall0
    launches = kernel.getLaunchCount()-launches_before
    policy = list_policy(stats_before, kernel.getListStats())
    ms = float(np.mean(per_step))
    value = 1e3/ms

    # ---- the same trajectory with the list rebuilt on every step (what the Reference platform does), and the same
    # evaluation of positions that do not move (the list is never rebuilt) -------------------------------------------
    rebuild = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=abi.NBS_FLAG_NO_LIST_REUSE))
    rebuild.initialize(s.system, s.force)
    rebuild_ms = float(np.mean(timed_steps(lambda: step_device(rebuild), moving, flush, warmup, args.steps)[0]))
    del rebuild
    frozen = MovingSystem(s.positions, dev)
    frozen.vel.zero_()
    static_ms = float(np.mean(timed_steps(lambda: kernel.execute_device(frozen.pos.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream),
                                          frozen, flush, warmup, args.steps)[0]))

    # ---- the moving system without slice energies (what an MD step between reports asks for) --------------------------
    forces_ms = float(np.mean(timed_steps(lambda: step_device(want=False), moving, flush, warmup, args.steps)[0]))

    # ---- end to end through the host-buffer API (what a plugin user calls): host positions in, host forces out -------
    pos_host = torch.tensor(s.positions, dtype=torch.float64).pin_memory()
    frc_host = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
    pos_np, frc_np = pos_host.numpy(), frc_host.numpy()
    e2e_times = []
    no_globals = np.zeros(0)
    e2e_warm = max(args.warmup, 3)
    for it in range(e2e_warm + args.steps):
        pos_np[:] = moving.host_positions(it)
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e_host = kernel._evaluate(pos_np, s.box, lam, no_globals, True, True, frc_np, accumulate=False)
        dt = time.perf_counter()-t0
        if it >= e2e_warm:
            e2e_times.append(dt)
    e2e_value = 1.0/float(np.mean(e2e_times))

    # ---- parity of THIS run against the CPU baseline: the un-moved positions, full evaluation ---------------------------
    frc_np[:] = 0
    e_host = kernel._evaluate(np.ascontiguousarray(s.positions), s.box, lam, np.zeros(0), True, True, frc_np)
    frc_check = frc_np.copy()

    # ---- per-kernel durations (CUDA events around every kernel, separate profiled context, list re-used) -------------
    prof = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=abi.NBS_FLAG_PROFILE))
    prof.initialize(s.system, s.force)
    acc = {}
    reps = 10
    for it in range(3 + reps):
        moving.advance(0)
        flush.fill_(1)
        torch.cuda.synchronize()
        prof.execute_device(moving.pos.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream)
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
    roofline = pair_roofline(pair_count, pair_ms, pair_count/max(stats[3], 1), traffic,
                             note="kernel_ms: the list built with the skin and re-used (the steady state of the moving system)")
    roofline_pme = {"bound": "hbm", "kernels": "k_spread + 3 plane-fused FFT/convolution kernels + k_gather", "achieved": pme_bytes/(pme_ms*1e-3)/1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": pme_bytes/(pme_ms*1e-3)/1e9/hbm_peak,
                    "peak_source": "hbm_gbs of MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "algorithmic": f"32*nS*G + 52*N = {pme_bytes} bytes", "kernel_ms": pme_ms,
                    "traffic": PME_TRAFFIC_BYTES.get(workload_name)}

    line = {
        "metric": "force+energy evals/s", "value": value, "unit": "evals/s", "n_gpus": 1, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{workload_name}: {s.description}", "atoms": n, "subsets": s.force.getNumSubsets(),
                   "pme_grid": grid, "cutoff_nm": 1.0, "ns_per_day_2fs": ns_per_day(value), "ns_per_day_4fs": ns_per_day(value, 4.0),
                   "l2": "256 MiB buffer written between steps, outside the per-step CUDA events",
                   "motion": MOTION,
                   "neighbour_list": "built with a skin, re-used until an atom has moved half of it (device-side check in every evaluation)",
                   "list_policy": policy, "interacting_pairs": pair_count,
                   "wall_s_timed_loop": t_wall},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(pos_np.nbytes),
                "d2h_bytes_per_step": int(frc_np.nbytes + 8*2*36 + 64), "ns_per_day_2fs": ns_per_day(e2e_value),
                "note": "host positions (pinned) in, host forces (pinned, overwritten) and slice energies out, every step, through "
                        "nbs_execute with host buffers; wall clock around the call"},
        "gpu_launches": int(launches),
        "rebuild_every_step": {"ms_per_step": rebuild_ms, "value": 1e3/rebuild_ms, "unit": "evals/s",
                               "note": "same trajectory, neighbour list rebuilt from scratch on every step (NBS_FLAG_NO_LIST_REUSE), like the Reference platform"},
        "static_positions": {"ms_per_step": static_ms, "value": 1e3/static_ms, "unit": "evals/s",
                             "note": "positions that do not move: the list is never rebuilt (round 1 measured this case with a rebuild every step)"},
        "forces_only": {"ms_per_step": forces_ms, "value": 1e3/forces_ms, "unit": "evals/s", "ns_per_day_2fs": ns_per_day(1e3/forces_ms),
                        "note": "same trajectory without slice energies / dE/dlambda (single-precision PME grids, no double-precision pair energies)"},
        "roofline": roofline,
        "roofline_pme": roofline_pme,
        "kernel_ms": {k: round(v, 5) for k, v in acc.items()},
        "slice_energy_checksum": float(np.abs(energies).sum()),
        "reference_cuda": {"value": None, "note": "the plugin's own CUDA platform needs OpenMM, which neither this image nor the GPU box has "
                                                  "(profiles/r02_probe_openmm.log): the 10x target of BASELINE.json has no measured denominator"},
    }
    del prof
    if not args.no_scale_anchor and workload_name == "C3":
        # the scaling curve (bench.py under torch.distributed.run) is quoted on C5: its one-GPU point through the plain,
        # unsharded library, same protocol
        try:
            big = load_workload("C5")
            kb = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
            kb.initialize(big.system, big.force)
            mb = MovingSystem(big.positions, dev)
            fb = torch.zeros((big.force.getNumParticles(), 3), dtype=torch.float64, device=dev)
            lb = np.ones((big.force.getNumSlices(), 2))
            tb = timed_steps(lambda: kb.execute_device(mb.pos.data_ptr(), big.box, fb.data_ptr(), lb, stream=stream), mb, flush, 3, 10)[0]
            line["scale_anchor"] = {"workload": f"C5: {big.description}", "n_gpus": 1, "ms_per_step": float(np.mean(tb)),
                                    "value": 1e3/float(np.mean(tb)), "unit": "evals/s", "steps": 10, "warmup": 3,
                                    "note": "C5 on this one GPU, unsharded library, same moving-system protocol: the N = 1 point of the scaling curve"}
            del kb, mb, fb
        except Exception as exc:                       # noqa: BLE001
            line["scale_anchor"] = {"error": str(exc)}
    if not args.no_cpu_baseline:
        desc = kernel.desc
        base, ref = cpu_baseline(workload_name, s, desc, args.cpu_baseline_seconds)
        line["cpu_baseline"] = base
        scale = np.maximum(np.abs(ref.slice_energies), 1.0)
        line["parity_vs_cpu_baseline"] = {
            "force_rel_rms": float(np.sqrt(((frc_check-ref.forces)**2).sum()/(ref.forces**2).sum())),
            "max_energy_err_over_max_absE_1": float(np.max(np.abs(e_host-ref.slice_energies)/scale)),
            "positions": "the un-moved configuration",
        }
    print(json.dumps(line))


if __name__ == "__main__":
    main()

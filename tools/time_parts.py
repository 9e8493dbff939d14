"""Whole-evaluation time (graph replay, device-resident positions) of the full evaluation, direct space alone and
reciprocal space alone, with and without slice energies -- shows how much of the two concurrent chains overlaps.
usage: time_parts.py [config] [reps]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
s = systems.make_system(name)
n = s.force.getNumParticles()
pos = torch.tensor(s.positions, dtype=torch.float64, device="cuda")
frc = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
lam = np.ones((s.force.getNumSlices(), 2))
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
out = []
for label, direct, recip, energy in (("full+E", True, True, True), ("direct+E", True, False, True), ("recip+E", False, True, True),
                                     ("full", True, True, False), ("direct", True, False, False), ("recip", False, True, False)):
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
    kernel.initialize(s.system, s.force)
    times = []
    for it in range(6 + reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        kernel.execute_device(pos.data_ptr(), s.box, frc.data_ptr(), lam, includeDirect=direct, includeReciprocal=recip,
                              want_energies=energy, stream=stream)
        b.record()
        torch.cuda.synchronize()
        if it >= 6:
            times.append(1e3*a.elapsed_time(b))
    # back-to-back calls, wall clock, no flush: what the host sees per call
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(reps):
        kernel.execute_device(pos.data_ptr(), s.box, frc.data_ptr(), lam, includeDirect=direct, includeReciprocal=recip,
                              want_energies=energy, stream=stream)
    wall = 1e6*(time.perf_counter() - t0)/reps
    out.append(f"{label}={np.mean(times):.1f}/{np.min(times):.1f}/wall {wall:.1f}")
print(os.environ.get("NBS_B200_LIBRARY", "default").split("/")[-1], name, "eval_us mean/min/back-to-back wall:", "  ".join(out))

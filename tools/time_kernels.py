"""Per-kernel CUDA-event times (profiled context, serial) and whole-evaluation time (graph replay) of one
workload.  NBS_B200_LIBRARY selects an alternative build of the CUDA library (tuning experiments)."""
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
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
want_energy = (sys.argv[3] != "forces") if len(sys.argv) > 3 else True
s = systems.make_system(name)
n = s.force.getNumParticles()
pos = torch.tensor(s.positions, dtype=torch.float64, device="cuda")
frc = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
lam = np.ones((s.force.getNumSlices(), 2))
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")

precision = os.environ.get("NBS_PRECISION", "mixed")          # single | mixed | double (the platform's Precision property)
prof = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=nbs.abi.NBS_FLAG_PROFILE, properties={"Precision": precision}))
prof.initialize(s.system, s.force)
acc = {}
for it in range(3 + reps):
    flush.fill_(1)
    torch.cuda.synchronize()
    prof.execute_device(pos.data_ptr(), s.box, frc.data_ptr(), lam, want_energies=want_energy)
    if it >= 3:
        for k, t in prof.getKernelTimes():
            acc[k] = acc.get(k, 0.0) + 1e3*t/reps
kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(properties={"Precision": precision}))
kernel.initialize(s.system, s.force)
times = []
for it in range(5 + reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    kernel.execute_device(pos.data_ptr(), s.box, frc.data_ptr(), lam, want_energies=want_energy)
    b.record()
    torch.cuda.synchronize()
    if it >= 5:
        times.append(1e3*a.elapsed_time(b))
print(os.environ.get("NBS_B200_LIBRARY", "default"), precision, name, "energy" if want_energy else "forces",
      "eval_us %.1f (min %.1f)" % (np.mean(times), np.min(times)), " ".join(f"{k}={v:.1f}" for k, v in acc.items()))

"""A few device-resident evaluations of one workload (for ncu launch lists / captures)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
s = systems.make_system(name)
kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
kernel.initialize(s.system, s.force)
n = s.force.getNumParticles()
pos = torch.tensor(s.positions, dtype=torch.float64, device="cuda")
frc = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
lam = np.ones((s.force.getNumSlices(), 2))
for _ in range(reps):
    e = kernel.execute_device(pos.data_ptr(), s.box, frc.data_ptr(), lam, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("launches", kernel.getLaunchCount(), "energy checksum", float(np.abs(e).sum()))

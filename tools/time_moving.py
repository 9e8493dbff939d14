"""Whole-evaluation time along a ballistic trajectory (every atom its own thermal velocity), device-resident
positions, with the neighbour list re-used (default policy) and rebuilt on every evaluation.
usage: time_moving.py [config] [steps] [sigma_nm_per_step] [skin_nm]"""
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
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0009      # nm per step and component: ~300 K, 12 amu, 2 fs
skin = float(sys.argv[4]) if len(sys.argv) > 4 else -1.0
s = systems.make_system(name)
n = s.force.getNumParticles()
rng = np.random.default_rng(1)
vel = torch.tensor(rng.normal(scale=sigma, size=(n, 3)), dtype=torch.float64, device="cuda")
pos0 = torch.tensor(s.positions, dtype=torch.float64, device="cuda")
pos = pos0.clone()
frc = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
lam = np.ones((s.force.getNumSlices(), 2))
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
for label, flags in (("reuse", 0), ("rebuild", nbs.abi.NBS_FLAG_NO_LIST_REUSE)):
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=flags))
    kernel.initialize(s.system, s.force)
    if skin >= 0 and flags == 0:
        kernel.setListSkin(skin)
    times = []
    for t in range(steps + 10):
        pos.copy_(pos0 + t*vel)
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        kernel.execute_device(pos.data_ptr(), s.box, frc.data_ptr(), lam)
        b.record()
        torch.cuda.synchronize()
        if t >= 10:
            times.append(1e3*a.elapsed_time(b))
    st = kernel.getListStats()
    times = np.array(times)
    print(f"{name} {label} skin={st['skin']:.3f} sigma={sigma}: mean {times.mean():.1f} us  median {np.median(times):.1f}  min {times.min():.1f}  "
          f"max {times.max():.1f}  evals {st['evaluations']} builds {st['builds']} redone {st['redone']}  nlist {kernel.getNlistStats()[:5]}")

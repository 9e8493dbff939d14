"""Developer diagnostic (run on a GPU box through gpurun): CUDA path vs the CPU oracle, verbose."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
from oracle import oracle  # noqa: E402  (diagnostic tool: the oracle is the checker)


def compare(name, kind="port", lam_mode="mixed", check_pairs=True):
    s = systems.make_system(name)
    nsl = s.force.getNumSlices()
    lam = np.ones((nsl, 2))
    if lam_mode == "mixed":
        rng = np.random.default_rng(5)
        lam = rng.uniform(0.2, 1.0, size=(nsl, 2))
    platform = nbs.Platform(flags=nbs.abi.NBS_FLAG_PROFILE)
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    desc = kernel.desc
    for direct, recip in ((True, True), (True, False), (False, True)):
        forces = np.zeros((s.force.getNumParticles(), 3))
        t0 = time.time()
        e_gpu = kernel._evaluate(s.positions, s.box, lam, np.zeros(0), direct, recip, forces)
        t_gpu = time.time()-t0
        ref = oracle.evaluate(desc, s.positions, s.box, lam, None, direct, recip, kind=kind, want_pairs=False)
        frms = np.sqrt(((forces-ref.forces)**2).sum()/(ref.forces**2).sum())
        scale = np.maximum(np.abs(ref.slice_energies), 1.0)
        erel = np.abs(e_gpu-ref.slice_energies)/np.maximum(np.abs(ref.slice_energies), 1e-300)
        print(f"[{name}] direct={direct} recip={recip}: force relRMS {frms:.3e}  max |dE|/max(|E|,1) {np.max(np.abs(e_gpu-ref.slice_energies)/scale):.3e}"
              f"  max rel {np.max(np.where(np.abs(ref.slice_energies) > 1e-3, erel, 0)):.3e}  gpu wall {t_gpu*1e3:.1f} ms")
        print("   |dE| per slice/term:", np.array2string(np.abs(e_gpu-ref.slice_energies).ravel(), precision=2))
        print("   E ref             :", np.array2string(ref.slice_energies.ravel(), precision=4))
        if direct and recip:
            print("   E gpu", np.array2string(e_gpu.ravel(), precision=6))
            print("   E ref", np.array2string(ref.slice_energies.ravel(), precision=6))
            print("   kernel times (ms):", ", ".join(f"{n}={t:.3f}" for n, t in kernel.getKernelTimes()))
            print("   nlist stats:", kernel.getNlistStats())
        if direct and check_pairs:
            count, h, _ = kernel.getPairSet(with_pairs=False)
            print(f"   pairs gpu {count} ref {ref.pair_count}  hash equal: {h == ref.pair_hash}")
    return kernel


if __name__ == "__main__":
    names = sys.argv[1:] or ["C1", "C2"]
    for n in names:
        compare(n)

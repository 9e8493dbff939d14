"""Per-slice energy and force errors of the CUDA path against the CPU oracle for one workload (developer diagnostic)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
from oracle import oracle  # noqa: E402  (diagnostic: the oracle is the checker)
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
s = systems.make_system(name)
precision = sys.argv[2] if len(sys.argv) > 2 else "mixed"          # single | mixed | double
k = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(properties={"Precision": precision}))
k.initialize(s.system, s.force)
lam = np.ones((s.force.getNumSlices(), 2))
for direct, recip in ((True, False), (False, True), (True, True)):
    f = np.zeros_like(s.positions)
    e = k._evaluate(s.positions, s.box, lam, np.zeros(0), direct, recip, f)
    r = oracle.evaluate(k.desc, s.positions, s.box, lam, None, direct, recip, kind="port")
    err = np.abs(e - r.slice_energies)/np.maximum(np.abs(r.slice_energies), 1.0)
    print(name, "direct" if direct else "", "recip" if recip else "", "force relRMS %.2e" % np.sqrt(((f-r.forces)**2).sum()/(r.forces**2).sum()),
          "max energy err %.2e" % err.max(), "at", np.unravel_index(err.argmax(), err.shape))
    print(np.array2string(err, precision=1))
    print(np.array2string(e, precision=9), np.array2string(r.slice_energies, precision=9))

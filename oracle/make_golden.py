"""Generates tests/golden/*.npz from the REFERENCE backend (oracle/_ref: the reference's own unmodified
Reference-platform TUs, compiled where they lie under /root/reference).  Run here, where the
reference tree exists; the fixtures travel to the GPU box.

    python oracle/make_golden.py

Each fixture holds the inputs' identity (config name + lambdas; positions are regenerated
deterministically by systems.make_system and their checksum is stored) and the reference outputs:
forces, slice energies [nSl][2], interacting-pair count and hash.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
from oracle import oracle  # noqa: E402

CASES = [("C1", 11), ("C2", 12)]
# beyond the BASELINE configurations: plain Ewald, LJPME, triclinic boxes (systems.make_variant); global parameter
# values for the T1 systems' offset parameter travel in the fixture
VARIANT_CASES = [("C1_ewald", 13), ("C1_ljpme", 14), ("T1_pme", 15), ("T1_ljpme", 16)]


def lambdas_for(nsl, seed):
    return np.random.default_rng(seed).uniform(0.2, 1.0, size=(nsl, 2))


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, seed in CASES + VARIANT_CASES:
        s = systems.make_variant(name) if name in systems.VARIANTS else systems.make_system(name)
        desc = nbs.build_desc(s.system, s.force, legal_grid=True)
        lam = lambdas_for(s.force.getNumSlices(), seed)
        gv = np.full(max(s.force.getNumGlobalParameters(), 1), 0.45) if name in systems.VARIANTS else None
        data = {"lambdas": lam, "positions_checksum": np.array([s.positions.sum(), (s.positions**2).sum()])}
        if gv is not None:
            data["global_values"] = gv
        for tag, (direct, recip) in {"full": (True, True), "direct": (True, False), "recip": (False, True)}.items():
            r = oracle.evaluate(desc, s.positions, s.box, lam, gv, direct, recip, kind="reference")
            data[f"{tag}_energies"] = r.slice_energies
            data[f"{tag}_forces"] = r.forces.astype(np.float64)
            if direct:
                data["pair_count"] = np.array([r.pair_count], dtype=np.int64)
                data["pair_hash"] = np.array([r.pair_hash], dtype=np.uint64)
        np.savez_compressed(os.path.join(out, f"{name}_reference.npz"), **data)
        print("wrote", name, data["pair_count"], data["full_energies"].ravel()[:4])


if __name__ == "__main__":
    main()

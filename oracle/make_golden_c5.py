"""Generates tests/golden/C5_reference.npz: the STMV-size configuration (1,066,628 atoms, 180^3 grid) evaluated ONCE by
the reference's own compiled TUs (oracle/_ref) -- about a minute of CPU for the arithmetic plus the restated neighbour
list.  The full force array would be 25 MB, so the fixture keeps
  * the 3 x 2 slice energies of the full, direct-only and reciprocal-only evaluations,
  * the interacting-pair count and hash,
  * the forces of a fixed sample of 4,096 atoms (every evaluation),
  * sum |F|^2 over all atoms and the per-component force sums (every evaluation),
which is what the 1-GPU parity test and every multi-GPU bench line compare against.

    python oracle/make_golden_c5.py
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
from oracle import oracle  # noqa: E402

SAMPLE = 4096


def sample_indices(n):
    return np.sort(np.random.default_rng(20261018).choice(n, size=SAMPLE, replace=False))


def main():
    s = systems.make_system("C5")
    n = s.force.getNumParticles()
    desc = nbs.build_desc(s.system, s.force, legal_grid=True)
    lam = np.random.default_rng(15).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    idx = sample_indices(n)
    data = {"lambdas": lam, "sample": idx, "positions_checksum": np.array([s.positions.sum(), (s.positions**2).sum()])}
    for tag, (direct, recip) in {"full": (True, True), "direct": (True, False), "recip": (False, True)}.items():
        t0 = time.time()
        r = oracle.evaluate(desc, s.positions, s.box, lam, None, direct, recip, kind="reference", want_pairs=False)
        print(tag, "took %.1f s" % (time.time() - t0), r.timings, flush=True)
        data[f"{tag}_energies"] = r.slice_energies
        data[f"{tag}_forces_sample"] = r.forces[idx].astype(np.float64)
        data[f"{tag}_force_sumsq"] = np.array([(r.forces**2).sum()])
        data[f"{tag}_force_sum"] = r.forces.sum(axis=0)
        if direct:
            data["pair_count"] = np.array([r.pair_count], dtype=np.int64)
            data["pair_hash"] = np.array([r.pair_hash], dtype=np.uint64)
    out = os.path.join(ROOT, "tests", "golden", "C5_reference.npz")
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes", data["pair_count"], data["full_energies"].ravel())


if __name__ == "__main__":
    main()

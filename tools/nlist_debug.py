"""Neighbour-list statistics of a random (optionally triclinic) test system: blocks, list entries, exclusion-list
entries, capacities after growth.  Usage: nlist_debug.py seed nsub n L tiltB tiltCx tiltCy"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
from test_oracle_fixtures import random_system

seed, nsub, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
L = float(sys.argv[4])
tilt = tuple(float(v) for v in sys.argv[5:8])
rng = np.random.default_rng(seed)
system, force, positions = random_system(nbs, rng, n=n, nsub=nsub, L=L, tilt=tilt if any(tilt) else None)
ctx = nbs.Context(system, nbs.Platform())
ctx.setPositions(positions)
try:
    ctx.getState(getEnergy=True, getForces=True)
    print("ok", dict(zip(("blocks", "entries", "tiles", "tile_pairs", "x_entries", "capJ", "columns", "bins"),
                         ctx.impls[0].kernel.getNlistStats())))
except Exception as e:
    print("FAILED", e)

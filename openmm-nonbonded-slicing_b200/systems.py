"""Deterministic synthetic systems for the BASELINE.json configurations.

Counter-based RNG (SplitMix64) so that the same (config, seed) gives the same system everywhere.
Water is rigid 3-site TIP3P geometry on a jittered simple-cubic lattice; solute subsets are bonded
chains (three atoms per lattice site, consecutive atoms bonded) with exceptions from the
1-2/1-3/1-4 rule (createExceptionsFromBonds(bonds, 1/1.2, 0.5), as tests/TestSlicedNonbondedForce.h:149
does).  Boxes are cubic and PME grids have nx == ny == nz; alpha and the grid are set
explicitly.  After generation every pair is moved out of the guard band
|r^2 - r_c^2| < 1e-6 nm^2 so that the interacting-pair set is well defined.
"""
import ctypes as C
import math
import os

import numpy as np

from .api import SlicedNonbondedForce, System

ALPHA = 2.628261          # sqrt(-ln(2*5e-4))/1.0
CUTOFF = 1.0
GUARD_BAND = 1e-6

CONFIGS = {
    # name: waters, solute block sizes (subset 0..k-1; solvent is the last subset), box, grid, dispersion, net charge
    "C1": dict(waters=216, solute=[], solute_waters=8, box=2.0, grid=18, dispersion=True, seed=1235,
               description="TIP3P water box, 648 atoms, 2 subsets solute(8 waters)/solvent, PME 18^3"),
    "C2": dict(waters=2500, solute=[30], box=4.22, grid=36, dispersion=True, seed=1236, net_charge=[1.0],
               description="alchemical solvation: 30-atom ligand in 2,500 waters (7,530 atoms), 2 subsets, PME 36^3"),
    "C3": dict(waters=7023, solute=[2450, 39], box=6.2, grid=64, dispersion=False, seed=1237, net_charge=[-3.0, 1.0],
               description="DHFR-size: 23,558 atoms, 3 subsets protein/ligand/solvent, PME 64^3, all slice energies"),
    "C4": dict(waters=25750, solute=[6000, 6000, 2974], box=9.73, grid=80, dispersion=False, seed=1238,
               net_charge=[2.0, -2.0, 0.0],
               description="ApoA1-size: 92,224 atoms, 4 subsets, PME 80^3, dE/dlambda for every slice and term"),
    "C5": dict(waters=319988, solute=[106664], box=21.68, grid=180, dispersion=False, seed=1239, net_charge=[0.0],
               description="STMV-size: 1,066,628 atoms, 2 subsets, PME 180^3"),
}


def splitmix64(seed, stream, n):
    """n uniform doubles in [0, 1) from counter-based SplitMix64 (seed, stream) -- stateless."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed)*np.uint64(0x9E3779B97F4A7C15) + np.uint64(stream)*np.uint64(0xD1B54A32D192ED03)
        x = base + (np.arange(1, n+1, dtype=np.uint64))*np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30)))*np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27)))*np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64)*(1.0/9007199254740992.0)


def _snake_sites(M, count):
    """First `count` sites of an M^3 lattice in boustrophedon order (consecutive sites adjacent)."""
    k = np.arange(count)
    iz = k//(M*M)
    r = k - iz*M*M
    iy = r//M
    ix = r - iy*M
    iy = np.where(iz % 2 == 1, M-1-iy, iy)
    ix = np.where((r//M) % 2 == 1, M-1-ix, ix)
    return np.stack([ix, iy, iz], axis=1)


def _random_rotations(u):
    """Uniform random rotation matrices from 3 uniforms per row (Shoemake quaternions)."""
    u1, u2, u3 = u[:, 0], u[:, 1], u[:, 2]
    q = np.stack([np.sqrt(1-u1)*np.sin(2*np.pi*u2), np.sqrt(1-u1)*np.cos(2*np.pi*u2),
                  np.sqrt(u1)*np.sin(2*np.pi*u3), np.sqrt(u1)*np.cos(2*np.pi*u3)], axis=1)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(u), 3, 3))
    R[:, 0, 0] = 1-2*(y*y+z*z); R[:, 0, 1] = 2*(x*y-z*w); R[:, 0, 2] = 2*(x*z+y*w)
    R[:, 1, 0] = 2*(x*y+z*w); R[:, 1, 1] = 1-2*(x*x+z*z); R[:, 1, 2] = 2*(y*z-x*w)
    R[:, 2, 0] = 2*(x*z-y*w); R[:, 2, 1] = 2*(y*z+x*w); R[:, 2, 2] = 1-2*(x*x+y*y)
    return R


_tools = None


def _band_pairs(positions, L, cutoff, band):
    global _tools
    if _tools is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libnbs_hosttools.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `make -C {os.path.dirname(path)}`")
        _tools = C.CDLL(path)
        _tools.nbs_tools_band_pairs.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double,
                                                C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    lengths = np.array([L, L, L], dtype=np.float64)
    cap = 1 << 18
    pairs = np.zeros((cap, 2), dtype=np.int32)
    count = C.c_int64()
    _tools.nbs_tools_band_pairs(pos.shape[0], pos.ctypes.data_as(C.POINTER(C.c_double)), lengths.ctypes.data_as(C.POINTER(C.c_double)),
                                cutoff, band, cap, pairs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(count))
    assert count.value <= cap
    return pairs[:count.value]


class SyntheticSystem:
    """A generated configuration: ``system``, ``force`` (SlicedNonbondedForce), ``positions``, ``box``."""

    def __init__(self, name, system, force, positions, box, description, lambda_names):
        self.name, self.system, self.force, self.positions, self.box = name, system, force, positions, box
        self.description = description
        self.lambda_names = lambda_names


def _bulk_exceptions_from_chain(first, count, charges, sigmas, epsilons, coulomb14Scale, lj14Scale):
    """Exceptions of a linear chain of `count` atoms starting at `first` (1-2, 1-3 excluded, 1-4 scaled):
    what createExceptionsFromBonds gives for consecutive-atom bonds, vectorised."""
    out = []
    for sep in (1, 2, 3):
        if count <= sep:
            continue
        i = np.arange(first, first+count-sep)
        j = i + sep
        if sep < 3:
            qq, sg, ep = np.zeros(len(i)), np.ones(len(i)), np.zeros(len(i))
        else:
            qq = coulomb14Scale*charges[i]*charges[j]
            sg = 0.5*(sigmas[i]+sigmas[j])
            ep = lj14Scale*np.sqrt(epsilons[i]*epsilons[j])
        out.append(np.stack([i, j, qq, sg, ep], axis=1))
    return np.concatenate(out) if out else np.zeros((0, 5))


def make_system(name, seed=None, derivatives=True, guard_band=GUARD_BAND):
    cfg = CONFIGS[name]
    seed = cfg["seed"] if seed is None else seed
    L = cfg["box"]
    waters = cfg["waters"]
    solute = cfg["solute"]
    n_solute = sum(solute)
    solute_sites = [(s+2)//3 for s in solute]
    n_sites = waters + sum(solute_sites)
    M = int(math.ceil(n_sites**(1.0/3.0) - 1e-9))
    while M**3 < n_sites:
        M += 1
    spacing = L/M
    sites = (_snake_sites(M, n_sites) + 0.5)*spacing
    n = n_solute + 3*waters
    positions = np.zeros((n, 3))
    charges, sigmas, epsilons = np.zeros(n), np.zeros(n), np.zeros(n)
    subsets = np.zeros(n, dtype=np.int32)
    exceptions = []

    # solute blocks: subset b occupies a contiguous run of sites, three atoms per site along a chain
    atom, site = 0, 0
    for b, size in enumerate(solute):
        k = np.arange(size)
        s = site + k//3
        # direction of travel along the snake so that consecutive atoms stay ~0.1 nm apart
        nxt = sites[np.minimum(s+1, n_sites-1)] - sites[s]
        prv = sites[s] - sites[np.maximum(s-1, 0)]
        direction = np.where((np.abs(nxt).sum(axis=1) > 0)[:, None], nxt, prv)
        direction = direction/np.maximum(np.linalg.norm(direction, axis=1), 1e-12)[:, None]
        positions[atom:atom+size] = sites[s] + direction*((k % 3)-1)[:, None]*(spacing/3.0)
        u = splitmix64(seed, 10+b, 3*size).reshape(size, 3)
        q = u[:, 0] - 0.5
        net = cfg.get("net_charge", [0.0]*len(solute))[b]
        q += (net - q.sum())/size
        charges[atom:atom+size] = q
        sigmas[atom:atom+size] = 0.25 + 0.1*u[:, 1]
        epsilons[atom:atom+size] = 0.2 + 0.6*u[:, 2]
        subsets[atom:atom+size] = b
        exceptions.append(_bulk_exceptions_from_chain(atom, size, charges, sigmas, epsilons, 1/1.2, 0.5))
        atom += size
        site += solute_sites[b]

    # waters
    solvent_subset = len(solute)
    w = np.arange(waters)
    theta = math.radians(104.52)
    rOH = 0.09572
    local = np.array([[0.0, 0.0, 0.0],
                      [rOH*math.sin(theta/2), rOH*math.cos(theta/2), 0.0],
                      [-rOH*math.sin(theta/2), rOH*math.cos(theta/2), 0.0]])
    R = _random_rotations(splitmix64(seed, 1, 3*waters).reshape(waters, 3))
    wpos = sites[site + w][:, None, :] + np.einsum("wij,aj->wai", R, local)
    positions[atom:] = wpos.reshape(-1, 3)
    charges[atom:] = np.tile([-0.834, 0.417, 0.417], waters)
    sigmas[atom:] = np.tile([0.315075, 1.0, 1.0], waters)
    epsilons[atom:] = np.tile([0.635968, 0.0, 0.0], waters)
    subsets[atom:] = solvent_subset
    o = atom + 3*w
    zeros, ones = np.zeros(waters), np.ones(waters)
    for a, b in ((0, 1), (0, 2), (1, 2)):
        exceptions.append(np.stack([o+a, o+b, zeros, ones, zeros], axis=1))
    if "solute_waters" in cfg:         # C1: the "solute" subset is the first few waters
        subsets[:] = 1
        subsets[atom:atom+3*cfg["solute_waters"]] = 0
        solvent_subset = 1

    # jitter every atom a little, then move pairs out of the guard band around the cutoff
    positions += (splitmix64(seed, 2, 3*n).reshape(n, 3) - 0.5)*0.04*np.array([1.0, 1.0, 1.0])
    # keep water rigid: the jitter above is per atom, so re-impose the geometry from the jittered O
    opos = positions[o]
    positions[atom:] = (opos[:, None, :] + np.einsum("wij,aj->wai", R, local)).reshape(-1, 3)
    for iteration in range(50):
        bad = _band_pairs(positions, L, CUTOFF, guard_band)
        if len(bad) == 0:
            break
        movers = np.unique(bad[:, 1])
        # move whole molecules for water (rigid), single atoms for solute
        shift = (splitmix64(seed, 100+iteration, 3*len(movers)).reshape(-1, 3) - 0.5)*2e-3
        for m, d in zip(movers, shift):
            if m >= atom:
                mol = atom + 3*((m-atom)//3)
                positions[mol:mol+3] += d
            else:
                positions[m] += d
    else:
        raise RuntimeError("could not clear the cutoff guard band")

    system = System()
    system.setDefaultPeriodicBoxVectors([L, 0, 0], [0, L, 0], [0, 0, L])
    system._masses = [1.0]*n
    force = SlicedNonbondedForce(len(solute)+1 if "solute_waters" not in cfg else 2)
    force.setNonbondedMethod(SlicedNonbondedForce.PME)
    force.setCutoffDistance(CUTOFF)
    force.setPMEParameters(ALPHA, cfg["grid"], cfg["grid"], cfg["grid"])
    force.setUseDispersionCorrection(cfg["dispersion"])
    force._particles = np.stack([charges, sigmas, epsilons], axis=1).tolist()
    exc = np.concatenate(exceptions)
    force._exceptions = [[int(e[0]), int(e[1]), e[2], e[3], e[4]] for e in exc.tolist()]
    force._exceptionMap = {(min(e[0], e[1]), max(e[0], e[1])): k for k, e in enumerate(force._exceptions)}
    force._subsets = {int(i): int(s) for i, s in enumerate(subsets) if s != 0}
    system.addForce(force)

    # scaling parameters: one per (slice, term), all with derivatives -> every slice energy observable
    lambda_names = []
    nS = force.getNumSubsets()
    if name == "C2":
        force.addGlobalParameter("lambda_elec", 1.0)
        force.addGlobalParameter("lambda_vdw", 1.0)
        force.addScalingParameter("lambda_elec", 0, 1, True, False)
        force.addScalingParameter("lambda_vdw", 0, 1, False, True)
        lambda_names = ["lambda_elec", "lambda_vdw"]
    else:
        for i in range(nS):
            for j in range(i, nS):
                for term, (c, l) in (("c", (True, False)), ("v", (False, True))):
                    pname = f"lam_{term}_{i}{j}"
                    force.addGlobalParameter(pname, 1.0)
                    force.addScalingParameter(pname, i, j, c, l)
                    lambda_names.append(pname)
    if derivatives:
        for pname in lambda_names:
            force.addEnergyParameterDerivative(pname)
    box = np.diag([L, L, L])
    return SyntheticSystem(name, system, force, positions, box, cfg["description"], lambda_names)


# ---------------------------------------------------------------------------------------------
# Small variants for the methods and box shapes beyond the BASELINE configurations: what the golden
# fixtures tests/golden/{C1_ewald, C1_ljpme, T1_pme, T1_ljpme}_reference.npz are generated from
# (oracle/make_golden.py).  Everything comes from SplitMix64, so the systems are identical everywhere.
# ---------------------------------------------------------------------------------------------
VARIANTS = {
    "C1_ewald": "C1 with the method switched to plain Ewald (kmax from calcEwaldParameters at tolerance 5e-4)",
    "C1_ljpme": "C1 with LJPME, dispersion grid 12^3 at alpha_d = 2.4",
    "T1_pme": "300 atoms on a sheared lattice, triclinic box (tilt 0.3, -0.2, 0.4), 3 subsets, PME 20^3, 1-4 exceptions, offsets",
    "T1_ljpme": "the same system with LJPME (dispersion grid 12^3)",
}


def _triclinic_test_system(method, seed=4242, n=300, nsub=3, L=2.6, tilt=(0.3, -0.2, 0.4), grid=20):
    system = System()
    force = SlicedNonbondedForce(nsub)
    force.setNonbondedMethod(method)
    force.setCutoffDistance(CUTOFF)
    force.setPMEParameters(2.8, grid, grid, grid)
    box = np.array([[L, 0, 0], [tilt[0]*L, L, 0], [tilt[1]*L, tilt[2]*L, L]], dtype=float)
    system.setDefaultPeriodicBoxVectors(*box)
    side = int(math.ceil(n**(1/3)))
    sites = np.array([(i, j, k) for i in range(side) for j in range(side) for k in range(side)][:n], dtype=float)
    jitter = (splitmix64(seed, 1, 3*n).reshape(n, 3) - 0.5)*0.1
    images = np.floor(splitmix64(seed, 2, 3*n).reshape(n, 3)*5) - 2        # unwrapped input: whole box vectors added
    positions = ((sites + 0.5)/side + images) @ box + jitter
    charges = (splitmix64(seed, 3, n) - 0.5)*1.6
    sigmas = 0.15 + 0.15*splitmix64(seed, 4, n)
    epsilons = 0.1 + 0.9*splitmix64(seed, 5, n)
    subsets = np.floor(splitmix64(seed, 6, n)*nsub).astype(int)
    for i in range(n):
        system.addParticle(1.0)
        force.addParticle(float(charges[i]), float(sigmas[i]), float(epsilons[i]))
        force.setParticleSubset(i, int(subsets[i]))
    force.createExceptionsFromBonds([(i, i+1) for i in range(0, n-1) if i % 5 != 4], 1/1.2, 0.5)
    force.addGlobalParameter("off", 0.3)
    force.addParticleParameterOffset("off", 3, 0.5, 0.01, 0.2)
    force.addExceptionParameterOffset("off", 2, 0.2, 0.01, 0.1)
    force.setExceptionsUsePeriodicBoundaryConditions(True)
    system.addForce(force)
    return system, force, positions, box


def make_variant(name):
    """See VARIANTS.  Returns a SyntheticSystem; ``box`` is the 3x3 matrix of box vectors (rows)."""
    if name == "C1_ewald":
        s = make_system("C1")
        s.force.setNonbondedMethod(SlicedNonbondedForce.Ewald)
    elif name == "C1_ljpme":
        s = make_system("C1")
        s.force.setNonbondedMethod(SlicedNonbondedForce.LJPME)
        s.force.setLJPMEParameters(2.4, 12, 12, 12)
    elif name in ("T1_pme", "T1_ljpme"):
        method = SlicedNonbondedForce.PME if name == "T1_pme" else SlicedNonbondedForce.LJPME
        system, force, positions, box = _triclinic_test_system(method)
        if name == "T1_ljpme":
            force.setLJPMEParameters(2.4, 12, 12, 12)
        s = SyntheticSystem(name, system, force, positions, box, VARIANTS[name], [])
    else:
        raise KeyError(name)
    s.name, s.description = name, VARIANTS[name]
    return s

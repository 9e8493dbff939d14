"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against
the CPU oracle and the committed reference fixtures.

Tolerances (BASELINE.json north_star):
  * interacting-pair set and exclusion set: bit-exact (count + order-independent 64-bit hash, and the
    sorted pair lists themselves on the smaller systems);
  * forces: relative RMS error sqrt(sum |F-Fref|^2 / sum |Fref|^2) <= 1e-5;
  * per-slice energies and dE/dlambda: |E - Eref| <= 1e-5 * max(|Eref|, 1) for every slice and term, in
    direct-only, reciprocal-only and full evaluations (the floor of 1 kJ/mol is the reference tests' own
    assertion semantics, AssertionUtilities.h:20-27).  Slice energies are small differences of large
    sums (protein-ligand Coulomb in C3: +12.98 - 9.36 = 3.62 kJ/mol out of ~10^4 of pair terms), which
    is why every energy term is evaluated in double precision on the device ("Precision").
"""
import importlib
import os

import numpy as np
import pytest

from helpers import assert_equal_tol, assert_equal_vec, force_rel_rms, TOL
from test_oracle_fixtures import random_system

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F_TOL = 1e-5
E_TOL = 1e-5


@pytest.fixture(scope="module")
def systems():
    return importlib.import_module("openmm-nonbonded-slicing_b200.systems")


@pytest.fixture(scope="module")
def platform(nbs):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    nbs.abi.load_library()
    return nbs.Platform()


def check_energies(found, expected, scale=None):
    scale = np.maximum(np.abs(expected) if scale is None else scale, 1.0)
    err = np.abs(found-expected)/scale
    assert err.max() <= E_TOL, f"slice energy error {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}\n{found}\n{expected}"


def three_way(kernel, desc, s_positions, box, lam, oracle_eval):
    """direct-only, reciprocal-only and full evaluations against the oracle (or fixture) callback."""
    parts = {}
    n = s_positions.shape[0]
    for tag, (direct, recip) in {"direct": (True, False), "recip": (False, True), "full": (True, True)}.items():
        forces = np.zeros((n, 3))
        e = kernel._evaluate(s_positions, box, lam, np.zeros(0), direct, recip, forces)
        ref_e, ref_f, pair_count, pair_hash = oracle_eval(tag, direct, recip)
        parts[tag] = ref_e
        assert force_rel_rms(forces, ref_f) <= F_TOL, tag
        check_energies(e, ref_e)
        if direct:
            count, h, _ = kernel.getPairSet(with_pairs=False)
            assert (count, h) == (pair_count, pair_hash), f"pair set differs ({count} vs {pair_count})"


@pytest.mark.parametrize("name", ["C1", "C2", "C1_ewald", "C1_ljpme", "T1_pme", "T1_ljpme"])
def test_reference_fixture(nbs, platform, systems, name):
    """CUDA path vs the outputs of the reference's own compiled TUs (tests/golden, oracle/make_golden.py): the BASELINE
    configurations C1 / C2 and the variants beyond them -- plain Ewald, LJPME, a triclinic box (systems.VARIANTS)."""
    g = np.load(os.path.join(GOLDEN, f"{name}_reference.npz"))
    s = systems.make_variant(name) if name in systems.VARIANTS else systems.make_system(name)
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    if "global_values" in g.files:
        kernel._push_parameters(g["lambdas"], np.ascontiguousarray(g["global_values"], dtype=np.float64))

    def fixture(tag, direct, recip):
        return g[f"{tag}_energies"], g[f"{tag}_forces"], int(g["pair_count"][0]), int(g["pair_hash"][0])
    three_way(kernel, kernel.desc, s.positions, s.box, g["lambdas"], fixture)


@pytest.mark.parametrize("flags", ["NBS_FLAG_LINE_FFT", "NBS_FLAG_SORTED_PME", "NBS_FLAG_NO_GRAPH"])
def test_alternative_paths(nbs, systems, flags):
    """Paths that small systems do not take by default, forced on a small system and held to the same fixture:
    the line-at-a-time FFT kernels (planes beyond shared memory), PME from the cell-sorted records (large
    systems), plain stream launches instead of the captured CUDA graph."""
    g = np.load(os.path.join(GOLDEN, "C2_reference.npz"))
    s = systems.make_system("C2")
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=getattr(nbs.abi, flags)))
    kernel.initialize(s.system, s.force)

    def fixture(tag, direct, recip):
        return g[f"{tag}_energies"], g[f"{tag}_forces"], int(g["pair_count"][0]), int(g["pair_hash"][0])
    three_way(kernel, kernel.desc, s.positions, s.box, g["lambdas"], fixture)


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_baseline_configs_vs_oracle(nbs, platform, systems, oracle, name):
    s = systems.make_system(name)
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    lam = np.random.default_rng(3).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))

    def run(tag, direct, recip):
        r = oracle.evaluate(kernel.desc, s.positions, s.box, lam, None, direct, recip, kind="port")
        return r.slice_energies, r.forces, r.pair_count, r.pair_hash
    three_way(kernel, kernel.desc, s.positions, s.box, lam, run)


@pytest.mark.parametrize("method,n,switch", [("NoCutoff", 150, False), ("NoCutoff", 37, False), ("CutoffNonPeriodic", 400, False),
                                              ("CutoffNonPeriodic", 333, True)])
def test_nonperiodic_methods(nbs, platform, oracle, method, n, switch):
    """NoCutoff and CutoffNonPeriodic (reaction field, optional switching function): a free cluster anywhere in
    space, no box; 1-4 exceptions, offsets, three lambda settings; forces, slice energies, derivatives and the
    interacting-pair set against the oracle (ReferenceSlicedLJCoulombIxn.cpp:571-631)."""
    rng = np.random.default_rng(n)
    nsub = 3
    system = nbs.System()
    force = nbs.SlicedNonbondedForce(nsub)
    force.setNonbondedMethod(getattr(force, method))
    force.setCutoffDistance(1.0)
    if switch:
        force.setUseSwitchingFunction(True)
        force.setSwitchingDistance(0.8)
    side = int(np.ceil(n**(1/3)))
    sites = np.array([(i, j, k) for i in range(side) for j in range(side) for k in range(side)][:n], dtype=float)
    positions = sites*0.31 + rng.uniform(-0.05, 0.05, size=(n, 3)) + np.array([17.0, -250.0, 3.3])
    charges = rng.uniform(-0.8, 0.8, size=n)
    for i in range(n):
        system.addParticle(1.0)
        force.addParticle(charges[i], rng.uniform(0.15, 0.3), rng.uniform(0.1, 1.0))
        force.setParticleSubset(i, int(rng.integers(0, nsub)))
    force.createExceptionsFromBonds([(i, i+1) for i in range(0, n-1) if i % 5 != 4], 1/1.2, 0.5)
    force.addGlobalParameter("off", 0.3)
    force.addParticleParameterOffset("off", 3, 0.5, 0.01, 0.2)
    force.addExceptionParameterOffset("off", 2, 0.2, 0.01, 0.1)
    force.addGlobalParameter("lc", 0.7)
    force.addGlobalParameter("lv", 0.4)
    force.addScalingParameter("lc", 0, 1, True, False)
    force.addScalingParameter("lv", 0, 1, False, True)
    force.addEnergyParameterDerivative("lc")
    force.addEnergyParameterDerivative("lv")
    system.addForce(force)
    ctx = nbs.Context(system, platform)
    ref = nbs.Context(system, oracle.OraclePlatform("port"))
    for c in (ctx, ref):
        c.setPositions(positions)
    for lc, lv in ((0.7, 0.4), (0.0, 1.0), (1.0, 1.0)):
        for c in (ctx, ref):
            c.setParameter("lc", lc)
            c.setParameter("lv", lv)
        a = ctx.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        b = ref.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
        check_energies(ctx.impls[0].kernel.lastSliceEnergies, ref.impls[0].kernel.lastSliceEnergies)
        for name, value in b.getEnergyParameterDerivatives().items():
            assert_equal_tol(value, a.getEnergyParameterDerivatives()[name], E_TOL)
        r = ref.impls[0].kernel.lastResult
        count, h, _ = ctx.impls[0].kernel.getPairSet(with_pairs=False)
        if method == "NoCutoff":      # the reference loops over all pairs here (no neighbour list to compare with)
            excluded = {(min(e[0], e[1]), max(e[0], e[1])) for e in force._exceptions}
            assert count == n*(n-1)//2 - len(excluded)
        else:
            assert (count, h) == (r.pair_count, r.pair_hash)


def test_pair_and_exclusion_lists_exact(nbs, platform, systems, oracle):
    """The sorted pair list itself (not just its hash) and the exclusion set, on C1 and a random system."""
    s = systems.make_system("C1")
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    lam = np.ones((s.force.getNumSlices(), 2))
    kernel._evaluate(s.positions, s.box, lam, np.zeros(0), True, False, np.zeros((648, 3)))
    count, h, pairs = kernel.getPairSet(with_pairs=True)
    ref = oracle.evaluate(kernel.desc, s.positions, s.box, lam, None, True, False, kind="port", want_pairs=True)
    a = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    b = ref.pairs[np.lexsort((ref.pairs[:, 1], ref.pairs[:, 0]))]
    assert np.array_equal(a, b)
    assert h == int(nbs.abi.pair_hash(b[:, 0], b[:, 1]).sum(dtype=np.uint64))
    excl = kernel.getExclusionSet()
    expected = sorted({(min(e[0], e[1]), max(e[0], e[1])) for e in s.force._exceptions})
    assert [tuple(p) for p in excl.tolist()] == expected


@pytest.mark.parametrize("seed,nsub,grid,n", [(1, 3, (20, 20, 20), 300), (2, 1, (24, 18, 30), 333), (3, 4, (25, 21, 28), 97),
                                               (4, 2, (22, 26, 20), 1000), (5, 8, (20, 20, 20), 250)])
def test_random_systems(nbs, platform, oracle, seed, nsub, grid, n):
    """Ragged sizes (N not a multiple of 32), unwrapped coordinates, 1-4 exceptions, parameter offsets,
    net subset charges, odd / non-cubic grids with factors 3, 5, 7, 11, 13; checked through the Context API."""
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng, n=n, nsub=nsub, L=2.6 if n < 500 else 3.4, grid=grid)
    ctx = nbs.Context(system, platform)
    ref = nbs.Context(system, oracle.OraclePlatform("port"))
    for c in (ctx, ref):
        c.setPositions(positions)
        c.setParameter("off", 0.45)
    for lc, lv in ((0.7, 0.4), (0.0, 1.0), (1.0, 0.25)):
        for c in (ctx, ref):
            c.setParameter("lc", lc)
            c.setParameter("lv", lv)
        for groups in (0xFFFFFFFF,):
            a = ctx.getState(getEnergy=True, getForces=True, getParameterDerivatives=True, groups=groups)
            b = ref.getState(getEnergy=True, getForces=True, getParameterDerivatives=True, groups=groups)
            assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
            ea = ctx.impls[0].kernel.lastSliceEnergies
            eb = ref.impls[0].kernel.lastSliceEnergies
            k, r = ctx.impls[0].kernel, ref.impls[0].kernel.lastResult
            check_energies(ea, eb)
            for name, value in b.getEnergyParameterDerivatives().items():
                assert_equal_tol(value, a.getEnergyParameterDerivatives()[name], E_TOL)
            count, h, _ = k.getPairSet(with_pairs=False)
            assert (count, h) == (r.pair_count, r.pair_hash)
    # offsets change through setParameter only (no re-initialisation)
    for c in (ctx, ref):
        c.setParameter("off", -0.2)
    a = ctx.getState(getEnergy=True, getForces=True)
    b = ref.getState(getEnergy=True, getForces=True)
    assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
    assert_equal_tol(b.getPotentialEnergy(), a.getPotentialEnergy(), 1e-5)
    # updateParametersInContext
    q, sg, ep = force.getParticleParameters(5)
    force.setParticleParameters(5, q+0.3, sg, ep*0.5)
    for c in (ctx, ref):
        force.updateParametersInContext(c)
    a = ctx.getState(getEnergy=True, getForces=True)
    b = ref.getState(getEnergy=True, getForces=True)
    assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
    assert_equal_tol(b.getPotentialEnergy(), a.getPotentialEnergy(), 1e-5)


@pytest.mark.parametrize("seed,nsub,n,tol", [(31, 3, 300, 5e-4), (32, 1, 97, 1e-4), (33, 8, 1000, 5e-4)])
def test_ewald_random_systems(nbs, platform, oracle, seed, nsub, n, tol):
    """Plain Ewald (ReferenceSlicedLJCoulombIxn.cpp:256-358) on the device: erfc direct space and exclusion
    corrections as for PME, reciprocal sum over the half space of k vectors (csrc/k_ewald.cu); direct-only,
    reciprocal-only and full evaluations, derivatives and the pair set against the oracle."""
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng, n=n, nsub=nsub, L=2.6 if n < 500 else 3.4, method="Ewald")
    force.setEwaldErrorTolerance(tol)
    ctx = nbs.Context(system, platform)
    ref = nbs.Context(system, oracle.OraclePlatform("port"))
    for c in (ctx, ref):
        c.setPositions(positions)
        c.setParameter("off", 0.45)
    for lc, lv in ((0.7, 0.4), (0.0, 1.0)):
        for c in (ctx, ref):
            c.setParameter("lc", lc)
            c.setParameter("lv", lv)
        a = ctx.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        b = ref.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
        check_energies(ctx.impls[0].kernel.lastSliceEnergies, ref.impls[0].kernel.lastSliceEnergies)
        for name, value in b.getEnergyParameterDerivatives().items():
            assert_equal_tol(value, a.getEnergyParameterDerivatives()[name], E_TOL)
        count, h, _ = ctx.impls[0].kernel.getPairSet(with_pairs=False)
        r = ref.impls[0].kernel.lastResult
        assert (count, h) == (r.pair_count, r.pair_hash)
    kernel = ctx.impls[0].kernel
    lam = rng.uniform(0.2, 1.0, size=(force.getNumSlices(), 2))
    gv = np.array([0.45, 0.0, 1.0])

    def run(tag, direct, recip):
        r = oracle.evaluate(kernel.desc, positions, kernel_box, lam, gv, direct, recip, kind="port")
        return r.slice_energies, r.forces, r.pair_count, r.pair_hash
    kernel_box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)
    three_way(kernel, kernel.desc, positions, kernel_box, lam, run)


@pytest.mark.parametrize("seed,nsub,n,grid,dgrid", [(41, 3, 300, (20, 20, 20), (12, 12, 12)), (42, 1, 97, (24, 18, 30), (10, 14, 9)),
                                                     (43, 4, 1000, (22, 26, 22), (16, 15, 16)), (44, 3, 500, (22, 26, 20), (16, 15, 18))])
def test_ljpme_random_systems(nbs, platform, oracle, seed, nsub, n, grid, dgrid):
    """LJPME on the device against the reference's own compiled TUs (and, for grids with nx != nz, the port): direct space
    with the multiplicative C6 term taken out and the potential shift (ReferenceSlicedLJCoulombIxn.cpp:398-426),
    dispersion exclusion corrections (:487-504), self term (:211-212) and the second PME chain on the dispersion grid
    (ReferencePME.cpp:499-595, 814-871); direct-only, reciprocal-only, full; derivatives; pair set.
    Multi-subset cases keep nx == nz: the reference's gather indexes subset grids with sj*nz instead of sj*nx
    (ReferencePME.cpp:682), which only coincides with its own spreading when nx == nz."""
    kind = "reference" if grid[0] == grid[2] and dgrid[0] == dgrid[2] else "port"
    if not oracle.available(kind):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng, n=n, nsub=nsub, L=2.6 if n < 500 else 3.4, grid=grid, method="LJPME")
    force.setLJPMEParameters(2.4, *dgrid)
    ctx = nbs.Context(system, platform)
    ref = nbs.Context(system, oracle.OraclePlatform(kind))
    assert ctx.impls[0].kernel.getLJPMEParameters() == (2.4,)+tuple(dgrid)
    for c in (ctx, ref):
        c.setPositions(positions)
        c.setParameter("off", 0.45)
    for lc, lv in ((0.7, 0.4), (1.0, 0.0)):
        for c in (ctx, ref):
            c.setParameter("lc", lc)
            c.setParameter("lv", lv)
        a = ctx.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        b = ref.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
        check_energies(ctx.impls[0].kernel.lastSliceEnergies, ref.impls[0].kernel.lastSliceEnergies)
        for name, value in b.getEnergyParameterDerivatives().items():
            assert_equal_tol(value, a.getEnergyParameterDerivatives()[name], E_TOL)
        count, h, _ = ctx.impls[0].kernel.getPairSet(with_pairs=False)
        r = ref.impls[0].kernel.lastResult
        assert (count, h) == (r.pair_count, r.pair_hash)
    kernel = ctx.impls[0].kernel
    lam = rng.uniform(0.2, 1.0, size=(force.getNumSlices(), 2))
    gv = np.array([0.45, 0.0, 1.0])
    box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)

    def run(tag, direct, recip):
        r = oracle.evaluate(kernel.desc, positions, box, lam, gv, direct, recip, kind=kind)
        return r.slice_energies, r.forces, r.pair_count, r.pair_hash
    three_way(kernel, kernel.desc, positions, box, lam, run)


@pytest.mark.parametrize("seed,nsub,n,L,grid,tilt,method", [
    (61, 3, 300, 2.6, (20, 20, 20), (0.3, -0.2, 0.4), "PME"),
    (62, 2, 1000, 3.4, (24, 18, 30), (-0.5, 0.5, -0.5), "PME"),         # the extreme reduced form: images two boxes away in x
    (63, 4, 97, 2.2, (25, 21, 28), (0.5, 0.5, 0.5), "PME"),
    (64, 3, 333, 2.6, (20, 20, 20), (0.25, 0.1, -0.35), "CutoffPeriodic"),
    (65, 2, 2000, 4.4, (30, 30, 30), (-0.45, -0.3, 0.2), "PME"),
    (66, 3, 300, 2.6, (20, 20, 20), (0.3, -0.2, 0.4), "LJPME")])
def test_triclinic_random_systems(nbs, platform, oracle, seed, nsub, n, L, grid, tilt, method):
    """Triclinic boxes on the device: atoms wrapped into the rectangular brick [0,ax) x [0,by) x [0,cz) of the same
    lattice, images displaced by kx a + ky b + kz c in the list builder and the pair kernel, lattice fractions and
    the full reciprocal matrix in PME, OpenMM's minimum image for periodic exceptions; unwrapped input coordinates."""
    kind = "port"
    if method == "LJPME":
        if not oracle.available("reference"):
            pytest.skip("oracle/_ref not built")
        kind = "reference"
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng, n=n, nsub=nsub, L=L, grid=grid, method=method, tilt=tilt)
    if method == "LJPME":
        force.setLJPMEParameters(2.4, 12, 12, 12)
    force.setExceptionsUsePeriodicBoundaryConditions(seed % 2 == 1)
    ctx = nbs.Context(system, platform)
    ref = nbs.Context(system, oracle.OraclePlatform(kind))
    for c in (ctx, ref):
        c.setPositions(positions)
        c.setParameter("off", 0.45)
    for lc, lv in ((0.7, 0.4), (1.0, 0.0)):
        for c in (ctx, ref):
            c.setParameter("lc", lc)
            c.setParameter("lv", lv)
        a = ctx.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        b = ref.getState(getEnergy=True, getForces=True, getParameterDerivatives=True)
        count, h, _ = ctx.impls[0].kernel.getPairSet(with_pairs=False)
        r = ref.impls[0].kernel.lastResult
        assert (count, h) == (r.pair_count, r.pair_hash), f"pair set differs ({count} vs {r.pair_count})"
        assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
        check_energies(ctx.impls[0].kernel.lastSliceEnergies, ref.impls[0].kernel.lastSliceEnergies)
        for name, value in b.getEnergyParameterDerivatives().items():
            assert_equal_tol(value, a.getEnergyParameterDerivatives()[name], E_TOL)
    kernel = ctx.impls[0].kernel
    lam = rng.uniform(0.2, 1.0, size=(force.getNumSlices(), 2))
    gv = np.array([0.45, 0.0, 1.0])
    box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)

    def run(tag, direct, recip):
        r = oracle.evaluate(kernel.desc, positions, box, lam, gv, direct, recip, kind=kind)
        return r.slice_energies, r.forces, r.pair_count, r.pair_hash
    if method != "CutoffPeriodic":
        three_way(kernel, kernel.desc, positions, box, lam, run)


@pytest.mark.parametrize("method,flag,tilt", [("LJPME", "NBS_FLAG_FP32_ENERGY", None), ("Ewald", "NBS_FLAG_FP32_ENERGY", None),
                                              ("PME", "NBS_FLAG_SORTED_PME", (0.3, -0.2, 0.4)), ("PME", "NBS_FLAG_LINE_FFT", (0.3, -0.2, 0.4)),
                                              ("LJPME", "NBS_FLAG_NO_GRAPH", (-0.5, 0.5, -0.5)), ("Ewald", "NBS_FLAG_NO_GRAPH", None)])
def test_method_and_flag_combinations(nbs, oracle, method, flag, tilt):
    """The late additions on the paths they do not take by default: single-precision energies (looser energy
    tolerance: that is the plugin's "single" mode) with the LJPME / Ewald pair-kernel variants, a triclinic box through
    the sorted-PME and line-FFT paths, plain launches; and the forces-only evaluation (no slice energies requested)
    against the same oracle forces."""
    import torch
    rng = np.random.default_rng(91)
    system, force, positions = random_system(nbs, rng, n=400, nsub=3, L=2.8, grid=(24, 20, 24), method=method, tilt=tilt)
    if method == "LJPME":
        force.setLJPMEParameters(2.4, 14, 12, 14)
    fp32 = flag == "NBS_FLAG_FP32_ENERGY"
    ctx = nbs.Context(system, nbs.Platform(flags=getattr(nbs.abi, flag)))
    ref = nbs.Context(system, oracle.OraclePlatform("port"))
    for c in (ctx, ref):
        c.setPositions(positions)
        c.setParameter("lc", 0.6)
        c.setParameter("lv", 0.3)
    for _ in range(3):                      # plain launches, graph capture, graph replay
        a = ctx.getState(getEnergy=True, getForces=True)
    b = ref.getState(getEnergy=True, getForces=True)
    assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
    ea, eb = ctx.impls[0].kernel.lastSliceEnergies, ref.impls[0].kernel.lastSliceEnergies
    err = np.abs(ea-eb)/np.maximum(np.abs(eb), 1.0)
    assert err.max() <= (2e-4 if fp32 else E_TOL), err
    # forces only
    kernel = ctx.impls[0].kernel
    pos = torch.tensor(positions, dtype=torch.float64, device="cuda")
    frc = torch.zeros((400, 3), dtype=torch.float64, device="cuda")
    lam = np.ones((6, 2))
    lam[1] = [0.6, 0.3]
    box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)
    for _ in range(3):
        kernel.execute_device(pos.data_ptr(), box, frc.data_ptr(), lam, want_energies=False)
    torch.cuda.synchronize()
    assert force_rel_rms(frc.cpu().numpy(), b.getForces()) <= F_TOL


def test_non_reduced_box_is_refused(nbs, platform):
    """Box vectors outside OpenMM's reduced form are an error, not a silently different lattice."""
    system = nbs.System()
    system.setDefaultPeriodicBoxVectors([3, 0, 0], [2.0, 3, 0], [0, 0, 3])
    force = nbs.SlicedNonbondedForce(1)
    force.setNonbondedMethod(force.CutoffPeriodic)
    for k in range(2):
        system.addParticle(1.0)
        force.addParticle(1.0, 0.3, 0.1)
    system.addForce(force)
    ctx = nbs.Context(system, platform)
    ctx.setPositions([[0, 0, 0], [1, 0, 0]])
    with pytest.raises(Exception, match="reduced form"):
        ctx.getState(getEnergy=True)


def test_ewald_c1(nbs, platform, systems, oracle):
    """The TIP3P box of BASELINE.json config 0 with the method switched to Ewald, three repeated evaluations
    (plain launches, graph capture, graph replay)."""
    s = systems.make_system("C1")
    s.force.setNonbondedMethod(s.force.Ewald)
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    lam = np.random.default_rng(5).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    r = oracle.evaluate(kernel.desc, s.positions, s.box, lam, None, True, True, kind="port")
    for _ in range(3):
        forces = np.zeros((648, 3))
        e = kernel._evaluate(s.positions, s.box, lam, np.zeros(0), True, True, forces)
        assert force_rel_rms(forces, r.forces) <= F_TOL
        check_energies(e, r.slice_energies)


def test_tiny_and_degenerate_systems(nbs, platform, oracle):
    """Two particles; a subset with no particles; no exceptions at all; atoms exactly on the box edge."""
    for positions in ([[0, 0, 0], [0.3, 0.1, 0]], [[0.0, 2.0, 4.0], [3.9, 0.0, 0.05]]):
        system = nbs.System()
        system.setDefaultPeriodicBoxVectors([4, 0, 0], [0, 4, 0], [0, 0, 4])
        force = nbs.SlicedNonbondedForce(3)
        force.setNonbondedMethod(force.PME)
        force.setCutoffDistance(1.2)
        force.setPMEParameters(2.5, 30, 30, 30)
        system.addParticle(1.0)
        system.addParticle(1.0)
        force.addParticle(1.0, 0.2, 0.8)
        force.addParticle(-0.6, 0.25, 0.5)
        force.setParticleSubset(1, 2)
        system.addForce(force)
        ctx, ref = nbs.Context(system, platform), nbs.Context(system, oracle.OraclePlatform("port"))
        for c in (ctx, ref):
            c.setPositions(positions)
        a = ctx.getState(getEnergy=True, getForces=True)
        b = ref.getState(getEnergy=True, getForces=True)
        # two charges: the force is ~1 kJ/mol/nm, so use the reference tests' floor-1 vector comparison
        for k in range(2):
            assert_equal_vec(b.getForces()[k], a.getForces()[k], TOL)
        assert_equal_tol(b.getPotentialEnergy(), a.getPotentialEnergy(), 1e-5)


def test_periodic_cutoff_golden(nbs, platform):
    """testPeriodic (tests/TestSlicedNonbondedForce.h:358-392) on the CUDA path: reaction-field cutoff,
    one excluded pair, minimum image across the box."""
    system = nbs.System()
    sliced = nbs.SlicedNonbondedForce(1)
    for _ in range(3):
        system.addParticle(1.0)
        sliced.addParticle(1.0, 1, 0)
    sliced.addException(0, 1, 0.0, 1.0, 0.0)
    sliced.setNonbondedMethod(sliced.CutoffPeriodic)
    cutoff = 2.0
    sliced.setCutoffDistance(cutoff)
    system.setDefaultPeriodicBoxVectors([4, 0, 0], [0, 4, 0], [0, 0, 4])
    system.addForce(sliced)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [2, 0, 0], [3, 0, 0]])
    state = context.getState(getForces=True, getEnergy=True)
    eps = 78.3
    krf = (1.0/cutoff**3)*(eps-1.0)/(2.0*eps+1.0)
    crf = (1.0/cutoff)*(3.0*eps)/(2.0*eps+1.0)
    force = nbs.ONE_4PI_EPS0*(1.0-2.0*krf*1.0)
    forces = state.getForces()
    assert_equal_vec([force, 0, 0], forces[0], TOL)
    assert_equal_vec([-force, 0, 0], forces[1], TOL)
    assert_equal_vec([0, 0, 0], forces[2], TOL)
    assert_equal_tol(2*nbs.ONE_4PI_EPS0*(1.0+krf*1.0-crf), state.getPotentialEnergy(), TOL)


def test_periodic_exceptions_golden(nbs, platform):
    """testPeriodicExceptions (:394-430)."""
    system = nbs.System()
    sliced = nbs.SlicedNonbondedForce(1)
    for _ in range(2):
        system.addParticle(1.0)
        sliced.addParticle(1.0, 1, 0)
    sliced.addException(0, 1, 1.0, 1.0, 0.0)
    sliced.setNonbondedMethod(sliced.CutoffPeriodic)
    sliced.setCutoffDistance(2.0)
    system.setDefaultPeriodicBoxVectors([4, 0, 0], [0, 4, 0], [0, 0, 4])
    system.addForce(sliced)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [3, 0, 0]])
    state = context.getState(getForces=True, getEnergy=True)
    force = nbs.ONE_4PI_EPS0/9
    assert_equal_vec([-force, 0, 0], state.getForces()[0], TOL)
    assert_equal_vec([force, 0, 0], state.getForces()[1], TOL)
    assert_equal_tol(nbs.ONE_4PI_EPS0/3, state.getPotentialEnergy(), TOL)
    sliced.setExceptionsUsePeriodicBoundaryConditions(True)
    context.reinitialize(True)
    state = context.getState(getForces=True, getEnergy=True)
    force = nbs.ONE_4PI_EPS0
    assert_equal_vec([force, 0, 0], state.getForces()[0], TOL)
    assert_equal_vec([-force, 0, 0], state.getForces()[1], TOL)
    assert_equal_tol(nbs.ONE_4PI_EPS0, state.getPotentialEnergy(), TOL)


def test_direct_and_reciprocal_groups(nbs, platform):
    """testDirectAndReciprocal (:987-1029) on the CUDA path."""
    system = nbs.System()
    for _ in range(4):
        system.addParticle(1.0)
    system.setDefaultPeriodicBoxVectors([2, 0, 0], [0, 2, 0], [0, 0, 2])
    force = nbs.SlicedNonbondedForce(1)
    system.addForce(force)
    force.setNonbondedMethod(force.PME)
    force.setCutoffDistance(1.0)
    force.setReciprocalSpaceForceGroup(1)
    force.addParticle(1.0, 0.5, 1.0)
    force.addParticle(1.0, 0.5, 1.0)
    force.addParticle(-1.0, 0.5, 1.0)
    force.addParticle(-1.0, 0.5, 1.0)
    force.addException(0, 2, -2.0, 0.5, 3.0)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [1.5, 0, 0], [0, 0.5, 0.5], [0.2, 1.3, 0]])
    e1 = context.getState(getEnergy=True).getPotentialEnergy()
    e2 = context.getState(getEnergy=True, groups=1 << 0).getPotentialEnergy()
    e3 = context.getState(getEnergy=True, groups=1 << 1).getPotentialEnergy()
    assert_equal_tol(e1, e2+e3, 1e-4)
    assert e2 != 0 and e3 != 0
    force.setIncludeDirectSpace(False)
    context.reinitialize(True)
    assert_equal_tol(e3, context.getState(getEnergy=True).getPotentialEnergy(), 1e-4)


def test_errors_mirror_reference(nbs, platform):
    """Box smaller than twice the cutoff (ReferenceNonbondedSlicingKernels.cpp:200-204) and the PME query
    on a non-PME context (:321-328)."""
    system = nbs.System()
    system.setDefaultPeriodicBoxVectors([3, 0, 0], [0, 3, 0], [0, 0, 3])
    force = nbs.SlicedNonbondedForce(1)
    force.setNonbondedMethod(force.CutoffPeriodic)
    force.setCutoffDistance(1.2)
    for _ in range(2):
        system.addParticle(1.0)
        force.addParticle(0.5, 0.3, 0.2)
    system.addForce(force)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [1, 0, 0]])
    context.setPeriodicBoxVectors([2.3, 0, 0], [0, 3, 0], [0, 0, 3])
    with pytest.raises(nbs.abi.NbsError, match="less than twice the nonbonded cutoff") as err:
        context.getState(getEnergy=True)
    assert err.value.status == nbs.abi.NBS_ERR_BOX
    with pytest.raises(nbs.OpenMMException, match="not using PME"):
        force.getPMEParametersInContext(context)


def test_deterministic_forces(nbs, systems):
    """testDeterministicForces (platforms/cuda/tests/TestCudaSlicedNonbondedForce.cpp:109-141) in the reference's own
    shape: FULL PME in the triclinic box (6,0,0), (2.1,6,0), (-1.5,-0.5,6), 1,000 charges of +-1 at random positions,
    the platform's deterministic-forces property set -- two evaluations must agree bit for bit.  Direct space always
    does (64-bit fixed-point accumulators, fixed summation order); reciprocal space does because
    NBS_FLAG_DETERMINISTIC makes the charge spreading accumulate in 64-bit fixed point (pme.cc:108-109, 124-134).
    Without the flag the spreading uses floating-point atomics, whose order -- and rounding -- varies from run to
    run: expected, documented, and checked here only to stay within the parity tolerance."""
    rng = np.random.default_rng(0)
    n = 1000
    system = nbs.System()
    system.setDefaultPeriodicBoxVectors([6, 0, 0], [2.1, 6, 0], [-1.5, -0.5, 6])
    force = nbs.SlicedNonbondedForce(1)
    force.setNonbondedMethod(force.PME)
    for i in range(n):
        system.addParticle(1.0)
        force.addParticle(1.0 if i % 2 == 0 else -1.0, 1.0, 0.0)
    system.addForce(force)
    positions = (rng.random((n, 3)) - 0.5)*6
    box = np.array([[6, 0, 0], [2.1, 6, 0], [-1.5, -0.5, 6]], dtype=float)
    lam = np.ones((1, 2))
    for energies in (False, True):                      # fp32 grids (forces only) and fp64 grids
        kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=nbs.abi.NBS_FLAG_DETERMINISTIC))
        kernel.initialize(system, force)
        out = []
        for _ in range(4):
            f = np.zeros((n, 3))
            if energies:
                kernel._evaluate(positions, box, lam, np.zeros(0), True, True, f)
            else:
                import torch
                pos = torch.tensor(positions, dtype=torch.float64, device="cuda")
                fd = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
                kernel.execute_device(pos.data_ptr(), box, fd.data_ptr(), lam, want_energies=False)
                torch.cuda.synchronize()
                f = fd.cpu().numpy()
            out.append(f)
        for f in out[1:]:
            assert np.array_equal(out[0], f)
    plain = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
    plain.initialize(system, force)
    f = np.zeros((n, 3))
    plain._evaluate(positions, box, lam, np.zeros(0), True, True, f)
    assert force_rel_rms(f, out[0]) < 1e-6


@pytest.mark.parametrize("name,step", [("C2", 0.004), ("C2", 0.012), ("T1_pme", 0.006), ("C1", 0.01)])
def test_list_reuse_matches_rebuild(nbs, platform, systems, name, step):
    """A neighbour list built with a skin and re-used while atoms move (what the plugin's CUDA platform inherits
    from OpenMM, CommonNonbondedSlicingKernels.cpp:721) must never change a result: along a ballistic trajectory
    (every atom has its own velocity, `step` nm per evaluation for the fastest ones, reversed whenever the fastest
    atom is 0.08 nm from its start -- further out atoms overlap and forces leave the range of the 64-bit fixed-point
    accumulators, 2^31 kJ/mol/nm, which is OpenMM's own limit; atoms cross the periodic boundaries, of a triclinic box too) the interacting-pair set is IDENTICAL (count + hash) to that of a context that
    rebuilds its list on every evaluation, forces agree to 1e-5 (single-precision rounding in a different summation order; the last step is also held to the oracle), slice energies to 1e-9 -- including the evaluations
    that find the displacement limit exceeded and are redone with a fresh list."""
    s = systems.make_variant(name) if name in systems.VARIANTS else systems.make_system(name)
    n = s.force.getNumParticles()
    rng = np.random.default_rng(17)
    lam = rng.uniform(0.3, 1.0, size=(s.force.getNumSlices(), 2))
    velocity = rng.normal(size=(n, 3))
    velocity *= step/np.abs(velocity).max()
    reuse = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    reuse.initialize(s.system, s.force)
    fresh = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=nbs.abi.NBS_FLAG_NO_LIST_REUSE))
    fresh.initialize(s.system, s.force)
    gv = np.full(max(s.force.getNumGlobalParameters(), 1), 0.45) if name in systems.VARIANTS else np.zeros(0)
    steps = 50
    turn = max(2, int(round(0.08/step)))
    for t in range(steps):
        phase = t % (2*turn)
        pos = s.positions + (phase if phase <= turn else 2*turn - phase)*velocity
        fa, fb = np.zeros((n, 3)), np.zeros((n, 3))
        ea = reuse._evaluate(pos, s.box, lam, gv, True, True, fa)
        eb = fresh._evaluate(pos, s.box, lam, gv, True, True, fb)
        assert reuse.getPairSet(with_pairs=False)[:2] == fresh.getPairSet(with_pairs=False)[:2], t
        assert force_rel_rms(fa, fb) < 1e-5, t        # fp32 pair forces summed in a different order (the lists differ)
        # (Coulomb terms: the double-precision part agrees to 1e-12, the single-precision remainders of the erfc table --
        # 3 % of a term -- are summed in fp32 over a tile: 2e-9 measured, held to 2e-8; Lennard-Jones terms are fp32 values summed in fp32 over a tile, and
        # the tiles of a skin-padded list differ from those of a fresh one: measured up to 7e-7, held to 2e-6 -- a fifth
        # of the parity tolerance against the oracle, which the last step is also held to below; absolute floor 1e-5 kJ/mol:
        # a small slice total next to close contacts is the fp32 sum of terms a thousand times larger)
        assert np.allclose(ea[:, 0], eb[:, 0], rtol=2e-8, atol=1e-6) and np.allclose(ea[:, 1], eb[:, 1], rtol=2e-6, atol=1e-5), t
    from oracle import oracle as cpu
    r = cpu.evaluate(reuse.desc, pos, s.box, lam, gv if len(gv) else None, True, True, kind="port")
    assert force_rel_rms(fa, r.forces) <= F_TOL and (r.pair_count, r.pair_hash) == reuse.getPairSet(with_pairs=False)[:2]
    check_energies(ea, r.slice_energies)
    stats, stats_fresh = reuse.getListStats(), fresh.getListStats()
    assert stats_fresh["builds"] == steps and stats_fresh["redone"] == 0
    assert stats["builds"] < steps/2, stats                      # the list really was re-used ...
    assert stats["builds"] >= 2, stats                           # ... and really was rebuilt when atoms had moved
    assert stats["evaluations"] == steps + stats["redone"], stats


def test_list_reuse_survives_jumps_and_parameter_changes(nbs, platform, systems, oracle):
    """Things that silently invalidate a kept list: positions replaced wholesale (detected on the device by the
    displacement check), a new box, updated parameters, a changed offset parameter.  Each evaluation against the oracle."""
    s = systems.make_system("C2")
    n = s.force.getNumParticles()
    lam = np.random.default_rng(2).uniform(0.3, 1.0, size=(s.force.getNumSlices(), 2))
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    rng = np.random.default_rng(3)

    def check(pos, box):
        f = np.zeros((n, 3))
        e = kernel._evaluate(pos, box, lam, np.zeros(0), True, True, f)
        r = oracle.evaluate(kernel.desc, pos, box, lam, None, True, True, kind="port")
        assert force_rel_rms(f, r.forces) <= F_TOL
        check_energies(e, r.slice_energies)
        assert kernel.getPairSet(with_pairs=False)[:2] == (r.pair_count, r.pair_hash)
    check(s.positions, s.box)
    check(s.positions + rng.normal(scale=0.003, size=(n, 3)), s.box)          # small move: re-use
    assert kernel.getListStats()["reused_last"]
    shuffled = s.positions.copy()
    waters = shuffled[30:].reshape(-1, 3, 3)
    waters[[5, 900]] = waters[[900, 5]] + 0.0                                  # two molecules trade places: a jump
    check(shuffled, s.box)
    assert kernel.getListStats()["redone"] == 1
    check(shuffled*1.004, s.box*1.004)                                          # new box: fresh list, no redo
    assert kernel.getListStats()["redone"] == 1 and not kernel.getListStats()["reused_last"]
    check(shuffled*1.004, s.box*1.004)
    assert kernel.getListStats()["reused_last"]
    q, sig, eps = s.force.getParticleParameters(3)
    s.force.setParticleParameters(3, q + 0.1, sig*1.05, eps*0.9)
    kernel.desc = nbs.build_desc(s.system, s.force, **kernel._desc_options())
    kernel._update()                                                            # copyParametersToContext
    check(shuffled*1.004, s.box*1.004)
    assert not kernel.getListStats()["reused_last"]


def test_graph_replay_survives_box_change(nbs, platform, systems):
    """A, A, A, B, A: the evaluation at box A is captured into a CUDA graph on its second run; one evaluation at
    another box (a rejected barostat trial) rewrites the influence function on the device, and the graph -- which
    never contains k_eterm -- must not be replayed against it when the old box comes back."""
    import torch
    s = systems.make_system("C2")
    n = s.force.getNumParticles()
    lam = np.random.default_rng(5).uniform(0.3, 1.0, size=(s.force.getNumSlices(), 2))
    pos = torch.tensor(s.positions, dtype=torch.float64, device="cuda")
    box_a, box_b = s.box.copy(), s.box*1.013
    stream = torch.cuda.current_stream().cuda_stream

    def run(kernel, box):
        f = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
        e = kernel.execute_device(pos.data_ptr(), box, f.data_ptr(), lam, stream=stream)
        torch.cuda.synchronize()
        return e, f.cpu().numpy()
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    plain = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=nbs.abi.NBS_FLAG_NO_GRAPH))
    plain.initialize(s.system, s.force)
    # (the force buffer is a new allocation each time; PyTorch's caching allocator hands the same block back,
    # so the graph signature -- which holds the pointers -- repeats)
    results = [run(kernel, b) for b in (box_a, box_a, box_a, box_b, box_a, box_a, box_a)]
    ref_a, ref_b = run(plain, box_a), run(plain, box_b)
    for k in (0, 1, 2, 4, 5, 6):
        assert np.allclose(results[k][0], ref_a[0], rtol=1e-9, atol=1e-7), k
        assert force_rel_rms(results[k][1], ref_a[1]) < 1e-6, k
    assert np.allclose(results[3][0], ref_b[0], rtol=1e-9, atol=1e-7)
    assert not np.allclose(ref_a[0], ref_b[0], rtol=1e-6)


def c5_fixture_errors(golden, forces, energies, tag="full"):
    """Parity figures of a C5 evaluation against tests/golden/C5_reference.npz (the reference's own full-size CPU
    evaluation, oracle/make_golden_c5.py): relative RMS force error over the fixture's 4,096-atom sample, relative
    error of sum |F|^2 over all atoms, worst slice-energy error over max(|E|, 1)."""
    idx = golden["sample"]
    ref = golden[f"{tag}_forces_sample"]
    f_err = float(np.sqrt(((forces[idx]-ref)**2).sum()/(ref**2).sum()))
    sumsq_err = float(abs((forces**2).sum()/golden[f"{tag}_force_sumsq"][0] - 1.0))
    ref_e = golden[f"{tag}_energies"]
    e_err = float(np.max(np.abs(energies-ref_e)/np.maximum(np.abs(ref_e), 1.0)))
    return f_err, sumsq_err, e_err


def test_stmv_size_vs_reference_fixture(nbs, platform, systems):
    """C5 (1,066,628 atoms, 2 subsets, PME 180^3) against the reference's own full-size evaluation: slice energies of
    the full / direct-only / reciprocal-only evaluations, the interacting-pair count and hash (234 M pairs), forces
    of a fixed 4,096-atom sample and sum |F|^2 over all atoms."""
    g = np.load(os.path.join(GOLDEN, "C5_reference.npz"))
    s = systems.make_system("C5")
    assert np.allclose([s.positions.sum(), (s.positions**2).sum()], g["positions_checksum"], rtol=1e-13)
    n = s.force.getNumParticles()
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    for tag, (direct, recip) in {"full": (True, True), "direct": (True, False), "recip": (False, True)}.items():
        f = np.zeros((n, 3))
        e = kernel._evaluate(s.positions, s.box, g["lambdas"], np.zeros(0), direct, recip, f)
        f_err, sumsq_err, e_err = c5_fixture_errors(g, f, e, tag)
        assert f_err <= F_TOL, (tag, f_err)
        assert sumsq_err <= 2*F_TOL, (tag, sumsq_err)
        assert e_err <= E_TOL, (tag, e_err)
        if direct:
            count, h, _ = kernel.getPairSet(with_pairs=False)
            assert (count, h) == (int(g["pair_count"][0]), int(g["pair_hash"][0]))


def test_stmv_size_properties(nbs, platform, systems):
    """C5 (1,066,628 atoms, 180^3 grid): size-independent properties instead of the 5-minute oracle run --
    momentum conservation of the direct-space forces, the lambda-derivative identity
    sum_slices lambda * dE/dlambda = E (tests/TestSlicedNonbondedForce.h:1310-1317), linearity of the energy
    in lambda, and invariance under shifting every atom by whole box vectors."""
    s = systems.make_system("C5")
    n = s.force.getNumParticles()
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    nsl = s.force.getNumSlices()
    lam1 = np.ones((nsl, 2))
    f1 = np.zeros((n, 3))
    e1 = kernel._evaluate(s.positions, s.box, lam1, np.zeros(0), True, True, f1)
    fd = np.zeros((n, 3))
    kernel._evaluate(s.positions, s.box, lam1, np.zeros(0), True, False, fd)
    assert np.abs(fd.sum(axis=0)).max() < 1e-6*np.abs(fd).sum()            # Newton's third law
    count, _, _ = kernel.getPairSet(with_pairs=False)
    assert abs(count/n - 209) < 25                                          # ~209 pairs per atom at this density
    lam2 = np.random.default_rng(1).uniform(0.1, 1.0, size=(nsl, 2))
    f2 = np.zeros((n, 3))
    e2 = kernel._evaluate(s.positions, s.box, lam2, np.zeros(0), True, True, f2)
    assert np.allclose(e1, e2, rtol=1e-6, atol=1e-3)                        # slice energies do not depend on lambda
    shifted = s.positions + np.array([s.box[0, 0], -2*s.box[1, 1], 3*s.box[2, 2]])
    f3 = np.zeros((n, 3))
    e3 = kernel._evaluate(shifted, s.box, lam2, np.zeros(0), True, True, f3)
    assert force_rel_rms(f3, f2) < 1e-5
    assert np.allclose(e3, e2, rtol=1e-5, atol=1e-2)


def test_openmm_cuda_buffer_layouts(nbs, platform, systems, oracle):
    """The layouts the C++ adapter hands over (platform/B200NonbondedSlicingKernels.cpp): OpenMM CUDA's posq
    as float4 / double4 in a permuted, padded atom order + the atomIndex array, and its long-long
    fixed-point force buffer [3][paddedNumAtoms] (pme.cc:382-388), all device resident."""
    import ctypes as C
    import torch
    abi = nbs.abi
    s = systems.make_system("C2")
    n = s.force.getNumParticles()
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(platform)
    kernel.initialize(s.system, s.force)
    lam = np.random.default_rng(9).uniform(0.3, 1.0, size=(s.force.getNumSlices(), 2))
    ref = oracle.evaluate(kernel.desc, s.positions, s.box, lam, None, True, True, kind="port")
    rng = np.random.default_rng(4)
    order = rng.permutation(n).astype(np.int32)              # slot -> particle
    padded = ((n + 31)//32)*32 + 64
    for fmt, dtype in ((abi.NBS_POS_F64_XYZW, torch.float64), (abi.NBS_POS_F32_XYZW, torch.float32)):
        posq = torch.zeros((padded, 4), dtype=dtype, device="cuda")
        posq[:n, :3] = torch.tensor(s.positions[order], dtype=dtype, device="cuda")
        index = torch.tensor(order, dtype=torch.int32, device="cuda")
        forces = torch.zeros((3, padded), dtype=torch.int64, device="cuda")
        kernel._push_parameters(lam, np.zeros(0))
        args = abi.ExecArgs()
        args.struct_size = C.sizeof(abi.ExecArgs)
        args.positions_format, args.positions_space = fmt, abi.NBS_MEM_DEVICE
        args.forces_format, args.forces_space = abi.NBS_FORCE_I64_FIXED, abi.NBS_MEM_DEVICE
        args.positions, args.forces = posq.data_ptr(), forces.data_ptr()
        args.padded_num_atoms = padded
        args.atom_index = index.data_ptr()
        args.box[:] = list(s.box.reshape(9))
        args.include_forces = args.include_energy = args.include_direct = args.include_reciprocal = 1
        energies = np.zeros((s.force.getNumSlices(), 2))
        args.slice_energies = energies.ctypes.data_as(C.POINTER(C.c_double))
        args.stream = torch.cuda.current_stream().cuda_stream
        for _ in range(2):                                   # the buffer is accumulated into, like OpenMM's
            abi.check(kernel.lib.nbs_execute(kernel.handle, C.byref(args)))
        torch.cuda.synchronize()
        f = np.zeros((n, 3))
        f[order] = (forces.cpu().numpy().astype(np.float64)/2**32).T[:n]/2
        if fmt == abi.NBS_POS_F64_XYZW:
            assert force_rel_rms(f, ref.forces) <= F_TOL
            check_energies(energies, ref.slice_energies)
            count, h, _ = kernel.getPairSet(with_pairs=False)
            assert (count, h) == (ref.pair_count, ref.pair_hash)
        else:       # float32 coordinates are OpenMM's single-precision mode: positions themselves are rounded
            assert force_rel_rms(f, ref.forces) <= 2e-4


DOUBLE_CASES = [("PME", None, False), ("PME", (0.3, -0.2, 0.25), False), ("LJPME", None, False), ("Ewald", None, False),
                ("CutoffPeriodic", None, True), ("CutoffPeriodic", (0.25, 0.1, -0.3), False), ("CutoffNonPeriodic", None, True),
                ("NoCutoff", None, False)]


@pytest.mark.parametrize("method,tilt,switch", DOUBLE_CASES)
def test_double_precision_mode(nbs, oracle, method, tilt, switch):
    """NBS_FLAG_DOUBLE, the plugin's Precision = double (CommonNonbondedSlicingKernels.cpp:297-299; the reference registers
    every CUDA test in single, mixed and double, platforms/cuda/tests/CMakeLists.txt:22-24): direct space in double
    precision arithmetic from the exact coordinates, double PME grids / transforms / gather.  Against the oracle, which is
    double throughout: forces to 1e-7 relative RMS (measured 2e-8: a hundred times tighter than the mixed-precision default;
    what is left is the 32 fractional bits the coordinates carry -- 6.5e-10 nm in this box, 4e-8 of an r^-13 force at close
    contact -- not the arithmetic).  Slice energies are double-precision sums in the default mode already; they are held to
    the same 1e-5 * max(|E|, 1) here (measured: 6e-8 direct space, 2e-6 on a reciprocal-space cross term of 0.1 kJ/mol that
    is the difference of structure-factor products a thousand times larger -- the coordinate grid again, identical in
    both modes)."""
    rng = np.random.default_rng(77)
    system, force, positions = random_system(nbs, rng, n=400, nsub=3, L=2.8, grid=(24, 20, 24), method=method, tilt=tilt)
    if method == "LJPME":
        force.setLJPMEParameters(2.4, 14, 12, 14)
    if switch:
        force.setUseSwitchingFunction(True)
        force.setSwitchingDistance(0.8*force.getCutoffDistance())
    ctx = nbs.Context(system, nbs.Platform(flags=nbs.abi.NBS_FLAG_DOUBLE))
    mixed = nbs.Context(system, nbs.Platform())
    ref = nbs.Context(system, oracle.OraclePlatform("port"))
    for c in (ctx, mixed, ref):
        c.setPositions(positions)
        c.setParameter("lc", 0.6)
        c.setParameter("lv", 0.3)
    for groups in (0xFFFFFFFF,):
        for _ in range(3):                      # plain launches, graph capture, graph replay
            a = ctx.getState(getEnergy=True, getForces=True, groups=groups)
        m = mixed.getState(getEnergy=True, getForces=True, groups=groups)
        b = ref.getState(getEnergy=True, getForces=True, groups=groups)
        err_double, err_mixed = force_rel_rms(a.getForces(), b.getForces()), force_rel_rms(m.getForces(), b.getForces())
        assert err_double <= 1e-7, (err_double, err_mixed)
        assert err_double < 0.05*err_mixed, (err_double, err_mixed)          # and it really is another arithmetic
        ea, eb = ctx.impls[0].kernel.lastSliceEnergies, ref.impls[0].kernel.lastSliceEnergies
        err = np.abs(ea-eb)/np.maximum(np.abs(eb), 1.0)
        assert err.max() <= E_TOL, err


def test_double_precision_mode_c2_three_way(nbs, systems, oracle):
    """C2 (7,530 atoms, exceptions, offsets-free) in double-precision mode: direct-only, reciprocal-only and full
    evaluations against the oracle, forces to 1e-7, slice energies to the common tolerance."""
    g = np.load(os.path.join(GOLDEN, "C2_reference.npz"))
    s = systems.make_system("C2")
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform(flags=nbs.abi.NBS_FLAG_DOUBLE))
    kernel.initialize(s.system, s.force)
    n = s.force.getNumParticles()
    for tag, (direct, recip) in {"direct": (True, False), "recip": (False, True), "full": (True, True)}.items():
        f = np.zeros((n, 3))
        e = kernel._evaluate(s.positions, s.box, g["lambdas"], np.zeros(0), direct, recip, f)
        r = oracle.evaluate(kernel.desc, s.positions, s.box, g["lambdas"], None, direct, recip, kind="port")
        assert force_rel_rms(f, r.forces) <= 1e-7, tag
        err = np.abs(e-r.slice_energies)/np.maximum(np.abs(r.slice_energies), 1.0)
        assert err.max() <= E_TOL, (tag, err)


@pytest.mark.parametrize("lambda_elec,lambda_vdw", [(1.0, 1.0), (0.5, 1.0), (0.0, 0.5)])
def test_c2_alchemical_lambda_settings(nbs, platform, systems, oracle, lambda_elec, lambda_vdw):
    """C2 (alchemical solvation: a 30-atom ligand in 2,500 waters) at the three settings SURVEY 8(d) names --
    (lambda_elec, lambda_vdw) = (1, 1), (0.5, 1), (0, 0.5) -- through the Context API: total energy, forces and BOTH
    energy parameter derivatives against the reference's own TUs."""
    s = systems.make_system("C2")
    cpu = oracle.OraclePlatform("reference" if oracle.available("reference") else "port")
    contexts = [nbs.Context(s.system, platform), nbs.Context(s.system, cpu)]
    for c in contexts:
        c.setPositions(s.positions)
        c.setParameter("lambda_elec", lambda_elec)
        c.setParameter("lambda_vdw", lambda_vdw)
    a, b = (c.getState(getEnergy=True, getForces=True, getParameterDerivatives=True) for c in contexts)
    assert force_rel_rms(a.getForces(), b.getForces()) <= F_TOL
    assert abs(a.getPotentialEnergy() - b.getPotentialEnergy()) <= E_TOL*max(1.0, abs(b.getPotentialEnergy()))
    da, db = a.getEnergyParameterDerivatives(), b.getEnergyParameterDerivatives()
    assert set(da) == {"lambda_elec", "lambda_vdw"} == set(db)
    for name in db:
        assert abs(da[name] - db[name]) <= E_TOL*max(1.0, abs(db[name])), (name, da[name], db[name])

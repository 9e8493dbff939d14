"""CPU tests: the restated oracle against (a) the committed golden fixtures that oracle/make_golden.py
produced with the reference's own compiled TUs, (b) that compiled reference itself on random systems
(when oracle/_ref exists), and (c) the slicing identities of tests/TestSlicedNonbondedForce.h:1031-1457."""
import importlib
import os

import numpy as np
import pytest

from helpers import assert_equal_tol, force_rel_rms

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def systems():
    return importlib.import_module("openmm-nonbonded-slicing_b200.systems")


@pytest.mark.parametrize("name", ["C1", "C2", "C1_ewald", "C1_ljpme", "T1_pme", "T1_ljpme"])
def test_port_matches_reference_fixture(nbs, oracle, systems, name):
    """The restatement against committed outputs of the reference's own TUs (oracle/make_golden.py): the BASELINE
    configurations C1 / C2 and the variants beyond them -- plain Ewald, LJPME, a triclinic box (systems.VARIANTS)."""
    g = np.load(os.path.join(GOLDEN, f"{name}_reference.npz"))
    s = systems.make_variant(name) if name in systems.VARIANTS else systems.make_system(name)
    # the generator is deterministic: same positions as when the fixture was made
    assert np.allclose([s.positions.sum(), (s.positions**2).sum()], g["positions_checksum"], rtol=1e-13)
    desc = nbs.build_desc(s.system, s.force, legal_grid=True)
    gv = g["global_values"] if "global_values" in g.files else None
    for tag, (direct, recip) in {"full": (True, True), "direct": (True, False), "recip": (False, True)}.items():
        r = oracle.evaluate(desc, s.positions, s.box, g["lambdas"], gv, direct, recip, kind="port")
        assert force_rel_rms(r.forces, g[f"{tag}_forces"]) < 1e-12
        assert np.allclose(r.slice_energies, g[f"{tag}_energies"], rtol=1e-11, atol=1e-9)
        if direct:
            assert r.pair_count == int(g["pair_count"][0])
            assert r.pair_hash == int(g["pair_hash"][0])


def random_system(nbs, rng, n=300, nsub=3, L=2.6, grid=(20, 20, 20), with_offsets=True, net_charge=True, method="PME",
                  tilt=None):
    """``tilt`` = (bx, cx, cy) as fractions of (ax, ax, by) makes the box triclinic (OpenMM's reduced form needs
    each within +-0.5); the atoms then sit on the sheared lattice and are displaced by whole box vectors."""
    system = nbs.System()
    force = nbs.SlicedNonbondedForce(nsub)
    force.setNonbondedMethod(getattr(force, method))
    force.setCutoffDistance(1.0)
    force.setPMEParameters(2.8, *grid)
    box = np.array([[L, 0, 0], [0, L, 0], [0, 0, L]], dtype=float)
    if tilt is not None:
        box[1, 0], box[2, 0], box[2, 1] = tilt[0]*L, tilt[1]*L, tilt[2]*L
    system.setDefaultPeriodicBoxVectors(*box)
    side = int(np.ceil(n**(1/3)))
    sites = np.array([(i, j, k) for i in range(side) for j in range(side) for k in range(side)][:n], dtype=float)
    positions = ((sites+0.5)/side) @ box + rng.uniform(-0.05, 0.05, size=(n, 3))
    positions += rng.integers(-2, 3, size=(n, 3)) @ box      # unwrapped coordinates must work too
    charges = rng.uniform(-0.8, 0.8, size=n)
    if not net_charge:
        charges -= charges.mean()
    for i in range(n):
        system.addParticle(1.0)
        force.addParticle(charges[i], rng.uniform(0.15, 0.3), rng.uniform(0.1, 1.0))
        force.setParticleSubset(i, int(rng.integers(0, nsub)))
    bonds = [(i, i+1) for i in range(0, n-1) if i % 5 != 4]
    force.createExceptionsFromBonds(bonds, 1/1.2, 0.5)
    if with_offsets:
        force.addGlobalParameter("off", 0.3)
        force.addParticleParameterOffset("off", 3, 0.5, 0.01, 0.2)
        force.addExceptionParameterOffset("off", 2, 0.2, 0.01, 0.1)
    force.addGlobalParameter("lc", 0.7)
    force.addGlobalParameter("lv", 0.4)
    other = 1 if nsub > 1 else 0
    force.addScalingParameter("lc", 0, other, True, False)
    force.addScalingParameter("lv", 0, other, False, True)
    force.addEnergyParameterDerivative("lc")
    force.addEnergyParameterDerivative("lv")
    system.addForce(force)
    return system, force, positions


@pytest.mark.parametrize("seed", [1, 2])
def test_port_matches_compiled_reference(nbs, oracle, seed):
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng)
    states = []
    for kind in ("port", "reference"):
        ctx = nbs.Context(system, oracle.OraclePlatform(kind))
        ctx.setPositions(positions)
        ctx.setParameter("off", 0.45)
        states.append(ctx.getState(getEnergy=True, getForces=True, getParameterDerivatives=True))
    a, b = states
    assert force_rel_rms(a.getForces(), b.getForces()) < 1e-12
    assert_equal_tol(b.getPotentialEnergy(), a.getPotentialEnergy(), 1e-12)
    for name, value in b.getEnergyParameterDerivatives().items():
        assert_equal_tol(value, a.getEnergyParameterDerivatives()[name], 1e-12)


def test_dispersion_coefficients_agree(nbs, oracle, systems):
    """The host mirror of calcDispersionCorrections (product side) equals the oracle's restatement."""
    s = systems.make_system("C2")
    host = nbs.SlicedNonbondedForceImpl.calcDispersionCorrections(s.system, s.force)
    desc = nbs.build_desc(s.system, s.force)
    ref = oracle.dispersion_coefficients(desc, np.array([g[1] for g in s.force._globalParams]))
    assert np.allclose(host, ref, rtol=1e-12)
    assert np.any(np.asarray(host) != 0)


def test_slicing_equals_rescaled_parameters(nbs, oracle):
    slicing_equals_rescaled_parameters(nbs, oracle.OraclePlatform("port"), 1e-9, 1e-9)


def slicing_equals_rescaled_parameters(nbs, platform, etol, ftol):
    """testNonbondedSlicing (:1031-1318): scaling slice (0,1) and (1,1) by lambda equals a one-subset
    force whose subset-1 charges are scaled (Coulomb case); sum of all slice derivatives = energy.
    Both Contexts run on `platform` (an oracle, or the B200 platform in test_gpu_known_answers.py)."""
    rng = np.random.default_rng(7)
    n, L = 200, 7.0
    system1, system2 = nbs.System(), nbs.System()
    plain = nbs.SlicedNonbondedForce(1)
    plain.setNonbondedMethod(plain.PME)
    plain.setCutoffDistance(3.0)
    plain.setPMEParameters(1.1, 24, 24, 24)
    plain.setUseDispersionCorrection(True)
    positions = rng.uniform(0, L, size=(n, 3))
    # keep particles apart
    side = 6
    sites = np.array([(i, j, k) for i in range(side) for j in range(side) for k in range(side)][:n], dtype=float)
    positions = (sites+0.5)*L/side + rng.uniform(-0.1, 0.1, size=(n, 3))
    charges = np.array([1-2*(k % 2) for k in range(n)], dtype=float)
    for k in range(n):
        system1.addParticle(1.0)
        system2.addParticle(1.0)
        plain.addParticle(charges[k], 0.5, 1.0)
    for k in range(0, n, 2):
        plain.addException(k, k+1, charges[k]*charges[k+1], 0.5, 1.0)
    for sysm in (system1, system2):
        sysm.setDefaultPeriodicBoxVectors([L, 0, 0], [0, L, 0], [0, 0, L])
    sliced = nbs.SlicedNonbondedForce(plain, 2)
    in1 = rng.random(n) < 0.5
    for k in range(n):
        if in1[k]:
            sliced.setParticleSubset(k, 1)
    sliced.addGlobalParameter("lambda", 1)
    sliced.addScalingParameter("lambda", 0, 1, True, False)
    sliced.addGlobalParameter("lambdaSq", 1)
    sliced.addScalingParameter("lambdaSq", 1, 1, True, False)
    system1.addForce(plain)
    system2.addForce(sliced)
    ctx2 = nbs.Context(system2, platform)
    ctx2.setPositions(positions)
    energies = {}
    for lam in (1.0, 0.0, 0.5):
        for k in range(n):
            plain.setParticleParameters(k, charges[k]*(lam if in1[k] else 1.0), 0.5, 1.0)
        for e in range(plain.getNumExceptions()):
            p1, p2 = plain.getExceptionParameters(e)[:2]
            scale = 1.0
            if in1[p1] != in1[p2]:
                scale = lam
            elif in1[p1]:
                scale = lam*lam
            plain.setExceptionParameters(e, p1, p2, charges[p1]*charges[p2]*scale, 0.5, 1.0)
        ctx1 = nbs.Context(system1, platform)
        ctx1.setPositions(positions)
        ctx2.setParameter("lambda", lam)
        ctx2.setParameter("lambdaSq", lam*lam)
        for groups in (0xFFFFFFFF,):
            s1 = ctx1.getState(getEnergy=True, getForces=True, groups=groups)
            s2 = ctx2.getState(getEnergy=True, getForces=True, groups=groups)
            assert_equal_tol(s1.getPotentialEnergy(), s2.getPotentialEnergy(), etol)
            assert force_rel_rms(s2.getForces(), s1.getForces()) < ftol
        energies[lam] = s2.getPotentialEnergy()
    # derivatives (:1279-1317)
    sliced.addEnergyParameterDerivative("lambda")
    sliced.addEnergyParameterDerivative("lambdaSq")
    sliced.addGlobalParameter("remainderC", 1.0)
    sliced.addScalingParameter("remainderC", 0, 0, True, False)
    sliced.addEnergyParameterDerivative("remainderC")
    ctx2.reinitialize(True)
    ctx2.setParameter("lambda", 1.0)
    ctx2.setParameter("lambdaSq", 1.0)
    d = ctx2.getState(getEnergy=True, getParameterDerivatives=True)
    derivs = d.getEnergyParameterDerivatives()
    assert_equal_tol(energies[1.0]-energies[0.0], derivs["lambda"]+derivs["lambdaSq"], etol)


@pytest.mark.parametrize("seed,nsub,tol", [(11, 3, 5e-4), (12, 1, 1e-4), (13, 4, 1e-5)])
def test_ewald_port_matches_compiled_reference(nbs, oracle, seed, nsub, tol):
    """Plain Ewald (ReferenceSlicedLJCoulombIxn.cpp:256-358): the restatement against the reference's own TU,
    direct-only / reciprocal-only / full, with kmax from the restated calcEwaldParameters."""
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng, n=150, nsub=nsub, L=2.3 + 0.1*nsub, method="Ewald")
    force.setEwaldErrorTolerance(tol)
    desc = nbs.build_desc(system, force)
    assert min(desc.desc.ewald_kmax) >= 1 and all(k % 2 == 1 for k in desc.desc.ewald_kmax)
    box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)
    lam = rng.uniform(0.1, 1.0, size=(force.getNumSlices(), 2))
    gv = np.array([0.3, 0.7, 0.4])
    for direct, recip in ((True, True), (True, False), (False, True)):
        a = oracle.evaluate(desc, positions, box, lam, gv, direct, recip, kind="port")
        b = oracle.evaluate(desc, positions, box, lam, gv, direct, recip, kind="reference")
        assert force_rel_rms(a.forces, b.forces) < 1e-12
        assert np.allclose(a.slice_energies, b.slice_energies, rtol=1e-11, atol=1e-9)
        assert (a.pair_count, a.pair_hash) == (b.pair_count, b.pair_hash)


def test_ewald_agrees_with_pme(nbs, oracle):
    """Second opinion on both reciprocal sums (they share no code): at a tight tolerance the plain Ewald sum and
    PME give the same forces and slice energies (cf. the method matrix of tests/TestSlicedNonbondedForce.h:1493-1500,
    whose arbiter -- OpenMM's NonbondedForce -- is not available here)."""
    rng = np.random.default_rng(21)
    system, force, positions = random_system(nbs, rng, n=160, nsub=3, L=2.4, net_charge=False)
    box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)
    lam = rng.uniform(0.1, 1.0, size=(6, 2))
    gv = np.array([0.3, 0.7, 0.4])
    force.setEwaldErrorTolerance(1e-6)
    force.setPMEParameters(0, 0, 0, 0)
    results = {}
    for method in (force.Ewald, force.PME):
        force.setNonbondedMethod(method)
        results[method] = oracle.evaluate(nbs.build_desc(system, force), positions, box, lam, gv, True, True, kind="port")
    a, b = results[force.Ewald], results[force.PME]
    assert force_rel_rms(a.forces, b.forces) < 1e-5
    assert np.abs(a.slice_energies-b.slice_energies).max() < 1e-5*np.abs(b.slice_energies).max()


def test_ewald_agrees_with_pme_on_c1(nbs, oracle, systems):
    """The same second opinion on BASELINE config 0 (the TIP3P box): with both reciprocal sums converged (tolerance
    1e-6: 77^3 PME grid, 9 Ewald vectors per axis) the absolute forces and slice energies agree -- this is what pins
    absolute PME values here, where OpenMM's own NonbondedForce (the reference tests' arbiter) is not available."""
    s = systems.make_system("C1")
    lam = np.ones((s.force.getNumSlices(), 2))
    s.force.setEwaldErrorTolerance(1e-6)
    s.force.setPMEParameters(0, 0, 0, 0)
    results = {}
    for method in (s.force.Ewald, s.force.PME):
        s.force.setNonbondedMethod(method)
        results[method] = oracle.evaluate(nbs.build_desc(s.system, s.force), s.positions, s.box, lam, None, True, True, kind="port")
    a, b = results[s.force.Ewald], results[s.force.PME]
    assert force_rel_rms(a.forces, b.forces) < 5e-6                      # measured 7e-7
    assert np.abs(a.slice_energies-b.slice_energies).max() < 5e-3        # kJ/mol; measured 3e-4, out of a self energy of 6e4


def test_ewald_rejects_triclinic(nbs, oracle):
    system = nbs.System()
    system.setDefaultPeriodicBoxVectors([3, 0, 0], [0.5, 3, 0], [0, 0, 3])
    force = nbs.SlicedNonbondedForce(1)
    force.setNonbondedMethod(force.Ewald)
    system.addParticle(1.0)
    force.addParticle(1.0, 0.3, 0.1)
    system.addForce(force)
    with pytest.raises(nbs.OpenMMException, match="Ewald is not supported with non-rectangular boxes"):
        nbs.Context(system, oracle.OraclePlatform("port"))


@pytest.mark.parametrize("seed,tilt,method", [(51, (0.3, -0.2, 0.4), "PME"), (52, (-0.5, 0.5, -0.5), "PME"),
                                              (53, (0.25, 0.1, -0.35), "CutoffPeriodic")])
def test_triclinic_port_matches_compiled_reference(nbs, oracle, seed, tilt, method):
    """Triclinic boxes (testTriclinic :432-492 pins the minimum image analytically; this pins the rest): the
    restatement against the reference's own TUs on sheared random systems, including the extreme reduced form."""
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng, n=200, nsub=3, L=2.5, grid=(20, 20, 20), method=method, tilt=tilt)
    desc = nbs.build_desc(system, force)
    box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)
    lam = rng.uniform(0.1, 1.0, size=(force.getNumSlices(), 2))
    gv = np.array([0.3, 0.7, 0.4])
    for direct, recip in ((True, True), (True, False), (False, True)):
        a = oracle.evaluate(desc, positions, box, lam, gv, direct, recip, kind="port")
        b = oracle.evaluate(desc, positions, box, lam, gv, direct, recip, kind="reference")
        if np.abs(b.forces).max() > 0:
            assert force_rel_rms(a.forces, b.forces) < 1e-11
        else:
            assert np.abs(a.forces).max() == 0
        assert np.allclose(a.slice_energies, b.slice_energies, rtol=1e-10, atol=1e-9)
        assert (a.pair_count, a.pair_hash) == (b.pair_count, b.pair_hash)


@pytest.mark.parametrize("seed,nsub,tilt,grid,dgrid", [(71, 3, None, (20, 18, 20), (12, 12, 12)), (72, 1, None, (24, 18, 30), (10, 14, 9)),
                                                        (73, 4, (0.3, -0.2, 0.4), (20, 20, 20), (10, 14, 10))])
def test_ljpme_port_matches_compiled_reference(nbs, oracle, seed, nsub, tilt, grid, dgrid):
    """LJPME (ReferenceSlicedLJCoulombIxn.cpp:211-212, 241-253, 398-426, 487-504; ReferencePME.cpp:499-595, 814-871):
    the restatement against the reference's own TUs.  Multi-subset cases keep nx == nz (the reference's
    gather indexes subset grids with sj*nz)."""
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(seed)
    system, force, positions = random_system(nbs, rng, n=180, nsub=nsub, L=2.6, grid=grid, method="LJPME", tilt=tilt)
    force.setLJPMEParameters(2.4, *dgrid)
    desc = nbs.build_desc(system, force)
    assert desc.desc.use_switching_function == 0
    ctx = nbs.Context(system, oracle.OraclePlatform("port"))
    assert force.getLJPMEParametersInContext(ctx) == (2.4,)+tuple(dgrid)
    assert force.getPMEParametersInContext(ctx)[1:] == tuple(grid)
    box = np.array(system.getDefaultPeriodicBoxVectors()).reshape(9)
    lam = rng.uniform(0.1, 1.0, size=(force.getNumSlices(), 2))
    gv = np.array([0.3, 0.7, 0.4])
    for direct, recip in ((True, True), (True, False), (False, True)):
        a = oracle.evaluate(desc, positions, box, lam, gv, direct, recip, kind="port")
        b = oracle.evaluate(desc, positions, box, lam, gv, direct, recip, kind="reference")
        assert force_rel_rms(a.forces, b.forces) < 1e-12
        assert np.allclose(a.slice_energies, b.slice_energies, rtol=1e-11, atol=1e-9)
        assert (a.pair_count, a.pair_hash) == (b.pair_count, b.pair_hash)

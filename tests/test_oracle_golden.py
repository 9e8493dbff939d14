"""Pins the CPU oracles against the analytic known answers of the reference's own test suite
(tests/TestSlicedNonbondedForce.h; line ranges cited per test).  Runs on both oracle backends:
the restatement ("port") and -- when oracle/_ref was built from /root/reference -- the reference's
unmodified TUs ("reference")."""
import math

import numpy as np
import pytest

from helpers import TOL, assert_equal_tol, assert_equal_vec


def _kinds(oracle):
    return ["port"] + (["reference"] if oracle.available("reference") else [])


@pytest.fixture(params=["port", "reference"])
def platform(request, oracle):
    if request.param == "reference" and not oracle.available("reference"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return oracle.OraclePlatform(request.param)


def test_coulomb(nbs, platform):
    """testCoulomb :87-109"""
    system = nbs.System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    ff = nbs.SlicedNonbondedForce(1)
    ff.addParticle(0.5, 1, 0)
    ff.addParticle(-1.5, 1, 0)
    system.addForce(ff)
    assert not ff.usesPeriodicBoundaryConditions()
    assert not system.usesPeriodicBoundaryConditions()
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [2, 0, 0]])
    state = context.getState(getForces=True, getEnergy=True)
    force = nbs.ONE_4PI_EPS0*(-0.75)/4.0
    assert_equal_vec([-force, 0, 0], state.getForces()[0], TOL)
    assert_equal_vec([force, 0, 0], state.getForces()[1], TOL)
    assert_equal_tol(nbs.ONE_4PI_EPS0*(-0.75)/2.0, state.getPotentialEnergy(), TOL)


def test_lj(nbs, platform):
    """testLJ :111-135"""
    system = nbs.System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    ff = nbs.SlicedNonbondedForce(1)
    ff.addParticle(0, 1.2, 1)
    ff.addParticle(0, 1.4, 2)
    system.addForce(ff)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [2, 0, 0]])
    state = context.getState(getForces=True, getEnergy=True)
    x = 1.3/2.0
    eps = math.sqrt(2.0)
    force = 4.0*eps*(12*x**12-6*x**6)/2.0
    assert_equal_vec([-force, 0, 0], state.getForces()[0], TOL)
    assert_equal_vec([force, 0, 0], state.getForces()[1], TOL)
    assert_equal_tol(4.0*eps*(x**12-x**6), state.getPotentialEnergy(), TOL)


def _chain_force(nbs, method=None, cutoff=None, eps=None):
    system = nbs.System()
    sliced = nbs.SlicedNonbondedForce(1)
    if method is not None:
        sliced.setNonbondedMethod(method)
    for _ in range(5):
        system.addParticle(1.0)
        sliced.addParticle(0, 1.5, 0)
    if cutoff is not None:
        sliced.setCutoffDistance(cutoff)
        sliced.setReactionFieldDielectric(eps)
    sliced.createExceptionsFromBonds([(0, 1), (1, 2), (2, 3), (3, 4)], 0.0, 0.0)
    first14 = second14 = None
    for i in range(sliced.getNumExceptions()):
        p1, p2 = sliced.getExceptionParameters(i)[:2]
        if {p1, p2} == {0, 3}:
            first14 = i
        if {p1, p2} == {1, 4}:
            second14 = i
    system.addForce(sliced)
    return system, sliced, first14, second14


def test_exclusions_and_14(nbs, platform):
    """testExclusionsAnd14 :137-222"""
    system, sliced, first14, second14 = _chain_force(nbs)
    context = nbs.Context(system, platform)
    for i in range(1, 5):
        r = 1.0
        positions = [[0, j, 0] for j in range(5)]
        for j in range(5):
            sliced.setParticleParameters(j, 0, 1.5, 0)
        sliced.setParticleParameters(0, 0, 1.5, 1)
        sliced.setParticleParameters(i, 0, 1.5, 1)
        sliced.setExceptionParameters(first14, 0, 3, 0, 1.5, 0.5 if i == 3 else 0.0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0.0)
        positions[i] = [r, 0, 0]
        context.reinitialize()
        context.setPositions(positions)
        state = context.getState(getForces=True, getEnergy=True)
        x = 1.5/r
        force = 4.0*(12*x**12-6*x**6)/r
        energy = 4.0*(x**12-x**6)
        if i == 3:
            force *= 0.5
            energy *= 0.5
        if i < 3:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], state.getForces()[0], TOL)
        assert_equal_vec([force, 0, 0], state.getForces()[i], TOL)
        assert_equal_tol(energy, state.getPotentialEnergy(), TOL)
        # Coulomb
        sliced.setParticleParameters(0, 2, 1.5, 0)
        sliced.setParticleParameters(i, 2, 1.5, 0)
        sliced.setExceptionParameters(first14, 0, 3, 4/1.2 if i == 3 else 0, 1.5, 0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0)
        context.reinitialize()
        context.setPositions(positions)
        state = context.getState(getForces=True, getEnergy=True)
        force = nbs.ONE_4PI_EPS0*4/(r*r)
        energy = nbs.ONE_4PI_EPS0*4/r
        if i == 3:
            force /= 1.2
            energy /= 1.2
        if i < 3:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], state.getForces()[0], TOL)
        assert_equal_vec([force, 0, 0], state.getForces()[i], TOL)
        assert_equal_tol(energy, state.getPotentialEnergy(), TOL)


def test_cutoff(nbs, platform):
    """testCutoff :224-260"""
    system = nbs.System()
    ff = nbs.SlicedNonbondedForce(1)
    for _ in range(3):
        system.addParticle(1.0)
        ff.addParticle(1.0, 1, 0)
    ff.setNonbondedMethod(ff.CutoffNonPeriodic)
    cutoff, eps = 2.9, 50.0
    ff.setCutoffDistance(cutoff)
    ff.setReactionFieldDielectric(eps)
    system.addForce(ff)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [0, 2, 0], [0, 3, 0]])
    state = context.getState(getForces=True, getEnergy=True)
    K = nbs.ONE_4PI_EPS0
    krf = (1.0/cutoff**3)*(eps-1.0)/(2.0*eps+1.0)
    crf = (1.0/cutoff)*(3.0*eps)/(2.0*eps+1.0)
    force1 = K*(0.25-2.0*krf*2.0)
    force2 = K*(1.0-2.0*krf*1.0)
    forces = state.getForces()
    assert_equal_vec([0, -force1, 0], forces[0], TOL)
    assert_equal_vec([0, force1-force2, 0], forces[1], TOL)
    assert_equal_vec([0, force2, 0], forces[2], TOL)
    assert_equal_tol(K*(0.5+krf*4.0-crf) + K*(1.0+krf*1.0-crf), state.getPotentialEnergy(), TOL)


def test_cutoff14(nbs, platform):
    """testCutoff14 :262-356"""
    cutoff, eps = 3.5, 30.0
    system, sliced, first14, second14 = _chain_force(nbs, nbs.SlicedNonbondedForce.CutoffNonPeriodic, cutoff, eps)
    context = nbs.Context(system, platform)
    positions = [[float(k), 0, 0] for k in range(5)]
    context.setPositions(positions)
    for i in range(1, 5):
        sliced.setParticleParameters(0, 0, 1.5, 1)
        for j in range(1, 5):
            sliced.setParticleParameters(j, 0, 1.5, 0)
        sliced.setParticleParameters(i, 0, 1.5, 1)
        sliced.setExceptionParameters(first14, 0, 3, 0, 1.5, 0.5 if i == 3 else 0.0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0.0)
        context.reinitialize(True)
        state = context.getState(getForces=True, getEnergy=True)
        r = positions[i][0]
        x = 1.5/r
        force = 4.0*(12*x**12-6*x**6)/r
        energy = 4.0*(x**12-x**6)
        if i == 3:
            force *= 0.5
            energy *= 0.5
        if i < 3 or r > cutoff:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], state.getForces()[0], TOL)
        assert_equal_vec([force, 0, 0], state.getForces()[i], TOL)
        assert_equal_tol(energy, state.getPotentialEnergy(), TOL)
        q = 0.7
        sliced.setParticleParameters(0, q, 1.5, 0)
        sliced.setParticleParameters(i, q, 1.5, 0)
        sliced.setExceptionParameters(first14, 0, 3, q*q/1.2 if i == 3 else 0, 1.5, 0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0)
        context.reinitialize(True)
        state = context.getState(getForces=True, getEnergy=True)
        force = nbs.ONE_4PI_EPS0*q*q/(r*r)
        energy = nbs.ONE_4PI_EPS0*q*q/r
        if i == 3:
            force /= 1.2
            energy /= 1.2
        if i < 3 or r > cutoff:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], state.getForces()[0], TOL)
        assert_equal_vec([force, 0, 0], state.getForces()[i], TOL)
        assert_equal_tol(energy, state.getPotentialEnergy(), TOL)


def test_periodic(nbs, platform):
    """testPeriodic :358-392"""
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
    assert sliced.usesPeriodicBoundaryConditions() and system.usesPeriodicBoundaryConditions()
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


def test_periodic_exceptions(nbs, platform):
    """testPeriodicExceptions :394-430"""
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


def test_triclinic(nbs, platform):
    """testTriclinic :432-492 (positions from numpy's RNG instead of SFMT; the check is analytic)"""
    system = nbs.System()
    a, b, c = np.array([3.1, 0, 0]), np.array([0.4, 3.5, 0]), np.array([-0.1, -0.5, 4.0])
    system.setDefaultPeriodicBoxVectors(a, b, c)
    sliced = nbs.SlicedNonbondedForce(1)
    for _ in range(2):
        system.addParticle(1.0)
        sliced.addParticle(1.0, 1, 0)
    sliced.setNonbondedMethod(sliced.CutoffPeriodic)
    cutoff = 1.5
    sliced.setCutoffDistance(cutoff)
    system.addForce(sliced)
    context = nbs.Context(system, platform)
    rng = np.random.default_rng(0)
    eps = 78.3
    krf = (1.0/cutoff**3)*(eps-1.0)/(2.0*eps+1.0)
    crf = (1.0/cutoff)*(3.0*eps)/(2.0*eps+1.0)
    for _ in range(50):
        u = rng.random((2, 3))
        positions = [a*u[k, 0] + b*u[k, 1] + c*u[k, 2] for k in range(2)]
        context.setPositions(positions)
        best, delta = 100.0, None
        for i in (-1, 0, 1):
            for j in (-1, 0, 1):
                for k in (-1, 0, 1):
                    d = positions[1]-positions[0]+a*i+b*j+c*k
                    if d.dot(d) < best:
                        best, delta = d.dot(d), d
        distance = math.sqrt(best)
        state = context.getState(getForces=True, getEnergy=True)
        if distance >= cutoff:
            assert state.getPotentialEnergy() == 0.0
            assert np.all(state.getForces() == 0)
        else:
            force = delta*nbs.ONE_4PI_EPS0*(-1.0/distance**3+2.0*krf)
            assert_equal_tol(nbs.ONE_4PI_EPS0*(1.0/distance+krf*distance*distance-crf), state.getPotentialEnergy(), 1e-4)
            assert_equal_vec(force, state.getForces()[0], 1e-4)
            assert_equal_vec(-force, state.getForces()[1], 1e-4)


def test_dispersion_correction(nbs, platform):
    """testDispersionCorrection :614-681"""
    gridSize = 5
    numParticles = gridSize**3
    boxSize = gridSize*0.7
    cutoff = boxSize/3
    system = nbs.System()
    sliced = nbs.SlicedNonbondedForce(1)
    positions = []
    for i in range(gridSize):
        for j in range(gridSize):
            for k in range(gridSize):
                system.addParticle(1.0)
                sliced.addParticle(0, 1.1, 0.5)
                positions.append([i*boxSize/gridSize, j*boxSize/gridSize, k*boxSize/gridSize])
    sliced.setNonbondedMethod(sliced.CutoffPeriodic)
    sliced.setCutoffDistance(cutoff)
    system.setDefaultPeriodicBoxVectors([boxSize, 0, 0], [0, boxSize, 0], [0, 0, boxSize])
    system.addForce(sliced)
    context = nbs.Context(system, platform)
    context.setPositions(positions)
    energy1 = context.getState(getEnergy=True).getPotentialEnergy()
    sliced.setUseDispersionCorrection(False)
    context.reinitialize()
    context.setPositions(positions)
    energy2 = context.getState(getEnergy=True).getPotentialEnergy()
    term1 = (0.5*1.1**12/cutoff**9)/9
    term2 = (0.5*1.1**6/cutoff**3)/3
    expected = 8*math.pi*numParticles*numParticles*(term1-term2)/boxSize**3
    assert_equal_tol(expected, energy1-energy2, TOL)
    numType2 = 0
    for i in range(0, numParticles, 2):
        sliced.setParticleParameters(i, 0, 1, 1)
        numType2 += 1
    numType1 = numParticles-numType2
    sliced.updateParametersInContext(context)
    energy2 = context.getState(getEnergy=True).getPotentialEnergy()
    sliced.setUseDispersionCorrection(True)
    context.reinitialize()
    context.setPositions(positions)
    energy1 = context.getState(getEnergy=True).getPotentialEnergy()
    term1 = ((numType1*(numType1+1))//2)*(0.5*1.1**12/cutoff**9)/9
    term2 = ((numType1*(numType1+1))//2)*(0.5*1.1**6/cutoff**3)/3
    term1 += ((numType2*(numType2+1))//2)*(1*1.0**12/cutoff**9)/9
    term2 += ((numType2*(numType2+1))//2)*(1*1.0**6/cutoff**3)/3
    combinedSigma = 0.5*(1+1.1)
    combinedEpsilon = math.sqrt(1*0.5)
    term1 += (numType1*numType2)*(combinedEpsilon*combinedSigma**12/cutoff**9)/9
    term2 += (numType1*numType2)*(combinedEpsilon*combinedSigma**6/cutoff**3)/3
    term1 /= (numParticles*(numParticles+1))//2
    term2 /= (numParticles*(numParticles+1))//2
    expected = 8*math.pi*numParticles*numParticles*(term1-term2)/boxSize**3
    assert_equal_tol(expected, energy1-energy2, TOL)


def test_switching_function(nbs, platform):
    """testSwitchingFunction :760-813 (CutoffNonPeriodic leg)"""
    system = nbs.System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    ff = nbs.SlicedNonbondedForce(1)
    ff.addParticle(0, 1.2, 1)
    ff.addParticle(0, 1.4, 2)
    ff.setNonbondedMethod(ff.CutoffNonPeriodic)
    ff.setCutoffDistance(2.0)
    ff.setUseSwitchingFunction(True)
    ff.setSwitchingDistance(1.5)
    ff.setUseDispersionCorrection(False)
    system.addForce(ff)
    context = nbs.Context(system, platform)
    eps = math.sqrt(2.0)
    r = 1.0
    while r < 2.5:
        context.setPositions([[0, 0, 0], [r, 0, 0]])
        state = context.getState(getForces=True, getEnergy=True)
        x = 1.3/r
        expectedEnergy = 4.0*eps*(x**12-x**6)
        switchValue = 1.0
        if r > 1.5:
            t = (r-1.5)/0.5
            switchValue = 1+t*t*t*(-10+t*(15-t*6))
        if r >= 2.0:
            switchValue = 0.0
        assert_equal_tol(switchValue*expectedEnergy, state.getPotentialEnergy(), TOL)
        delta = 1e-3
        context.setPositions([[0, 0, 0], [r-delta, 0, 0]])
        e1 = context.getState(getEnergy=True).getPotentialEnergy()
        context.setPositions([[0, 0, 0], [r+delta, 0, 0]])
        e2 = context.getState(getEnergy=True).getPotentialEnergy()
        assert_equal_tol((e2-e1)/(2*delta), state.getForces()[0][0], 1e-3)
        r += 0.1


def test_parameter_offsets(nbs, platform):
    """testParameterOffsets :883-945"""
    system = nbs.System()
    for _ in range(4):
        system.addParticle(1.0)
    force = nbs.SlicedNonbondedForce(1)
    force.addParticle(0.0, 1.0, 0.5)
    force.addParticle(1.0, 0.5, 0.6)
    force.addParticle(-1.0, 2.0, 0.7)
    force.addParticle(0.5, 2.0, 0.8)
    force.addException(0, 3, 0.0, 1.0, 0.0)
    force.addException(2, 3, 0.5, 1.0, 1.5)
    force.addException(0, 1, 1.0, 1.5, 1.0)
    force.addGlobalParameter("p1", 0.0)
    force.addGlobalParameter("p2", 1.0)
    force.addParticleParameterOffset("p1", 0, 3.0, 0.5, 0.5)
    force.addParticleParameterOffset("p2", 1, 1.0, 1.0, 2.0)
    force.addExceptionParameterOffset("p1", 1, 0.5, 0.5, 1.5)
    system.addForce(force)
    context = nbs.Context(system, platform)
    context.setPositions([[i, 0, 0] for i in range(4)])
    assert len(context.getParameters()) == 2
    assert context.getParameter("p1") == 0.0 and context.getParameter("p2") == 1.0
    context.setParameter("p1", 0.5)
    context.setParameter("p2", 1.5)
    q = [0.0+3.0*0.5, 1.0+1.0*1.5, -1.0, 0.5]
    sg = [1.0+0.5*0.5, 0.5+1.0*1.5, 2.0, 2.0]
    ep = [0.5+0.5*0.5, 0.6+2.0*1.5, 0.7, 0.8]
    qq, ss, ee = {}, {}, {}
    for i in range(4):
        for j in range(i+1, 4):
            qq[i, j], ss[i, j], ee[i, j] = q[i]*q[j], 0.5*(sg[i]+sg[j]), math.sqrt(ep[i]*ep[j])
    qq[0, 3], ss[0, 3], ee[0, 3] = 0.0, 1.0, 0.0
    qq[2, 3], ss[2, 3], ee[2, 3] = 0.5+0.5*0.5, 1.0+0.5*0.5, 1.5+1.5*0.5
    qq[0, 1], ss[0, 1], ee[0, 1] = 1.0, 1.5, 1.0
    energy = 0.0
    for (i, j) in qq:
        dist = j-i
        x = ss[i, j]/dist
        energy += nbs.ONE_4PI_EPS0*qq[i, j]/dist + 4.0*ee[i, j]*(x**12-x**6)
    assert_equal_tol(energy, context.getState(getEnergy=True).getPotentialEnergy(), 1e-4)


def test_direct_and_reciprocal(nbs, platform):
    """testDirectAndReciprocal :987-1029"""
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
    e4 = context.getState(getEnergy=True).getPotentialEnergy()
    assert_equal_tol(e3, e4, 1e-4)


def test_ewald_exceptions(nbs, platform):
    """testEwaldExceptions :947-985 -- LJPME: adding a periodic exception changes the energy by exactly the
    exception's own Coulomb + LJ energy minus the pair's plain LJ energy (the pair leaves the direct-space sum
    and both reciprocal sums are backed out by the exclusion corrections, Coulomb and dispersion)."""
    system = nbs.System()
    for _ in range(4):
        system.addParticle(1.0)
    system.setDefaultPeriodicBoxVectors([2, 0, 0], [0, 2, 0], [0, 0, 2])
    force = nbs.SlicedNonbondedForce(1)
    system.addForce(force)
    force.setNonbondedMethod(force.LJPME)
    force.setCutoffDistance(1.0)
    force.addParticle(1.0, 0.5, 1.0)
    force.addParticle(1.0, 0.5, 1.0)
    force.addParticle(-1.0, 0.5, 1.0)
    force.addParticle(-1.0, 0.5, 1.0)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [1.5, 0, 0], [0, 0.5, 0.5], [0.2, 1.3, 0]])
    e1 = context.getState(getEnergy=True).getPotentialEnergy()
    force.addException(0, 1, 0.2, 0.8, 2.0)
    force.setExceptionsUsePeriodicBoundaryConditions(True)
    context.reinitialize(True)
    e2 = context.getState(getEnergy=True).getPotentialEnergy()
    r = 0.5
    expected = nbs.ONE_4PI_EPS0*(0.2-1.0)/r + 4*2.0*((0.8/r)**12-(0.8/r)**6) - 4*1.0*((0.5/r)**12-(0.5/r)**6)
    assert_equal_tol(expected, e2-e1, 1e-4)


def test_two_forces(nbs, platform):
    """testTwoForces :815-881 -- two SlicedNonbondedForces in one System (separate kernel instances that must not share
    mutable state), separate force groups, updateParametersInContext on each, then both switched to PME."""
    system = nbs.System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    nb1 = nbs.SlicedNonbondedForce(1)
    nb1.addParticle(-1.5, 1, 1.2)
    nb1.addParticle(0.5, 1, 1.0)
    system.addForce(nb1)
    nb2 = nbs.SlicedNonbondedForce(1)
    nb2.addParticle(0.4, 1.4, 0.5)
    nb2.addParticle(0.3, 1.8, 1.0)
    nb2.setForceGroup(1)
    system.addForce(nb2)
    context = nbs.Context(system, platform)
    context.setPositions([[0, 0, 0], [1.5, 0, 0]])
    K = nbs.ONE_4PI_EPS0
    e1 = context.getState(getEnergy=True, groups=1 << 0).getPotentialEnergy()
    assert_equal_tol(K*(-1.5*0.5)/1.5 + 4.0*math.sqrt(1.2*1.0)*((1.0/1.5)**12-(1.0/1.5)**6), e1, TOL)
    e2 = context.getState(getEnergy=True, groups=1 << 1).getPotentialEnergy()
    assert_equal_tol(K*(0.4*0.3)/1.5 + 4.0*math.sqrt(0.5*1.0)*((1.6/1.5)**12-(1.6/1.5)**6), e2, TOL)
    assert_equal_tol(e1+e2, context.getState(getEnergy=True).getPotentialEnergy(), TOL)
    nb1.setParticleParameters(0, -1.2, 1.1, 1.4)
    nb1.updateParametersInContext(context)
    nb2.setParticleParameters(0, 0.5, 1.6, 0.6)
    nb2.updateParametersInContext(context)
    e1 = context.getState(getEnergy=True, groups=1 << 0).getPotentialEnergy()
    assert_equal_tol(K*(-1.2*0.5)/1.5 + 4.0*math.sqrt(1.4*1.0)*((1.05/1.5)**12-(1.05/1.5)**6), e1, TOL)
    e2 = context.getState(getEnergy=True, groups=1 << 1).getPotentialEnergy()
    assert_equal_tol(K*(0.5*0.3)/1.5 + 4.0*math.sqrt(0.6*1.0)*((1.7/1.5)**12-(1.7/1.5)**6), e2, TOL)
    # (PME runs in the System's default 2 nm box, exactly twice the default cutoff, as in the reference's test)
    nb1.setNonbondedMethod(nb1.PME)
    nb2.setNonbondedMethod(nb2.PME)
    context.reinitialize(True)
    e1 = context.getState(getEnergy=True, groups=1 << 0).getPotentialEnergy()
    e2 = context.getState(getEnergy=True, groups=1 << 1).getPotentialEnergy()
    assert_equal_tol(e1+e2, context.getState(getEnergy=True).getPotentialEnergy(), TOL)
    assert e1 != 0 and e2 != 0 and e1 != e2


SEPARATION_CASES = [(m, x) for m in ("NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "Ewald", "PME", "LJPME") for x in (False, True)]


def scaling_parameter_separation(nbs, platform, method, exceptions, tol):
    """testScalingParameterSeparation :1320-1456 -- one parameter scaling both terms of slice (0,1) against separate
    Coulomb / LJ parameters; parameters on the diagonal slices; E = sum lambda dE/dlambda because every slice is
    scaled; overall, direct-space-only and reciprocal-space-only groups."""
    num_molecules = 100
    n = 2*num_molecules
    cutoff = 3.5
    L = 7.0 if exceptions else 10.0
    rng = np.random.default_rng(17)
    systems, forces = [nbs.System(), nbs.System()], [nbs.SlicedNonbondedForce(2), nbs.SlicedNonbondedForce(2)]
    M = int(round(num_molecules**(1/3)))
    if M*M*M < num_molecules:
        M += 1
    positions = np.zeros((n, 3))
    subset1 = rng.random(n) < 0.5
    for f in forces:
        f.setNonbondedMethod(getattr(f, method))
        f.setCutoffDistance(cutoff)
        f.setUseDispersionCorrection(True)
        f.setReciprocalSpaceForceGroup(1)
        f.setEwaldErrorTolerance(1e-4)
    for k in range(num_molecules):
        iz = k//(M*M)
        iy = (k - iz*M*M)//M
        ix = k - M*(iy + iz*M)
        center = np.array([ix+0.5, iy+0.5, iz+0.5])*L/M
        delta = np.array([0.5-ix % 2, 0.5-iy % 2, 0.5-iz % 2])/2
        i, j = 2*k, 2*k+1
        positions[i], positions[j] = center+delta, center-delta
        for f in forces:
            f.addParticle(1-2*(i % 2), 1, 1)
            f.addParticle(1-2*(j % 2), 1, 1)
            if exceptions:
                f.addException(i, j, float((1-2*(i % 2))*(1-2*(j % 2))), 1, 1)
    for s in systems:
        for _ in range(n):
            s.addParticle(1.0)
        s.setDefaultPeriodicBoxVectors([L, 0, 0], [0, L, 0], [0, 0, L])
    for f in forces:
        for k in range(n):
            if subset1[k]:
                f.setParticleSubset(k, 1)
    lam, value = 0.5, 0.3
    s1, s2 = forces
    s1.addGlobalParameter("lambda", lam)
    s1.addScalingParameter("lambda", 0, 1, True, True)
    s1.addEnergyParameterDerivative("lambda")
    s2.addGlobalParameter("lambdaCoulomb", lam)
    s2.addGlobalParameter("lambdaLJ", lam)
    s2.addScalingParameter("lambdaCoulomb", 0, 1, True, False)
    s2.addScalingParameter("lambdaLJ", 0, 1, False, True)
    s2.addEnergyParameterDerivative("lambdaCoulomb")
    s2.addEnergyParameterDerivative("lambdaLJ")
    s1.addGlobalParameter("alpha", value)
    s1.addScalingParameter("alpha", 0, 0, True, True)
    s1.addEnergyParameterDerivative("alpha")
    s1.addGlobalParameter("beta", value)
    s1.addScalingParameter("beta", 1, 1, True, True)
    s1.addEnergyParameterDerivative("beta")
    s2.addGlobalParameter("gamma", value)
    s2.addScalingParameter("gamma", 0, 0, True, True)
    s2.addScalingParameter("gamma", 1, 1, True, True)
    s2.addEnergyParameterDerivative("gamma")
    systems[0].addForce(s1)
    systems[1].addForce(s2)
    contexts = [nbs.Context(systems[0], platform), nbs.Context(systems[1], platform)]
    for c in contexts:
        c.setPositions(positions)
    for groups in (0xFFFFFFFF, 1 << 0, 1 << 1):
        st1 = contexts[0].getState(getEnergy=True, getForces=True, getParameterDerivatives=True, groups=groups)
        st2 = contexts[1].getState(getEnergy=True, getForces=True, getParameterDerivatives=True, groups=groups)
        d1, d2 = st1.getEnergyParameterDerivatives(), st2.getEnergyParameterDerivatives()
        assert_equal_tol(st1.getPotentialEnergy(), st2.getPotentialEnergy(), tol)
        for fa, fb in zip(st1.getForces(), st2.getForces()):
            assert_equal_vec(fa, fb, tol)
        assert_equal_tol(d1["lambda"], d2["lambdaCoulomb"]+d2["lambdaLJ"], tol)
        assert_equal_tol(st1.getPotentialEnergy(), lam*d1["lambda"]+value*(d1["alpha"]+d1["beta"]), tol)
        assert_equal_tol(d1["alpha"]+d1["beta"], d2["gamma"], tol)


@pytest.mark.parametrize("method,exceptions", SEPARATION_CASES)
def test_scaling_parameter_separation(nbs, platform, method, exceptions):
    scaling_parameter_separation(nbs, platform, method, exceptions, 1e-4)


SLICING_CASES = [(m, o, x, lj) for m in ("NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "Ewald", "PME", "LJPME")
                 for o in (False, True) for x in (False, True) for lj in (False, True)]


def nonbonded_slicing(nbs, platform, method, offsets, exceptions, lj, tol):
    """testNonbondedSlicing :1030-1318, the reference's 48-run matrix (:1493-1497): a two-subset SlicedNonbondedForce
    whose slices (0,1) and (1,1) are scaled by Context parameters must equal an UNSLICED force (here a one-subset
    SlicedNonbondedForce standing in for OpenMM's NonbondedForce) whose subset-1 charges (Coulomb runs) or epsilons (LJ
    runs) are rescaled by hand -- at parameter values 1, 0 and 0.5, for the direct-space group, the reciprocal-space group
    and both; then E(1) - E(0) = sum of the two parameter derivatives, and with a third parameter on slice (0,0) the
    three derivatives add up to the energy of the scaled term.  Offsets: a particle charge offset, a particle epsilon
    offset and (with exceptions) the same on two exceptions, all driven by one global parameter."""
    include_lj, include_coulomb = lj, not lj
    num_molecules = 100
    n = 2*num_molecules
    cutoff = 3.5
    L = 7.0 if exceptions else 10.0
    rng = np.random.default_rng(23)
    system1, system2 = nbs.System(), nbs.System()
    for sysm in (system1, system2):
        for _ in range(n):
            sysm.addParticle(1.0)
        sysm.setDefaultPeriodicBoxVectors([L, 0, 0], [0, L, 0], [0, 0, L])
    plain = nbs.SlicedNonbondedForce(1)
    plain.setNonbondedMethod(getattr(plain, method))
    plain.setCutoffDistance(cutoff)
    plain.setUseDispersionCorrection(True)
    plain.setReciprocalSpaceForceGroup(1)
    plain.setEwaldErrorTolerance(1e-4)

    def q(k):
        return float(1 - 2*(k % 2))
    M = int(round(num_molecules**(1/3)))
    if M*M*M < num_molecules:
        M += 1
    eps = 1.0
    positions = np.zeros((n, 3))
    for k in range(num_molecules):
        iz = k//(M*M)
        iy = (k - iz*M*M)//M
        ix = k - M*(iy + iz*M)
        center = np.array([ix+0.5, iy+0.5, iz+0.5])*L/M
        delta = np.array([0.5-ix % 2, 0.5-iy % 2, 0.5-iz % 2])/2
        i, j = 2*k, 2*k+1
        positions[i], positions[j] = center+delta, center-delta
        plain.addParticle(q(i), 1, eps)
        plain.addParticle(q(j), 1, eps)
        if exceptions:
            plain.addException(i, j, q(i)*q(j), 1, eps)
    particle_offsets, exception_offsets = [], []
    if offsets:
        particle_offsets = [(0, "offsetLambda", 1.0, 0.0, 0.0), (1, "offsetLambda", 0.0, 0.0, 1.0)]
        if exceptions:
            exception_offsets = [(0, "offsetLambda", 1.0, 0.0, 0.0), (1, "offsetLambda", 0.0, 0.0, 1.0)]
        plain.addGlobalParameter("offsetLambda", 0.0)
        for k, name, dq, ds, de in particle_offsets:
            plain.addParticleParameterOffset(name, k, dq, ds, de)
        for k, name, dq, ds, de in exception_offsets:
            plain.addExceptionParameterOffset(name, k, dq, ds, de)
    sliced = nbs.SlicedNonbondedForce(plain, 2)
    in1 = rng.random(n) < 0.5
    for k in range(n):
        if in1[k]:
            sliced.setParticleSubset(k, 1)
    param01 = "lambda" if include_coulomb else "sqrtLambda"
    sliced.addGlobalParameter(param01, 1)
    sliced.addScalingParameter(param01, 0, 1, include_coulomb, include_lj)
    param11 = "lambdaSq" if include_coulomb else "lambda"
    sliced.addGlobalParameter(param11, 1)
    sliced.addScalingParameter(param11, 1, 1, include_coulomb, include_lj)
    system1.addForce(plain)
    system2.addForce(sliced)
    particle_scale = [("lambda" if include_coulomb else "one", "lambda" if include_lj else "one") if in1[k] else ("one", "one") for k in range(n)]
    num_exceptions = plain.getNumExceptions()
    exception_scale = []
    for k in range(num_exceptions):
        i, j = plain.getExceptionParameters(k)[:2]
        if in1[i] != in1[j] or in1[i]:
            name = param01 if in1[i] != in1[j] else param11
            exception_scale.append((name if include_coulomb else "one", name if include_lj else "one"))
        else:
            exception_scale.append(("one", "one"))
    context1, context2 = nbs.Context(system1, platform), nbs.Context(system2, platform)
    context1.setPositions(positions)
    context2.setPositions(positions)

    def compare():
        """direct-space group, reciprocal-space group, everything (:1151-1171); returns the total energy"""
        energy = None
        for groups in (1 << 0, 1 << 1, 0xFFFFFFFF):
            st1 = context1.getState(getEnergy=True, getForces=True, groups=groups)
            st2 = context2.getState(getEnergy=True, getForces=True, groups=groups)
            assert_equal_tol(st1.getPotentialEnergy(), st2.getPotentialEnergy(), tol)
            for fa, fb in zip(st1.getForces(), st2.getForces()):
                assert_equal_vec(fa, fb, tol)
            energy = st1.getPotentialEnergy()
        return energy

    def rescale(value):
        for k in range(n):
            plain.setParticleParameters(k, q(k)*value[particle_scale[k][0]], 1, eps*value[particle_scale[k][1]])
        for k in range(num_exceptions):
            plain.setExceptionParameters(k, 2*k, 2*k+1, q(2*k)*q(2*k+1)*value[exception_scale[k][0]], 1, eps*value[exception_scale[k][1]])
        for idx, (k, name, dq, ds, de) in enumerate(particle_offsets):
            plain.setParticleParameterOffset(idx, name, k, dq*value[particle_scale[k][0]], ds, de*value[particle_scale[k][1]])
        for idx, (k, name, dq, ds, de) in enumerate(exception_offsets):
            plain.setExceptionParameterOffset(idx, name, k, dq*value[exception_scale[k][0]], ds, de*value[exception_scale[k][1]])
        plain.updateParametersInContext(context1)
        context2.setParameter(param01, value[param01])
        context2.setParameter(param11, value[param11])

    energy_lambda_one = compare()
    rescale({"one": 1.0, "lambda": 0.0, "sqrtLambda": 0.0, "lambdaSq": 0.0})
    energy_lambda_zero = compare()
    rescale({"one": 1.0, "lambda": 0.5, "sqrtLambda": np.sqrt(0.5), "lambdaSq": 0.25})
    compare()
    # derivatives (:1279-1286): the energy is linear in each scaling parameter
    sliced.addEnergyParameterDerivative(param01)
    sliced.addEnergyParameterDerivative(param11)
    context2.reinitialize(True)
    derivs = context2.getState(getParameterDerivatives=True).getEnergyParameterDerivatives()
    assert_equal_tol(energy_lambda_one - energy_lambda_zero, derivs[param01] + derivs[param11], tol)
    # sum of derivatives (:1288-1317): only the scaled term left in the unsliced force, every slice of it scaled
    for k in range(n):
        plain.setParticleParameters(k, q(k) if include_coulomb else 0.0, 1, eps if include_lj else 0.0)
    for k in range(num_exceptions):
        plain.setExceptionParameters(k, 2*k, 2*k+1, q(2*k)*q(2*k+1) if include_coulomb else 0.0, 1, eps if include_lj else 0.0)
    for idx, (k, name, dq, ds, de) in enumerate(particle_offsets):
        plain.setParticleParameterOffset(idx, name, k, dq if include_coulomb else 0.0, ds, de if include_lj else 0.0)
    for idx, (k, name, dq, ds, de) in enumerate(exception_offsets):
        plain.setExceptionParameterOffset(idx, name, k, dq if include_coulomb else 0.0, ds, de if include_lj else 0.0)
    plain.updateParametersInContext(context1)
    energy = context1.getState(getEnergy=True).getPotentialEnergy()
    sliced.addGlobalParameter("remainder", 1.0)
    sliced.addScalingParameter("remainder", 0, 0, include_coulomb, include_lj)
    sliced.addEnergyParameterDerivative("remainder")
    context2.reinitialize(True)
    derivs = context2.getState(getEnergy=True, getParameterDerivatives=True).getEnergyParameterDerivatives()
    assert_equal_tol(energy, derivs[param01] + derivs[param11] + derivs["remainder"], tol)


@pytest.mark.parametrize("method,offsets,exceptions,lj", SLICING_CASES)
def test_nonbonded_slicing(nbs, platform, method, offsets, exceptions, lj):
    nonbonded_slicing(nbs, platform, method, offsets, exceptions, lj, 1e-4)


def test_parameter_clash(nbs, platform):
    """python/tests/TestSlicedNonbondedForce.py:51-67 and SlicedNonbondedForceImpl.cpp:114-131"""
    system = nbs.System()
    system.setDefaultPeriodicBoxVectors([4, 0, 0], [0, 4, 0], [0, 0, 4])
    system.addParticle(1.0)
    system.addParticle(1.0)
    force = nbs.SlicedNonbondedForce(1)
    force.addParticle(1.5, 1, 0)
    force.addParticle(-1.5, 1, 0)
    force.addGlobalParameter("param", 1)
    force.addScalingParameter("param", 0, 0, True, True)
    force.addParticleParameterOffset("param", 0, 1, 1, 0)
    system.addForce(force)
    with pytest.raises(Exception):
        nbs.Context(system, platform)


# ---- the reference tests that compare against OpenMM's own NonbondedForce / Reference platform -------------------------
# (NonbondedForce does not exist here: a one-subset SlicedNonbondedForce stands in for it, and "the Reference platform" is
# the compiled-reference oracle; the bodies take the two platforms to compare, so the GPU suite re-uses them)
def _compare_states(context_a, context_b, groups, etol, ftol):
    from helpers import force_rel_rms
    sa = context_a.getState(getEnergy=True, getForces=True, groups=groups)
    sb = context_b.getState(getEnergy=True, getForces=True, groups=groups)
    assert_equal_tol(sb.getPotentialEnergy(), sa.getPotentialEnergy(), etol)
    assert force_rel_rms(sa.getForces(), sb.getForces()) <= ftol


def large_system(nbs, platform_a, platform_b, etol=1e-5, ftol=1e-5):
    """testLargeSystem :494-555 -- 600 two-particle molecules at random positions in a 20 nm box, an exclusion inside each
    molecule; NoCutoff, then CutoffNonPeriodic (2 nm), then CutoffPeriodic after reinitialize(true): `platform_a` against
    `platform_b` (the reference compares its platform with the Reference platform).  (Its HarmonicBondForce sits in
    another force group and is not part of the comparison.)"""
    num_molecules, cutoff, box = 600, 2.0, 20.0
    rng = np.random.default_rng(0)
    system = nbs.System()
    force = nbs.SlicedNonbondedForce(1)
    positions = np.zeros((2*num_molecules, 3))
    for i in range(num_molecules):
        eps = 0.1 if i < num_molecules//2 else 0.2
        force.addParticle(-1.0, 0.2, eps)
        force.addParticle(1.0, 0.1, eps)
        positions[2*i] = box*rng.random(3)
        positions[2*i+1] = positions[2*i] + [1.0, 0.0, 0.0]
        force.addException(2*i, 2*i+1, 0.0, 0.15, 0.0)
        system.addParticle(1.0)
        system.addParticle(1.0)
    system.setDefaultPeriodicBoxVectors([box, 0, 0], [0, box, 0], [0, 0, box])
    force.setNonbondedMethod(force.NoCutoff)
    system.addForce(force)
    contexts = [nbs.Context(system, platform_a), nbs.Context(system, platform_b)]
    for c in contexts:
        c.setPositions(positions)
    _compare_states(contexts[0], contexts[1], 0xFFFFFFFF, etol, ftol)
    force.setNonbondedMethod(force.CutoffNonPeriodic)
    force.setCutoffDistance(cutoff)
    for c in contexts:
        c.reinitialize(True)
    _compare_states(contexts[0], contexts[1], 0xFFFFFFFF, etol, ftol)
    force.setNonbondedMethod(force.CutoffPeriodic)
    for c in contexts:
        c.reinitialize(True)
    _compare_states(contexts[0], contexts[1], 0xFFFFFFFF, etol, ftol)


def changing_parameters(nbs, platform_a, platform_b, etol=1e-5, ftol=1e-5):
    """testChangingParameters :683-758 -- 600 molecules on a lattice, PME with a 2 nm cutoff in a 20 nm box (default PME
    parameters), direct-space and reciprocal-space groups compared separately; then every fifth particle gets
    1.5 x charge, 1.1 x sigma, 1.7 x epsilon through updateParametersInContext and everything is compared again."""
    num_molecules, cutoff, box = 600, 2.0, 20.0
    system = nbs.System()
    force = nbs.SlicedNonbondedForce(1)
    M = int(num_molecules**(1/3))
    if M*M*M < num_molecules:
        M += 1
    positions = np.zeros((2*num_molecules, 3))
    for k in range(num_molecules):
        iz = k//(M*M)
        iy = (k - iz*M*M)//M
        ix = k - M*(iy + iz*M)
        center = (np.array([ix, iy, iz]) + 0.5)*box/M
        delta = np.array([0.5 - ix % 2, 0.5 - iy % 2, 0.5 - iz % 2])/2
        eps = 0.1 if k < num_molecules//2 else 0.2
        force.addParticle(-1.0, 0.2, eps)
        force.addParticle(1.0, 0.1, eps)
        positions[2*k], positions[2*k+1] = center + delta, center - delta
        force.addException(2*k, 2*k+1, 0.0, 0.15, 0.0)
        system.addParticle(1.0)
        system.addParticle(1.0)
    force.setNonbondedMethod(force.PME)
    force.setCutoffDistance(cutoff)
    force.setForceGroup(1)
    force.setReciprocalSpaceForceGroup(3)
    system.addForce(force)
    system.setDefaultPeriodicBoxVectors([box, 0, 0], [0, box, 0], [0, 0, box])
    contexts = [nbs.Context(system, platform_a), nbs.Context(system, platform_b)]
    for c in contexts:
        c.setPositions(positions)
    _compare_states(contexts[0], contexts[1], 1 << 1, etol, ftol)
    _compare_states(contexts[0], contexts[1], 1 << 3, etol, ftol)
    for i in range(0, 2*num_molecules, 5):
        charge, sigma, epsilon = force.getParticleParameters(i)
        force.setParticleParameters(i, 1.5*charge, 1.1*sigma, 1.7*epsilon)
    for c in contexts:
        force.updateParametersInContext(c)
    _compare_states(contexts[0], contexts[1], 0xFFFFFFFF, etol, ftol)
    _compare_states(contexts[0], contexts[1], 1 << 1, etol, ftol)
    _compare_states(contexts[0], contexts[1], 1 << 3, etol, ftol)


def instantiate_from_nonbonded_force(nbs, platform, method, tol=TOL):
    """testInstantiateFromNonbondedForce :29-85 -- a force and its copy (the copy constructor the plugin uses to wrap an
    existing NonbondedForce: particles, exceptions, global parameters, particle and exception offsets) in ONE Context,
    in different force groups: direct-space groups agree, and after a parameter change the reciprocal-space groups do."""
    force = nbs.SlicedNonbondedForce(1)
    force.setCutoffDistance(2.0)
    force.setNonbondedMethod(getattr(force, method))
    for q, sig, eps in ((0.0, 1.0, 0.5), (1.0, 0.5, 0.6), (-1.0, 2.0, 0.7), (0.5, 2.0, 0.8), (-0.5, 2.0, 0.8)):
        force.addParticle(q, sig, eps)
    force.addException(0, 3, 0.0, 1.0, 0.0)
    force.addException(2, 3, 0.5, 1.0, 1.5)
    force.addException(0, 1, 1.0, 1.5, 1.0)
    force.addGlobalParameter("p1", 0.5)
    force.addGlobalParameter("p2", 1.0)
    force.addParticleParameterOffset("p1", 0, -2.0, 0.5, 0.5)
    force.addParticleParameterOffset("p2", 1, 1.0, 1.0, 2.0)
    force.addExceptionParameterOffset("p1", 1, 0.5, 0.5, 1.5)
    force.setReciprocalSpaceForceGroup(2)
    sliced = nbs.SlicedNonbondedForce(force, 1)
    sliced.setForceGroup(1)
    sliced.setReciprocalSpaceForceGroup(3)
    n = force.getNumParticles()
    system = nbs.System()
    L = float(n)
    system.setDefaultPeriodicBoxVectors([L, 0, 0], [0, L, 0], [0, 0, L])
    for _ in range(n):
        system.addParticle(1.0)
    system.addForce(force)
    system.addForce(sliced)
    context = nbs.Context(system, platform)
    context.setPositions(np.array([[float(i), 0.0, 0.0] for i in range(n)]))
    for ga, gb, p1 in ((1 << 0, 1 << 1, None), (1 << 2, 1 << 3, 1.0)):
        if p1 is not None:
            context.setParameter("p1", p1)
        sa = context.getState(getEnergy=True, getForces=True, groups=ga)
        sb = context.getState(getEnergy=True, getForces=True, groups=gb)
        assert_equal_tol(sa.getPotentialEnergy(), sb.getPotentialEnergy(), tol)
        for fa, fb in zip(sa.getForces(), sb.getForces()):
            assert_equal_vec(fa, fb, tol)


def test_large_system_port_vs_compiled_reference(nbs, oracle):
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    large_system(nbs, oracle.OraclePlatform("port"), oracle.OraclePlatform("reference"), 1e-10, 1e-10)


def test_changing_parameters_port_vs_compiled_reference(nbs, oracle):
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    changing_parameters(nbs, oracle.OraclePlatform("port"), oracle.OraclePlatform("reference"), 1e-9, 1e-9)


@pytest.mark.parametrize("method", ["NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "Ewald", "PME", "LJPME"])
def test_instantiate_from_nonbonded_force(nbs, platform, method):
    instantiate_from_nonbonded_force(nbs, platform, method)

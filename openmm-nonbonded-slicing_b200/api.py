"""Host-side mirror of the reference's operator interface for the SlicedNonbondedForce hot path.

Same names, argument meaning and error behaviour as the plugin, so the parity tests read like
the reference's own tests:

* ``SlicedNonbondedForce``        <- openmmapi/include/SlicedNonbondedForce.h:26-70 (+ the inherited
                                     OpenMM ``NonbondedForce`` accessors it uses [external])
* ``SlicedNonbondedForceImpl``    <- openmmapi/src/SlicedNonbondedForceImpl.cpp:33-148, 263-367
* ``CalcSlicedNonbondedForceKernel`` (interface) <- openmmapi/include/NonbondedSlicingKernels.h:27-85
* ``B200CalcSlicedNonbondedForceKernel`` -- the implementation over the C ABI (include/nbslice_b200.h)
* ``System`` / ``Context`` / ``State`` -- the minimum of OpenMM [external] needed to drive a force

The production drop-in is the C++ kernel under ``platform/``; this Python mirror exists because the
reference also ships a Python API (python/nonbondedslicing.i) and because it is what the tests and
``bench.py`` call.  All arithmetic happens in the CUDA library; nothing here computes forces.
"""
import ctypes as C
import math

import numpy as np

from . import abi

ONE_4PI_EPS0 = 138.93545764438198  # openmm/reference/SimTKOpenMMRealType.h [external], CODATA-2018


def sliceIndex(i, j):
    """openmmapi/include/SlicedNonbondedForce.h:22"""
    return i*(i+1)//2 + j if i > j else j*(j+1)//2 + i


class OpenMMException(Exception):
    pass


class System:
    def __init__(self):
        self._masses = []
        self._box = np.diag([2.0, 2.0, 2.0])
        self._forces = []

    def addParticle(self, mass):
        self._masses.append(float(mass))
        return len(self._masses)-1

    def getNumParticles(self):
        return len(self._masses)

    def setDefaultPeriodicBoxVectors(self, a, b, c):
        self._box = np.array([list(a), list(b), list(c)], dtype=np.float64)

    def getDefaultPeriodicBoxVectors(self):
        return self._box.copy()

    def addForce(self, force):
        self._forces.append(force)
        return len(self._forces)-1

    def getNumForces(self):
        return len(self._forces)

    def getForce(self, index):
        return self._forces[index]

    def usesPeriodicBoundaryConditions(self):
        return any(f.usesPeriodicBoundaryConditions() for f in self._forces)


class SlicedNonbondedForce:
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)
    _methodNames = ["NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "Ewald", "PME", "LJPME"]

    def __init__(self, *args):
        """SlicedNonbondedForce(numSubsets) or SlicedNonbondedForce(force, numSubsets)
        (openmmapi/src/SlicedNonbondedForce.cpp:28-82)."""
        if len(args) == 1:
            source, numSubsets = None, args[0]
        else:
            source, numSubsets = args
        self.numSubsets = int(numSubsets)
        self._particles = []          # [charge, sigma, epsilon]
        self._exceptions = []         # [p1, p2, chargeProd, sigma, epsilon]
        self._exceptionMap = {}
        self._subsets = {}
        self._globalParams = []       # [name, default]
        self._particleOffsets = []    # [paramIndex, particle, chargeScale, sigmaScale, epsilonScale]
        self._exceptionOffsets = []
        self._scalingParameters = []  # [paramIndex, subset1, subset2, includeCoulomb, includeLJ]
        self._derivatives = []        # global parameter indices
        self.nonbondedMethod = self.NoCutoff
        self.cutoffDistance = 1.0
        self.switchingDistance = -1.0
        self.useSwitchingFunction = False
        self.rfDielectric = 78.3
        self.ewaldErrorTol = 5e-4
        self.alpha, self.nx, self.ny, self.nz = 0.0, 0, 0, 0
        self.dalpha, self.dnx, self.dny, self.dnz = 0.0, 0, 0, 0
        self.useDispersionCorrection = True
        self.exceptionsUsePeriodic = False
        self.includeDirectSpace = True
        self.forceGroup = 0
        self.recipForceGroup = -1
        self.useCuFFT = False
        if source is not None:
            for name in ("nonbondedMethod", "cutoffDistance", "switchingDistance", "useSwitchingFunction",
                         "rfDielectric", "ewaldErrorTol", "alpha", "nx", "ny", "nz", "dalpha", "dnx", "dny", "dnz",
                         "useDispersionCorrection", "exceptionsUsePeriodic", "includeDirectSpace", "forceGroup",
                         "recipForceGroup"):
                setattr(self, name, getattr(source, name))
            self._particles = [list(p) for p in source._particles]
            self._exceptions = [list(e) for e in source._exceptions]
            self._exceptionMap = dict(source._exceptionMap)
            self._globalParams = [list(g) for g in source._globalParams]
            self._particleOffsets = [list(o) for o in source._particleOffsets]
            self._exceptionOffsets = [list(o) for o in source._exceptionOffsets]

    # ---- NonbondedForce accessors [external API, same semantics] -------------------------------
    def getNumParticles(self):
        return len(self._particles)

    def getNumExceptions(self):
        return len(self._exceptions)

    def getNumGlobalParameters(self):
        return len(self._globalParams)

    def getNumParticleParameterOffsets(self):
        return len(self._particleOffsets)

    def getNumExceptionParameterOffsets(self):
        return len(self._exceptionOffsets)

    def getNonbondedMethod(self):
        return self.nonbondedMethod

    def setNonbondedMethod(self, method):
        if method < 0 or method > 5:
            raise OpenMMException("NonbondedForce: Illegal value for nonbonded method")
        self.nonbondedMethod = method

    def getNonbondedMethodName(self):
        return self._methodNames[self.nonbondedMethod]

    def getCutoffDistance(self):
        return self.cutoffDistance

    def setCutoffDistance(self, distance):
        self.cutoffDistance = float(distance)

    def getUseSwitchingFunction(self):
        return self.useSwitchingFunction

    def setUseSwitchingFunction(self, use):
        self.useSwitchingFunction = bool(use)

    def getSwitchingDistance(self):
        return self.switchingDistance

    def setSwitchingDistance(self, distance):
        self.switchingDistance = float(distance)

    def getReactionFieldDielectric(self):
        return self.rfDielectric

    def setReactionFieldDielectric(self, dielectric):
        self.rfDielectric = float(dielectric)

    def getEwaldErrorTolerance(self):
        return self.ewaldErrorTol

    def setEwaldErrorTolerance(self, tol):
        self.ewaldErrorTol = float(tol)

    def getPMEParameters(self):
        return self.alpha, self.nx, self.ny, self.nz

    def setPMEParameters(self, alpha, nx, ny, nz):
        self.alpha, self.nx, self.ny, self.nz = float(alpha), int(nx), int(ny), int(nz)

    def getLJPMEParameters(self):
        return self.dalpha, self.dnx, self.dny, self.dnz

    def setLJPMEParameters(self, alpha, nx, ny, nz):
        self.dalpha, self.dnx, self.dny, self.dnz = float(alpha), int(nx), int(ny), int(nz)

    def getUseDispersionCorrection(self):
        return self.useDispersionCorrection

    def setUseDispersionCorrection(self, use):
        self.useDispersionCorrection = bool(use)

    def getExceptionsUsePeriodicBoundaryConditions(self):
        return self.exceptionsUsePeriodic

    def setExceptionsUsePeriodicBoundaryConditions(self, periodic):
        self.exceptionsUsePeriodic = bool(periodic)

    def getIncludeDirectSpace(self):
        return self.includeDirectSpace

    def setIncludeDirectSpace(self, include):
        self.includeDirectSpace = bool(include)

    def getForceGroup(self):
        return self.forceGroup

    def setForceGroup(self, group):
        if group < 0 or group > 31:
            raise OpenMMException("Force group must be between 0 and 31")
        self.forceGroup = int(group)

    def getReciprocalSpaceForceGroup(self):
        return self.recipForceGroup

    def setReciprocalSpaceForceGroup(self, group):
        if group < -1 or group > 31:
            raise OpenMMException("Force group must be between -1 and 31")
        self.recipForceGroup = int(group)

    def usesPeriodicBoundaryConditions(self):
        return self.nonbondedMethod in (self.CutoffPeriodic, self.Ewald, self.PME, self.LJPME)

    def addParticle(self, charge, sigma, epsilon):
        self._particles.append([float(charge), float(sigma), float(epsilon)])
        return len(self._particles)-1

    def getParticleParameters(self, index):
        return tuple(self._particles[index])

    def setParticleParameters(self, index, charge, sigma, epsilon):
        self._particles[index] = [float(charge), float(sigma), float(epsilon)]

    def addException(self, particle1, particle2, chargeProd, sigma, epsilon, replace=False):
        key = (min(particle1, particle2), max(particle1, particle2))
        if key in self._exceptionMap:
            if not replace:
                raise OpenMMException(
                    f"NonbondedForce: There is already an exception for particles {particle1} and {particle2}")
            index = self._exceptionMap[key]
            self._exceptions[index] = [particle1, particle2, float(chargeProd), float(sigma), float(epsilon)]
            return index
        self._exceptions.append([int(particle1), int(particle2), float(chargeProd), float(sigma), float(epsilon)])
        self._exceptionMap[key] = len(self._exceptions)-1
        return len(self._exceptions)-1

    def getExceptionParameters(self, index):
        return tuple(self._exceptions[index])

    def setExceptionParameters(self, index, particle1, particle2, chargeProd, sigma, epsilon):
        self._exceptions[index] = [int(particle1), int(particle2), float(chargeProd), float(sigma), float(epsilon)]

    def createExceptionsFromBonds(self, bonds, coulomb14Scale, lj14Scale):
        """OpenMM NonbondedForce::createExceptionsFromBonds [external]: 1-2 and 1-3 pairs become pure
        exclusions, 1-4 pairs get scaled Coulomb / Lorentz-Berthelot LJ parameters."""
        n = self.getNumParticles()
        for b in bonds:
            if b[0] < 0 or b[1] < 0 or b[0] >= n or b[1] >= n:
                raise OpenMMException("createExceptionsFromBonds: Illegal particle index in list of bonds")
        bonded12 = [set() for _ in range(n)]
        for a, b in bonds:
            bonded12[a].add(b)
            bonded12[b].add(a)
        exclusions = [set() for _ in range(n)]
        for i in range(n):
            frontier = {i}
            seen = {i}
            for _ in range(3):
                frontier = {k for j in frontier for k in bonded12[j]} - seen
                seen |= frontier
            exclusions[i] = seen - {i}
        for i in range(n):
            bonded13 = set()
            for j in bonded12[i]:
                bonded13 |= bonded12[j]
            for j in sorted(exclusions[i]):
                if j < i:
                    if j not in bonded13 and j not in bonded12[i]:     # a 1-4 interaction
                        q1, s1, e1 = self._particles[j]
                        q2, s2, e2 = self._particles[i]
                        self.addException(j, i, coulomb14Scale*q1*q2, 0.5*(s1+s2), lj14Scale*math.sqrt(e1*e2))
                    else:
                        self.addException(j, i, 0.0, 1.0, 0.0)

    def addGlobalParameter(self, name, defaultValue):
        self._globalParams.append([name, float(defaultValue)])
        return len(self._globalParams)-1

    def getGlobalParameterName(self, index):
        return self._globalParams[index][0]

    def getGlobalParameterDefaultValue(self, index):
        return self._globalParams[index][1]

    def setGlobalParameterDefaultValue(self, index, value):
        self._globalParams[index][1] = float(value)

    def _getGlobalParameterIndex(self, parameter):
        for i, (name, _) in enumerate(self._globalParams):
            if name == parameter:
                return i
        raise OpenMMException(f"There is no global parameter called '{parameter}'")

    def addParticleParameterOffset(self, parameter, particleIndex, chargeScale, sigmaScale, epsilonScale):
        self._particleOffsets.append([self._getGlobalParameterIndex(parameter), int(particleIndex),
                                      float(chargeScale), float(sigmaScale), float(epsilonScale)])
        return len(self._particleOffsets)-1

    def getParticleParameterOffset(self, index):
        o = self._particleOffsets[index]
        return (self._globalParams[o[0]][0], o[1], o[2], o[3], o[4])

    def setParticleParameterOffset(self, index, parameter, particleIndex, chargeScale, sigmaScale, epsilonScale):
        self._particleOffsets[index] = [self._getGlobalParameterIndex(parameter), int(particleIndex),
                                        float(chargeScale), float(sigmaScale), float(epsilonScale)]

    def addExceptionParameterOffset(self, parameter, exceptionIndex, chargeProdScale, sigmaScale, epsilonScale):
        self._exceptionOffsets.append([self._getGlobalParameterIndex(parameter), int(exceptionIndex),
                                       float(chargeProdScale), float(sigmaScale), float(epsilonScale)])
        return len(self._exceptionOffsets)-1

    def getExceptionParameterOffset(self, index):
        o = self._exceptionOffsets[index]
        return (self._globalParams[o[0]][0], o[1], o[2], o[3], o[4])

    def setExceptionParameterOffset(self, index, parameter, exceptionIndex, chargeProdScale, sigmaScale, epsilonScale):
        self._exceptionOffsets[index] = [self._getGlobalParameterIndex(parameter), int(exceptionIndex),
                                         float(chargeProdScale), float(sigmaScale), float(epsilonScale)]

    # ---- SlicedNonbondedForce proper (openmmapi/src/SlicedNonbondedForce.cpp:84-194) ------------
    def getNumSubsets(self):
        return self.numSubsets

    def getNumSlices(self):
        return self.numSubsets*(self.numSubsets+1)//2

    def getNumScalingParameters(self):
        return len(self._scalingParameters)

    def getNumEnergyParameterDerivatives(self):
        return len(self._derivatives)

    def setParticleSubset(self, index, subset):
        if not 0 <= index < self.getNumParticles():
            raise OpenMMException("Index out of range")
        if not 0 <= subset < self.numSubsets:
            raise OpenMMException("Subset out of range")
        self._subsets[index] = int(subset)

    def getParticleSubset(self, index):
        if not 0 <= index < self.getNumParticles():
            raise OpenMMException("Index out of range")
        return self._subsets.get(index, 0)

    def addScalingParameter(self, parameter, subset1, subset2, includeCoulomb, includeLJ):
        if not (includeCoulomb or includeLJ):
            raise OpenMMException("Keywords 'includeCoulomb' and 'includeLJ' cannot be both false")
        for s in (subset1, subset2):
            if not 0 <= s < self.numSubsets:
                raise OpenMMException("Subset out of range")
        info = [self._getGlobalParameterIndex(parameter), int(subset1), int(subset2), bool(includeCoulomb), bool(includeLJ)]
        for other in self._scalingParameters:
            if sliceIndex(other[1], other[2]) == sliceIndex(subset1, subset2) and \
                    ((other[3] and includeCoulomb) or (other[4] and includeLJ)):
                raise OpenMMException("A scaling parameter has already been defined for this slice & contribution(s)")
        self._scalingParameters.append(info)
        return len(self._scalingParameters)-1

    def getScalingParameter(self, index):
        p = self._scalingParameters[index]
        return (self._globalParams[p[0]][0], p[1], p[2], p[3], p[4])

    def setScalingParameter(self, index, parameter, subset1, subset2, includeCoulomb, includeLJ):
        if not (includeCoulomb or includeLJ):
            raise OpenMMException("Keywords 'includeCoulomb' and 'includeLJ' cannot be both false")
        info = [self._getGlobalParameterIndex(parameter), int(subset1), int(subset2), bool(includeCoulomb), bool(includeLJ)]
        for k, other in enumerate(self._scalingParameters):
            if k != index and sliceIndex(other[1], other[2]) == sliceIndex(subset1, subset2) and \
                    ((other[3] and includeCoulomb) or (other[4] and includeLJ)):
                raise OpenMMException("A scaling parameter has already been defined for this slice & contribution(s)")
        self._scalingParameters[index] = info

    def addEnergyParameterDerivative(self, parameter):
        index = None
        for k, p in enumerate(self._scalingParameters):
            if self._globalParams[p[0]][0] == parameter:
                index = k
        if index is None:
            raise OpenMMException(f"There is no scaling parameter called '{parameter}'")
        if index not in self._derivatives:
            self._derivatives.append(index)
        return self._derivatives.index(index)

    def getEnergyParameterDerivativeName(self, index):
        return self._globalParams[self._scalingParameters[self._derivatives[index]][0]][0]

    def getUseCuFFT(self):
        return self.useCuFFT

    def setUseCuFFT(self, use):
        self.useCuFFT = bool(use)

    def getPMEParametersInContext(self, context):
        return context._impl(self).getPMEParameters()

    def getLJPMEParametersInContext(self, context):
        """openmmapi/include/SlicedNonbondedForce.h:31, openmmapi/src/SlicedNonbondedForce.cpp:188-190"""
        return context._impl(self).getLJPMEParameters()

    def updateParametersInContext(self, context):
        context._impl(self).updateParametersInContext(context)


class SlicedNonbondedForceImpl:
    """openmmapi/src/SlicedNonbondedForceImpl.cpp"""

    def __init__(self, owner):
        self.owner = owner
        self.kernel = None

    def initialize(self, context):
        owner = self.owner
        system = context.getSystem()
        self.kernel = context.getPlatform().createKernel(CalcSlicedNonbondedForceKernel.Name(), context)
        if owner.getNumParticles() != system.getNumParticles():
            raise OpenMMException("SlicedNonbondedForce must have exactly as many particles as the System it belongs to.")
        if owner.getUseSwitchingFunction():
            if owner.getSwitchingDistance() < 0 or owner.getSwitchingDistance() >= owner.getCutoffDistance():
                raise OpenMMException("SlicedNonbondedForce: Switching distance must satisfy 0 <= r_switch < r_cutoff")
        for charge, sigma, epsilon in owner._particles:
            if sigma < 0:
                raise OpenMMException("SlicedNonbondedForce: sigma for a particle cannot be negative")
            if epsilon < 0:
                raise OpenMMException("SlicedNonbondedForce: epsilon for a particle cannot be negative")
        seen = set()
        n = owner.getNumParticles()
        for p1, p2, chargeProd, sigma, epsilon in owner._exceptions:
            for p in (p1, p2):
                if p < 0 or p >= n:
                    raise OpenMMException(f"SlicedNonbondedForce: Illegal particle index for an exception: {p}")
            key = (min(p1, p2), max(p1, p2))
            if key in seen:
                raise OpenMMException(f"SlicedNonbondedForce: Multiple exceptions are specified for particles {p1} and {p2}")
            seen.add(key)
            if sigma < 0:
                raise OpenMMException("SlicedNonbondedForce: sigma for an exception cannot be negative")
            if epsilon < 0:
                raise OpenMMException("SlicedNonbondedForce: epsilon for an exception cannot be negative")
        for o in owner._particleOffsets:
            if o[1] < 0 or o[1] >= n:
                raise OpenMMException(f"SlicedNonbondedForce: Illegal particle index for a particle parameter offset: {o[1]}")
        for o in owner._exceptionOffsets:
            if o[1] < 0 or o[1] >= owner.getNumExceptions():
                raise OpenMMException(f"SlicedNonbondedForce: Illegal exception index for an exception parameter offset: {o[1]}")
        if owner.usesPeriodicBoundaryConditions():
            box = system.getDefaultPeriodicBoxVectors()
            cutoff = owner.getCutoffDistance()
            if cutoff > 0.5*box[0][0] or cutoff > 0.5*box[1][1] or cutoff > 0.5*box[2][2]:
                raise OpenMMException("SlicedNonbondedForce: The cutoff distance cannot be greater than half the periodic box size.")
            if owner.getNonbondedMethod() == owner.Ewald and (box[1][0] != 0.0 or box[2][0] != 0.0 or box[2][1] != 0):
                raise OpenMMException("SlicedNonbondedForce: Ewald is not supported with non-rectangular boxes.  Use PME instead.")
        offsetParams = {o[0] for o in owner._particleOffsets} | {o[0] for o in owner._exceptionOffsets}
        for p in owner._scalingParameters:
            if p[0] in offsetParams:
                raise OpenMMException("SlicedNonbondedForce: Cannot use a global parameter for both slice energy scaling and parameter offset.")
        self.kernel.initialize(system, owner)

    def getDefaultParameters(self):
        return {name: value for name, value in self.owner._globalParams}

    def calcForcesAndEnergy(self, context, includeForces, includeEnergy, groups):
        """:135-142"""
        owner = self.owner
        includeDirect = owner.getIncludeDirectSpace() and (groups & (1 << owner.getForceGroup())) != 0
        reciprocalGroup = owner.getReciprocalSpaceForceGroup()
        if reciprocalGroup < 0:
            reciprocalGroup = owner.getForceGroup()
        includeReciprocal = (groups & (1 << reciprocalGroup)) != 0
        return self.kernel.execute(context, includeForces, includeEnergy, includeDirect, includeReciprocal)

    def updateParametersInContext(self, context):
        self.kernel.copyParametersToContext(context, self.owner)

    def getPMEParameters(self):
        return self.kernel.getPMEParameters()

    def getLJPMEParameters(self):
        """openmmapi/src/SlicedNonbondedForceImpl.cpp:365-367"""
        return self.kernel.getLJPMEParameters()

    @staticmethod
    def calcPMEParameters(system, force, lj=False):
        """OpenMM NonbondedForceImpl::calcPMEParameters [external] (formula for alpha and the
        default grid); explicit setPMEParameters values win."""
        alpha, nx, ny, nz = force.getLJPMEParameters() if lj else force.getPMEParameters()
        if alpha == 0.0:
            box = system.getDefaultPeriodicBoxVectors()
            tol = force.getEwaldErrorTolerance()
            alpha = math.sqrt(-math.log(2*tol))/force.getCutoffDistance()
            if lj:
                nx, ny, nz = (int(math.ceil(alpha*box[k][k]/(3*pow(tol, 0.2)))) for k in range(3))
            else:
                nx, ny, nz = (int(math.ceil(2*alpha*box[k][k]/(3*pow(tol, 0.2)))) for k in range(3))
            nx, ny, nz = max(nx, 6), max(ny, 6), max(nz, 6)
        return alpha, nx, ny, nz

    @staticmethod
    def calcEwaldParameters(system, force):
        """OpenMM NonbondedForceImpl::calcEwaldParameters [external], called at
        ReferenceNonbondedSlicingKernels.cpp:160-162: alpha from the error tolerance and, per axis, the
        smallest number of reciprocal vectors whose truncation error estimate
        ``0.05 sqrt(L alpha) k exp(-(pi k / (L alpha))^2)`` falls below the tolerance, made odd."""
        box = system.getDefaultPeriodicBoxVectors()
        tol = force.getEwaldErrorTolerance()
        alpha = math.sqrt(-math.log(2*tol))/force.getCutoffDistance()

        def find_zero(width):
            def value(arg):
                temp = arg*math.pi/(width*alpha)
                return tol - 0.05*math.sqrt(width*alpha)*arg*math.exp(-temp*temp)
            arg = 10
            v = value(arg)
            if v > 0.0:
                while v > 0.0 and arg > 0:
                    arg -= 1
                    v = value(arg)
                return arg+1
            while v < 0.0:
                arg += 1
                v = value(arg)
            return arg

        kmax = [find_zero(box[k][k]) for k in range(3)]
        kmax = [k+1 if k % 2 == 0 else k for k in kmax]
        return alpha, kmax[0], kmax[1], kmax[2]

    @staticmethod
    def _evalIntegral(r, rs, rc, sigma):
        """:150-185"""
        A = 1/(rc-rs)
        A2 = A*A
        A3 = A2*A
        sig6 = sigma**6
        rs2, rs3 = rs*rs, rs**3
        poly12 = (rs3*28*(6*rs2*A2 + 15*rs*A + 10) - r*rs2*945*(rs2*A2 + 2*rs*A + 1)
                  + r**2*rs*1080*(2*rs2*A2 + 3*rs*A + 1) - r**3*420*(6*rs2*A2 + 6*rs*A + 1)
                  + r**4*756*(2*rs*A2 + A) - r**5*378*A2)
        poly6 = (rs3*84*(6*rs2*A2 + 15*rs*A + 10) - r*rs2*3780*(rs2*A2 + 2*rs*A + 1)
                 + r**2*rs*7560*(2*rs2*A2 + 3*rs*A + 1))
        return sig6*A3*((sig6*poly12 - r**6*poly6)/(252*r**9)
                        - math.log(r)*10*(6*rs2*A2 + 6*rs*A + 1) + r*15*(2*rs*A2 + A) - r*r*3*A2)

    @staticmethod
    def calcDispersionCorrections(system, force):
        """:263-354.  The reference does its pair counting in 32-bit ``int`` (``count``,
        ``numParticles*(numParticles+1)``, ``8*numParticles*numParticles``); that wrap-around is
        reproduced so that the coefficients equal what the unchanged API library hands
        the C++ adapter."""
        def i32(x):
            x &= 0xFFFFFFFF
            return x - (1 << 32) if x >= (1 << 31) else x

        def cdiv2(x):          # C integer division by 2 truncates toward zero
            return -((-x)//2) if x < 0 else x//2

        numSlices = force.getNumSlices()
        result = [0.0]*numSlices
        if force.getNonbondedMethod() in (force.NoCutoff, force.CutoffNonPeriodic):
            return result
        n = system.getNumParticles()
        sigma = [p[1] for p in force._particles]
        epsilon = [p[2] for p in force._particles]
        defaults = [g[1] for g in force._globalParams]
        for o in force._particleOffsets:
            sigma[o[1]] += defaults[o[0]]*o[3]
            epsilon[o[1]] += defaults[o[0]]*o[4]
        classCounts = {}
        for i in range(force.getNumParticles()):
            key = (sigma[i], epsilon[i], force.getParticleSubset(i))
            classCounts[key] = classCounts.get(key, 0)+1
        classes = sorted(classCounts.items())      # std::map order
        sum1, sum2, sum3 = [0.0]*numSlices, [0.0]*numSlices, [0.0]*numSlices
        useSwitch = force.getUseSwitchingFunction()
        cutoff, switchDist = force.getCutoffDistance(), force.getSwitchingDistance()
        ev = SlicedNonbondedForceImpl._evalIntegral

        def add(slice_, count, sig, eps):
            sig6 = sig**2
            sig6 = sig6*sig6*sig6
            sum1[slice_] += count*eps*sig6*sig6
            sum2[slice_] += count*eps*sig6
            if useSwitch:
                sum3[slice_] += count*eps*(ev(cutoff, switchDist, cutoff, sig)-ev(switchDist, switchDist, cutoff, sig))

        for (sig, eps, subset), cnt in classes:
            add(subset*(subset+3)//2, cdiv2(i32(cnt*(cnt+1))), sig, eps)
        for a, ((sig1, eps1, s1), cnt1) in enumerate(classes):
            for (sig2, eps2, s2), cnt2 in classes[:a]:
                add(sliceIndex(s1, s2), i32(cnt1*cnt2), 0.5*(sig1+sig2), math.sqrt(eps1*eps2))
        numInteractions = float(cdiv2(i32(n*(n+1))))
        prefactor = i32(i32(8*n)*n)
        for s in range(numSlices):
            result[s] = prefactor*math.pi*(sum1[s]/numInteractions/(9*cutoff**9)
                                           - sum2[s]/numInteractions/(3*cutoff**3) + sum3[s]/numInteractions)
        return result


class CalcSlicedNonbondedForceKernel:
    """The plugin kernel interface, openmmapi/include/NonbondedSlicingKernels.h:27-85."""
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

    @staticmethod
    def Name():
        return "CalcSlicedNonbondedForce"

    def initialize(self, system, force):
        raise NotImplementedError

    def execute(self, context, includeForces, includeEnergy, includeDirect, includeReciprocal):
        raise NotImplementedError

    def copyParametersToContext(self, context, force):
        raise NotImplementedError

    def getPMEParameters(self):
        raise NotImplementedError

    def getLJPMEParameters(self):
        raise NotImplementedError


def build_desc(system, force, flags=0, device_index=0, legal_grid=False):
    """Everything ReferenceCalcSlicedNonbondedForceKernel::initialize reads from the Force
    (platforms/reference/src/ReferenceNonbondedSlicingKernels.cpp:59-185), flattened into an
    nbs_system_desc.  Returns the abi.DescArrays that own the memory."""
    n = force.getNumParticles()
    method = force.getNonbondedMethod()
    particles = np.array(force._particles, dtype=np.float64).reshape(n, 3)
    subsets = np.array([force.getParticleSubset(i) for i in range(n)], dtype=np.int32)
    exceptions = force._exceptions
    alpha, grid = 0.0, (0, 0, 0)
    if method in (force.PME, force.LJPME):
        alpha, nx, ny, nz = SlicedNonbondedForceImpl.calcPMEParameters(system, force, False)
        grid = (nx, ny, nz)
        if legal_grid:
            grid = tuple(findLegalFFTDimension(g) for g in grid)
    kmax = (0, 0, 0)
    if method == force.Ewald:
        alpha, kx, ky, kz = SlicedNonbondedForceImpl.calcEwaldParameters(system, force)
        kmax = (kx, ky, kz)
    dalpha, dgrid = 0.0, (0, 0, 0)
    if method == force.LJPME:
        dalpha, dnx, dny, dnz = SlicedNonbondedForceImpl.calcPMEParameters(system, force, True)
        dgrid = (dnx, dny, dnz)
        if legal_grid:
            dgrid = tuple(findLegalFFTDimension(g) for g in dgrid)
    dispersion = None
    if force.getUseDispersionCorrection():
        dispersion = SlicedNonbondedForceImpl.calcDispersionCorrections(system, force)
    # NoCutoff ignores the switch (:146-148); LJPME forces it off (:166)
    use_switch = force.getUseSwitchingFunction() and method not in (force.NoCutoff, force.LJPME)
    return abi.DescArrays(
        num_particles=n,
        num_subsets=force.getNumSubsets(),
        method=method,
        subsets=subsets,
        charges=particles[:, 0],
        sigmas=particles[:, 1],
        epsilons=particles[:, 2],
        num_exceptions=len(exceptions),
        num_global_params=force.getNumGlobalParameters(),
        exception_particles=np.array([[e[0], e[1]] for e in exceptions], dtype=np.int32).reshape(-1, 2),
        exception_params=np.array([[e[2], e[3], e[4]] for e in exceptions], dtype=np.float64).reshape(-1, 3),
        num_particle_offsets=len(force._particleOffsets),
        num_exception_offsets=len(force._exceptionOffsets),
        particle_offset_indices=np.array([[o[0], o[1]] for o in force._particleOffsets], dtype=np.int32).reshape(-1, 2),
        particle_offset_scales=np.array([o[2:5] for o in force._particleOffsets], dtype=np.float64).reshape(-1, 3),
        exception_offset_indices=np.array([[o[0], o[1]] for o in force._exceptionOffsets], dtype=np.int32).reshape(-1, 2),
        exception_offset_scales=np.array([o[2:5] for o in force._exceptionOffsets], dtype=np.float64).reshape(-1, 3),
        cutoff=force.getCutoffDistance(),
        switching_distance=force.getSwitchingDistance(),
        rf_dielectric=force.getReactionFieldDielectric(),
        ewald_alpha=alpha,
        pme_grid=grid,
        use_switching_function=int(use_switch),
        exceptions_use_periodic=int(force.getExceptionsUsePeriodicBoundaryConditions()),
        device_index=device_index,
        flags=flags,
        dispersion_coefficients=dispersion,
        ewald_kmax=kmax,
        dispersion_alpha=dalpha,
        dispersion_grid=dgrid,
    )


class SlicedKernelBase(CalcSlicedNonbondedForceKernel):
    """Platform-independent part of a kernel: which scaling parameter drives which (slice, term),
    which derivatives were requested, and the lambda-weighted epilogue
    (ReferenceNonbondedSlicingKernels.cpp:74-86, 252-265, 343-347).  Subclasses provide
    ``_create`` / ``_update`` / ``_evaluate``."""
    Coul, vdW = 0, 1

    def initialize(self, system, force):
        self.numParticles = force.getNumParticles()
        self.numSubsets = force.getNumSubsets()
        self.numSlices = force.getNumSlices()
        self.nonbondedMethod = force.getNonbondedMethod()
        requested = {force.getEnergyParameterDerivativeName(i) for i in range(force.getNumEnergyParameterDerivatives())}
        self.sliceScalingParams = [[("", False), ("", False)] for _ in range(self.numSlices)]
        for index in range(force.getNumScalingParameters()):
            name, i, j, includeCoulomb, includeLJ = force.getScalingParameter(index)
            info = (name, name in requested)
            if includeCoulomb:
                self.sliceScalingParams[sliceIndex(i, j)][self.Coul] = info
            if includeLJ:
                self.sliceScalingParams[sliceIndex(i, j)][self.vdW] = info
        self.globalNames = [force.getGlobalParameterName(i) for i in range(force.getNumGlobalParameters())]
        self.desc = build_desc(system, force, **self._desc_options())
        self.ewaldAlpha = self.desc.desc.ewald_alpha
        self.gridSize = tuple(self.desc.desc.pme_grid)
        self.num14 = self._count14(force)
        self._create()

    @staticmethod
    def _count14(force):
        withOffsets = {o[1] for o in force._exceptionOffsets}
        return sum(1 for k, e in enumerate(force._exceptions) if e[2] != 0.0 or e[4] != 0.0 or k in withOffsets)

    def _desc_options(self):
        return {}

    def execute(self, context, includeForces, includeEnergy, includeDirect, includeReciprocal):
        lambdas = np.ones((self.numSlices, 2))
        for s in range(self.numSlices):
            for t in range(2):
                name = self.sliceScalingParams[s][t][0]
                if name != "":
                    lambdas[s, t] = context.getParameter(name)
        globalValues = np.array([context.getParameter(name) for name in self.globalNames], dtype=np.float64)
        box = context.getPeriodicBoxVectors()
        sliceEnergies = self._evaluate(context.positions, box, lambdas, globalValues, includeDirect,
                                       includeReciprocal, context.forces)
        self.lastSliceEnergies = sliceEnergies
        energy = float((lambdas*sliceEnergies).sum()) if includeEnergy else 0.0
        for s in range(self.numSlices):
            for t in range(2):
                name, hasDerivative = self.sliceScalingParams[s][t]
                if hasDerivative:
                    context.energyParameterDerivatives[name] = context.energyParameterDerivatives.get(name, 0.0) + sliceEnergies[s, t]
        return energy

    def copyParametersToContext(self, context, force):
        """ReferenceNonbondedSlicingKernels.cpp:270-319"""
        if force.getNumParticles() != self.numParticles:
            raise OpenMMException("updateParametersInContext: The number of particles has changed")
        if self._count14(force) != self.num14:
            raise OpenMMException("updateParametersInContext: The number of non-excluded exceptions has changed")
        self.desc = build_desc(context.getSystem(), force, **self._desc_options())
        self._update()

    def getPMEParameters(self):
        if self.nonbondedMethod not in (self.PME, self.LJPME):
            raise OpenMMException("getPMEParametersInContext: This Context is not using PME or LJPME")
        return (self.ewaldAlpha,)+tuple(self.gridSize)

    def getLJPMEParameters(self):
        if self.nonbondedMethod != self.LJPME:
            raise OpenMMException("getPMEParametersInContext: This Context is not using LJPME")
        return (self.desc.desc.dispersion_alpha,)+tuple(self.desc.desc.dispersion_grid)


def findLegalFFTDimension(minimum):
    """Smallest size >= minimum whose prime factors are all <= 13 -- what the plugin's GPU platforms do
    with the PME grid (platforms/common/include/FFT3DFactory.h:31-47, used at
    CommonNonbondedSlicingKernels.cpp:441-443); the Reference platform keeps the size as given."""
    n = max(int(minimum), 1)
    while True:
        m = n
        for f in (2, 3, 5, 7, 11, 13):
            while m % f == 0:
                m //= f
        if m == 1:
            return n
        n += 1


_NO_GLOBALS = np.zeros(0)


class B200CalcSlicedNonbondedForceKernel(SlicedKernelBase):
    """The kernel of the "B200" platform: forwards to the CUDA library through the C ABI."""

    def __init__(self, platform):
        self.platform = platform
        self.handle = None
        self.lib = abi.load_library()

    def _desc_options(self):
        return {"flags": self.platform.flags, "device_index": self.platform.deviceIndex, "legal_grid": True}

    def _create(self):
        handle = C.c_void_p()
        abi.check(self.lib.nbs_create(C.byref(self.desc.desc), C.byref(handle)))
        self.handle = handle
        self._lastLambdas = None
        self._lastGlobals = None

    def _update(self):
        abi.check(self.lib.nbs_update_parameters(self.handle, C.byref(self.desc.desc)))
        self._lastGlobals = None

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.nbs_destroy(self.handle)
            self.handle = None

    def _push_parameters(self, lambdas, globalValues):
        lambdas = np.asarray(lambdas, dtype=np.float64)
        if self._lastLambdas is None or not np.array_equal(lambdas, self._lastLambdas):
            lam = np.ascontiguousarray(lambdas, dtype=np.float64)
            abi.check(self.lib.nbs_set_lambdas(self.handle, lam.ctypes.data_as(C.POINTER(C.c_double))))
            self._lastLambdas = lam.copy()
        if len(globalValues) and (self._lastGlobals is None or not np.array_equal(globalValues, self._lastGlobals)):
            abi.check(self.lib.nbs_set_global_parameters(self.handle, globalValues.ctypes.data_as(C.POINTER(C.c_double))))
            self._lastGlobals = globalValues.copy()

    def _exec_args(self, key):
        """A persistent nbs_exec_args per calling pattern: the per-call work is then a handful of field stores (an
        evaluation takes a quarter of a millisecond; rebuilding the structure in Python was a tenth of that)."""
        cache = self.__dict__.setdefault("_argsCache", {})
        entry = cache.get(key)
        if entry is None:
            args = abi.ExecArgs()
            args.struct_size = C.sizeof(abi.ExecArgs)
            args.include_forces = 1
            args.include_energy = 1
            energies = np.zeros((self.numSlices, 2))
            entry = cache[key] = {"args": args, "energies": energies, "energies_ptr": energies.ctypes.data_as(C.POINTER(C.c_double)),
                                  "box": None}
        return entry

    @staticmethod
    def _set_box(entry, box):
        b = np.asarray(box, dtype=np.float64).reshape(9)
        if entry["box"] is None or not np.array_equal(b, entry["box"]):
            entry["args"].box[:] = b.tolist()
            entry["box"] = b.copy()

    def _evaluate(self, positions, box, lambdas, globalValues, includeDirect, includeReciprocal, forces, accumulate=True):
        """Host buffers in, host buffers out (the Reference platform's calling pattern).  `accumulate`: forces are ADDED to
        `forces` (what execute() of the plugin's kernel interface does); False overwrites them, which lets the library copy
        straight into the caller's buffer."""
        self._push_parameters(lambdas, globalValues)
        entry = self._exec_args(("host", bool(accumulate)))
        args = entry["args"]
        args.positions_format = abi.NBS_POS_F64_XYZ
        args.positions_space = abi.NBS_MEM_HOST
        args.forces_format = abi.NBS_FORCE_F64_XYZ
        args.forces_space = abi.NBS_MEM_HOST
        args.forces_accumulate = 1 if accumulate else 0
        pos = positions if (isinstance(positions, np.ndarray) and positions.dtype == np.float64 and positions.flags.c_contiguous) \
            else np.ascontiguousarray(positions, dtype=np.float64)
        args.positions = pos.ctypes.data
        assert forces.dtype == np.float64 and forces.flags.c_contiguous
        args.forces = forces.ctypes.data
        self._set_box(entry, box)
        args.include_direct = int(includeDirect)
        args.include_reciprocal = int(includeReciprocal)
        args.slice_energies = entry["energies_ptr"]
        args.stream = None
        abi.check(self.lib.nbs_execute(self.handle, C.byref(args)))
        return entry["energies"].copy()

    # ---- device-resident evaluation used by bench.py and multi-GPU drivers ---------------------
    def execute_device(self, positions_ptr, box, forces_ptr, lambdas, includeDirect=True, includeReciprocal=True,
                       stream=0, forces_format=abi.NBS_FORCE_F64_XYZ, padded_num_atoms=0, accumulate=0,
                       want_energies=True):
        """One evaluation with positions (double[N][3]) and forces already resident in HBM."""
        self._push_parameters(lambdas, _NO_GLOBALS)
        entry = self._exec_args(("device", forces_format))
        args = entry["args"]
        args.positions_format = abi.NBS_POS_F64_XYZ
        args.positions_space = abi.NBS_MEM_DEVICE
        args.forces_format = forces_format
        args.forces_space = abi.NBS_MEM_DEVICE
        args.forces_accumulate = accumulate
        args.positions = positions_ptr
        args.forces = forces_ptr
        args.padded_num_atoms = padded_num_atoms
        self._set_box(entry, box)
        args.include_direct = int(includeDirect)
        args.include_reciprocal = int(includeReciprocal)
        args.slice_energies = entry["energies_ptr"] if want_energies else None
        args.stream = stream
        abi.check(self.lib.nbs_execute(self.handle, C.byref(args)))
        return entry["energies"].copy() if want_energies else None

    # ---- parity diagnostics ----------------------------------------------------------------------
    def getPairSet(self, with_pairs=True):
        count = C.c_int64()
        h = C.c_uint64()
        abi.check(self.lib.nbs_get_pair_set(self.handle, 0, None, C.byref(count), C.byref(h)))
        pairs = None
        if with_pairs:
            pairs = np.zeros((count.value, 2), dtype=np.int32)
            abi.check(self.lib.nbs_get_pair_set(self.handle, count.value, pairs.ctypes.data_as(C.POINTER(C.c_int32)),
                                                C.byref(count), C.byref(h)))
        return count.value, h.value, pairs

    def getExclusionSet(self):
        count = C.c_int64()
        abi.check(self.lib.nbs_get_exclusion_set(self.handle, 0, None, C.byref(count)))
        pairs = np.zeros((count.value, 2), dtype=np.int32)
        abi.check(self.lib.nbs_get_exclusion_set(self.handle, count.value, pairs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(count)))
        return pairs

    def getKernelTimes(self):
        cap = 64
        names = (C.c_char_p*cap)()
        ms = (C.c_float*cap)()
        count = C.c_int32()
        abi.check(self.lib.nbs_get_kernel_times(self.handle, cap, names, ms, C.byref(count)))
        return [(names[i].decode(), ms[i]) for i in range(count.value)]

    def getLaunchCount(self):
        v = C.c_int64()
        abi.check(self.lib.nbs_get_launch_count(self.handle, C.byref(v)))
        return v.value

    def getNlistStats(self):
        v = (C.c_int64*8)()
        abi.check(self.lib.nbs_get_nlist_stats(self.handle, v))
        return list(v)

    def setListSkin(self, skin):
        """Neighbour-list padding in nm (0 = rebuild on every evaluation, like the Reference platform)."""
        abi.check(self.lib.nbs_set_list_skin(self.handle, float(skin)))

    def getListStats(self):
        v = (C.c_double*8)()
        abi.check(self.lib.nbs_get_list_stats(self.handle, v))
        return {"evaluations": int(v[0]), "builds": int(v[1]), "redone": int(v[2]), "max_displacement": v[3],
                "skin": v[4], "list_valid": bool(v[5]), "reused_last": bool(v[6]), "max_step_growth": v[7]}


class Platform:
    """The "B200" platform: creates kernels backed by the CUDA library."""

    def __init__(self, deviceIndex=0, flags=0, properties=None):
        self.deviceIndex = deviceIndex
        self.flags = flags
        self.properties = {"Precision": "mixed", "DeterministicForces": "false"}
        for name, value in (properties or {}).items():
            self.setPropertyDefaultValue(name, value)

    def getName(self):
        return "B200"

    # the two properties of OpenMM's CUDA platform that reach this force (CudaNonbondedSlicingKernels.cpp:22-33,
    # CommonNonbondedSlicingKernels.cpp:297-299): Precision = single | mixed | double, DeterministicForces = true | false
    def getPropertyDefaultValue(self, name):
        return self.properties[name]

    def setPropertyDefaultValue(self, name, value):
        if name == "Precision":
            if value not in ("single", "mixed", "double"):
                raise OpenMMException("Illegal value for Precision: " + str(value))
            self.flags &= ~(abi.NBS_FLAG_DOUBLE | abi.NBS_FLAG_FP32_ENERGY)
            self.flags |= {"single": abi.NBS_FLAG_FP32_ENERGY, "mixed": 0, "double": abi.NBS_FLAG_DOUBLE}[value]
        elif name == "DeterministicForces":
            self.flags &= ~abi.NBS_FLAG_DETERMINISTIC
            if str(value).lower() == "true":
                self.flags |= abi.NBS_FLAG_DETERMINISTIC
        else:
            raise OpenMMException("Illegal property name: " + str(name))
        self.properties[name] = value

    def createKernel(self, name, context):
        if name != CalcSlicedNonbondedForceKernel.Name():
            raise OpenMMException(f"Called createKernel() on a Platform which does not support the requested kernel: {name}")
        return B200CalcSlicedNonbondedForceKernel(self)


class State:
    def __init__(self, energy, forces, derivatives, positions):
        self._energy, self._forces, self._derivatives, self._positions = energy, forces, derivatives, positions

    def getPotentialEnergy(self):
        return self._energy

    def getForces(self):
        return self._forces

    def getPositions(self):
        return self._positions

    def getEnergyParameterDerivatives(self):
        return self._derivatives


class Context:
    """Just enough of OpenMM's Context/ContextImpl [external] to drive SlicedNonbondedForce objects."""

    def __init__(self, system, platform):
        self.system = system
        self.platform = platform
        self.positions = np.zeros((system.getNumParticles(), 3))
        self.forces = np.zeros((system.getNumParticles(), 3))
        self.box = system.getDefaultPeriodicBoxVectors()
        self.energyParameterDerivatives = {}
        self._initialize()

    def _initialize(self):
        self.parameters = {}
        self.impls = []
        for force in self.system._forces:
            impl = SlicedNonbondedForceImpl(force)
            self.impls.append(impl)
            for name, value in impl.getDefaultParameters().items():
                self.parameters.setdefault(name, value)
        for impl in self.impls:
            impl.initialize(self)

    def _impl(self, force):
        for impl in self.impls:
            if impl.owner is force:
                return impl
        raise OpenMMException("This Force is not part of this Context's System")

    def getSystem(self):
        return self.system

    def getPlatform(self):
        return self.platform

    def reinitialize(self, preserveState=False):
        positions, box, params = self.positions.copy(), self.box.copy(), dict(self.parameters)
        self._initialize()
        if preserveState:
            self.positions, self.box = positions, box
            for name, value in params.items():
                if name in self.parameters:
                    self.parameters[name] = value
        else:
            self.positions = np.zeros_like(positions)
            self.box = self.system.getDefaultPeriodicBoxVectors()

    def setPositions(self, positions):
        positions = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
        if positions.shape[0] != self.system.getNumParticles():
            raise OpenMMException("Called setPositions() on a Context with the wrong number of positions")
        self.positions = positions.copy()

    def setPeriodicBoxVectors(self, a, b, c):
        self.box = np.array([list(a), list(b), list(c)], dtype=np.float64)

    def getPeriodicBoxVectors(self):
        return self.box

    def getParameter(self, name):
        if name not in self.parameters:
            raise OpenMMException(f"Called getParameter() with invalid parameter name: {name}")
        return self.parameters[name]

    def setParameter(self, name, value):
        if name not in self.parameters:
            raise OpenMMException(f"Called setParameter() with invalid parameter name: {name}")
        self.parameters[name] = float(value)

    def getParameters(self):
        return dict(self.parameters)

    def getState(self, getEnergy=False, getForces=False, getParameterDerivatives=False, getPositions=False, groups=0xFFFFFFFF):
        if isinstance(groups, (set, frozenset, list, tuple)):
            groups = sum(1 << g for g in groups)
        self.forces = np.zeros((self.system.getNumParticles(), 3))
        self.energyParameterDerivatives = {}
        for impl in self.impls:
            for i in range(impl.owner.getNumEnergyParameterDerivatives()):
                self.energyParameterDerivatives.setdefault(impl.owner.getEnergyParameterDerivativeName(i), 0.0)
        energy = 0.0
        for impl in self.impls:
            energy += impl.calcForcesAndEnergy(self, getForces, getEnergy, groups)
        return State(energy if getEnergy else None, self.forces.copy() if getForces else None,
                     dict(self.energyParameterDerivatives) if getParameterDerivatives else None,
                     self.positions.copy() if getPositions else None)

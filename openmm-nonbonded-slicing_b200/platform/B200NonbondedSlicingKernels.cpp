// Host side of the drop-in: reads the Force exactly like ReferenceCalcSlicedNonbondedForceKernel::initialize
// (platforms/reference/src/ReferenceNonbondedSlicingKernels.cpp:59-185), hands the flattened description to
// the CUDA library, and per evaluation does what ReferenceCalcSlicedNonbondedForceKernel::execute does
// around the arithmetic (:187-268): read the lambdas and offset parameters from the Context (:339-347),
// run the device path, return sum(lambda * E_slice) and add E_slice to the requested derivatives (:252-265).
// Device buffers are OpenMM CUDA's own: posq (float4/double4, OpenMM's atom order + atomIndex), the
// long-long fixed-point force buffer, the current stream -- nothing is copied through the host.
#include "B200NonbondedSlicingKernels.h"
#include "SlicedNonbondedForce.h"
#include "internal/SlicedNonbondedForceImpl.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include <map>
#include <set>

using namespace NonbondedSlicing;
using namespace OpenMM;
using namespace std;

struct B200CalcSlicedNonbondedForceKernel::Description {
    nbs_system_desc desc;
    vector<int32_t> subsets, exceptionParticles, particleOffsetIndices, exceptionOffsetIndices;
    vector<double> charges, sigmas, epsilons, exceptionParams, particleOffsetScales, exceptionOffsetScales, dispersion;
};

B200CalcSlicedNonbondedForceKernel::~B200CalcSlicedNonbondedForceKernel() {
    if (handle != nullptr) {
        cu.setAsCurrent();
        nbs_destroy(handle);
    }
}

void B200CalcSlicedNonbondedForceKernel::check(int status) const {
    if (status != NBS_OK)
        throw OpenMMException(nbs_last_error());
}

int B200CalcSlicedNonbondedForceKernel::findLegalFFTDimension(int minimum) {
    // platforms/common/include/FFT3DFactory.h:31-47 (factors up to 13, the VkFFT set)
    for (int n = max(minimum, 1); ; n++) {
        int m = n;
        for (int f : {2, 3, 5, 7, 11, 13})
            while (m % f == 0) m /= f;
        if (m == 1) return n;
    }
}

void B200CalcSlicedNonbondedForceKernel::describe(const System& system, const SlicedNonbondedForce& force, Description& d) const {
    const int n = force.getNumParticles();
    map<string, int> globalIndex;
    for (int i = 0; i < force.getNumGlobalParameters(); i++)
        globalIndex[force.getGlobalParameterName(i)] = i;
    d.subsets.resize(n); d.charges.resize(n); d.sigmas.resize(n); d.epsilons.resize(n);
    for (int i = 0; i < n; i++) {
        d.subsets[i] = force.getParticleSubset(i);
        force.getParticleParameters(i, d.charges[i], d.sigmas[i], d.epsilons[i]);
    }
    const int numExceptions = force.getNumExceptions();
    d.exceptionParticles.resize(2*numExceptions); d.exceptionParams.resize(3*numExceptions);
    for (int i = 0; i < numExceptions; i++) {
        int p1, p2;
        force.getExceptionParameters(i, p1, p2, d.exceptionParams[3*i], d.exceptionParams[3*i+1], d.exceptionParams[3*i+2]);
        d.exceptionParticles[2*i] = p1;
        d.exceptionParticles[2*i+1] = p2;
    }
    for (int i = 0; i < force.getNumParticleParameterOffsets(); i++) {
        string param; int particle; double charge, sigma, epsilon;
        force.getParticleParameterOffset(i, param, particle, charge, sigma, epsilon);
        d.particleOffsetIndices.insert(d.particleOffsetIndices.end(), {globalIndex[param], particle});
        d.particleOffsetScales.insert(d.particleOffsetScales.end(), {charge, sigma, epsilon});
    }
    for (int i = 0; i < force.getNumExceptionParameterOffsets(); i++) {
        string param; int exception; double charge, sigma, epsilon;
        force.getExceptionParameterOffset(i, param, exception, charge, sigma, epsilon);
        d.exceptionOffsetIndices.insert(d.exceptionOffsetIndices.end(), {globalIndex[param], exception});
        d.exceptionOffsetScales.insert(d.exceptionOffsetScales.end(), {charge, sigma, epsilon});
    }
    nbs_system_desc& s = d.desc;
    s = nbs_system_desc();
    s.struct_size = sizeof(nbs_system_desc);
    s.num_particles = n;
    s.num_subsets = force.getNumSubsets();
    s.method = (int) force.getNonbondedMethod();
    s.subsets = d.subsets.data(); s.charges = d.charges.data(); s.sigmas = d.sigmas.data(); s.epsilons = d.epsilons.data();
    s.num_exceptions = numExceptions;
    s.num_global_params = force.getNumGlobalParameters();
    s.exception_particles = d.exceptionParticles.data();
    s.exception_params = d.exceptionParams.data();
    s.num_particle_offsets = force.getNumParticleParameterOffsets();
    s.num_exception_offsets = force.getNumExceptionParameterOffsets();
    s.particle_offset_indices = d.particleOffsetIndices.data();
    s.particle_offset_scales = d.particleOffsetScales.data();
    s.exception_offset_indices = d.exceptionOffsetIndices.data();
    s.exception_offset_scales = d.exceptionOffsetScales.data();
    s.cutoff = force.getCutoffDistance();
    s.switching_distance = force.getSwitchingDistance();
    s.rf_dielectric = force.getReactionFieldDielectric();
    // NoCutoff ignores the switch (:146-148); LJPME forces it off (:166)
    s.use_switching_function = force.getUseSwitchingFunction() && force.getNonbondedMethod() != SlicedNonbondedForce::NoCutoff
                               && force.getNonbondedMethod() != SlicedNonbondedForce::LJPME;
    s.exceptions_use_periodic = force.getExceptionsUsePeriodicBoundaryConditions();
    if (force.getNonbondedMethod() == SlicedNonbondedForce::PME || force.getNonbondedMethod() == SlicedNonbondedForce::LJPME) {
        int nx, ny, nz;
        SlicedNonbondedForceImpl::calcPMEParameters(system, force, s.ewald_alpha, nx, ny, nz, false);   // :163-167
        s.pme_grid[0] = findLegalFFTDimension(nx);      // like CommonNonbondedSlicingKernels.cpp:441-443
        s.pme_grid[1] = findLegalFFTDimension(ny);
        s.pme_grid[2] = findLegalFFTDimension(nz);
    }
    if (force.getNonbondedMethod() == SlicedNonbondedForce::Ewald) {
        int kx, ky, kz;
        SlicedNonbondedForceImpl::calcEwaldParameters(system, force, s.ewald_alpha, kx, ky, kz);         // :158-162
        s.ewald_kmax[0] = kx; s.ewald_kmax[1] = ky; s.ewald_kmax[2] = kz;
    }
    if (force.getNonbondedMethod() == SlicedNonbondedForce::LJPME) {
        int nx, ny, nz;
        SlicedNonbondedForceImpl::calcPMEParameters(system, force, s.dispersion_alpha, nx, ny, nz, true);  // :172-173
        s.dispersion_grid[0] = findLegalFFTDimension(nx);
        s.dispersion_grid[1] = findLegalFFTDimension(ny);
        s.dispersion_grid[2] = findLegalFFTDimension(nz);
    }
    if (force.getUseDispersionCorrection()) {           // :181-184, the unchanged API library does the maths
        d.dispersion = SlicedNonbondedForceImpl::calcDispersionCorrections(system, force);
        s.dispersion_coefficients = d.dispersion.data();
    }
    s.device_index = cu.getDeviceIndex();
    s.flags = cu.getPlatformData().deterministicForces ? NBS_FLAG_DETERMINISTIC : 0;
    // the platform's Precision property (CommonNonbondedSlicingKernels.cpp:297-299): "double" selects double-precision
    // forces, "single" single-precision energies; "mixed" is the library's default (fp32 forces, fp64 energies)
    if (cu.getUseDoublePrecision()) s.flags |= NBS_FLAG_DOUBLE;
    else if (!cu.getUseMixedPrecision()) s.flags |= NBS_FLAG_FP32_ENERGY;
}

void B200CalcSlicedNonbondedForceKernel::initialize(const System& system, const SlicedNonbondedForce& force) {
    cu.setAsCurrent();
    if (cu.getPlatformData().contexts.size() > 1)
        throw OpenMMException("SlicedNonbondedForce (B200): one CUDA context per Context; use one process per GPU "
                              "(the multi-GPU driver shards i-blocks and PME grids across ranks with NCCL)");
    numParticles = force.getNumParticles();
    numSlices = force.getNumSlices();
    nonbondedMethod = CalcSlicedNonbondedForceKernel::NonbondedMethod(force.getNonbondedMethod());
    // which scaling parameter drives which (slice, term), and which derivatives were requested (:70-86)
    set<string> requestedDerivatives;
    for (int i = 0; i < force.getNumEnergyParameterDerivatives(); i++)
        requestedDerivatives.insert(force.getEnergyParameterDerivativeName(i));
    sliceScalingParams.assign(numSlices, {});
    for (int index = 0; index < force.getNumScalingParameters(); index++) {
        string name; int i, j; bool includeCoulomb, includeLJ;
        force.getScalingParameter(index, name, i, j, includeCoulomb, includeLJ);
        ScalingParameterInfo info;
        info.name = name;
        info.hasDerivative = requestedDerivatives.count(name) > 0;
        int slice = sliceIndex(i, j);
        if (includeCoulomb) sliceScalingParams[slice][0] = info;
        if (includeLJ) sliceScalingParams[slice][1] = info;
    }
    globalNames.clear();
    for (int i = 0; i < force.getNumGlobalParameters(); i++)
        globalNames.push_back(force.getGlobalParameterName(i));
    Description d;
    describe(system, force, d);
    ewaldAlpha = d.desc.ewald_alpha;
    for (int k = 0; k < 3; k++) gridSize[k] = d.desc.pme_grid[k];
    check(nbs_create(&d.desc, &handle));
    lastLambdas.clear();
    lastGlobals.clear();
}

double B200CalcSlicedNonbondedForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy,
                                                   bool includeDirect, bool includeReciprocal) {
    cu.setAsCurrent();
    // computeParameters (:339-392): lambdas default to 1, offsets read the Context's global parameters
    vector<double> lambdas(2*numSlices, 1.0), globals(globalNames.size());
    for (int slice = 0; slice < numSlices; slice++)
        for (int term = 0; term < 2; term++)
            if (!sliceScalingParams[slice][term].name.empty())
                lambdas[2*slice+term] = context.getParameter(sliceScalingParams[slice][term].name);
    for (size_t k = 0; k < globalNames.size(); k++)
        globals[k] = context.getParameter(globalNames[k]);
    if (lambdas != lastLambdas) { check(nbs_set_lambdas(handle, lambdas.data())); lastLambdas = lambdas; }
    if (globals != lastGlobals) { check(nbs_set_global_parameters(handle, globals.data())); lastGlobals = globals; }

    nbs_exec_args args = nbs_exec_args();
    args.struct_size = sizeof(nbs_exec_args);
    args.positions_format = cu.getUseDoublePrecision() ? NBS_POS_F64_XYZW : NBS_POS_F32_XYZW;
    args.positions_space = NBS_MEM_DEVICE;
    args.positions = (const void*) cu.getPosq().getDevicePointer();
    args.atom_index = (const int32_t*) cu.getAtomIndexArray().getDevicePointer();
    args.forces_format = NBS_FORCE_I64_FIXED;       // OpenMM's own accumulator: no conversion, no copy
    args.forces_space = NBS_MEM_DEVICE;
    args.forces = (void*) cu.getLongForceBuffer().getDevicePointer();
    args.padded_num_atoms = cu.getPaddedNumAtoms();
    Vec3 a, b, c;
    cu.getPeriodicBoxVectors(a, b, c);
    for (int k = 0; k < 3; k++) { args.box[k] = a[k]; args.box[3+k] = b[k]; args.box[6+k] = c[k]; }
    args.include_forces = includeForces;
    args.include_energy = includeEnergy;
    args.include_direct = includeDirect;
    args.include_reciprocal = includeReciprocal;
    vector<double> sliceEnergies(2*numSlices, 0.0);
    bool anyDerivative = false;
    for (auto& pair : sliceScalingParams) anyDerivative = anyDerivative || pair[0].hasDerivative || pair[1].hasDerivative;
    args.slice_energies = (includeEnergy || anyDerivative) ? sliceEnergies.data() : nullptr;   // NULL => force-only kernels
    args.stream = (void*) cu.getCurrentStream();
    check(nbs_execute(handle, &args));

    double energy = 0;                                // :252-257
    if (includeEnergy)
        for (int k = 0; k < 2*numSlices; k++) energy += lambdas[k]*sliceEnergies[k];
    map<string, double>& energyParamDerivs = cu.getEnergyParamDerivWorkspace();     // :259-265
    for (int slice = 0; slice < numSlices; slice++)
        for (int term = 0; term < 2; term++) {
            const ScalingParameterInfo& info = sliceScalingParams[slice][term];
            if (info.hasDerivative) energyParamDerivs[info.name] += sliceEnergies[2*slice+term];
        }
    return energy;
}

void B200CalcSlicedNonbondedForceKernel::copyParametersToContext(ContextImpl& context, const SlicedNonbondedForce& force) {
    cu.setAsCurrent();
    if (force.getNumParticles() != numParticles)      // :271-272 (the library re-checks, incl. the 1-4 count :297-298)
        throw OpenMMException("updateParametersInContext: The number of particles has changed");
    Description d;
    describe(context.getSystem(), force, d);
    check(nbs_update_parameters(handle, &d.desc));
    lastGlobals.clear();
}

void B200CalcSlicedNonbondedForceKernel::getPMEParameters(double& alpha, int& nx, int& ny, int& nz) const {
    if (nonbondedMethod != PME && nonbondedMethod != LJPME)     // :321-328
        throw OpenMMException("getPMEParametersInContext: This Context is not using PME or LJPME");
    alpha = ewaldAlpha; nx = gridSize[0]; ny = gridSize[1]; nz = gridSize[2];
}

void B200CalcSlicedNonbondedForceKernel::getLJPMEParameters(double& alpha, int& nx, int& ny, int& nz) const {
    if (nonbondedMethod != LJPME)                               // :330-337
        throw OpenMMException("getPMEParametersInContext: This Context is not using LJPME");
    int32_t gx, gy, gz;
    check(nbs_get_ljpme_parameters(handle, &alpha, &gx, &gy, &gz));
    nx = gx; ny = gy; nz = gz;
}

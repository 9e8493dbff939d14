#ifndef B200_NONBONDED_SLICING_KERNELS_H_
#define B200_NONBONDED_SLICING_KERNELS_H_
// The B200 platform kernel: a drop-in implementation of the plugin's own kernel interface
// CalcSlicedNonbondedForceKernel (openmmapi/include/NonbondedSlicingKernels.h:27-85) for OpenMM's "CUDA"
// platform.  It stands exactly where CudaCalcSlicedNonbondedForceKernel stands
// (platforms/cuda/include/CudaNonbondedSlicingKernels.h): same constructor shape, same five virtuals,
// same registration symbols -- but all device work goes through the C ABI in include/nbslice_b200.h.
// SlicedNonbondedForce, SlicedNonbondedForceImpl, serialization and the SWIG wrapper are untouched.
#include "NonbondedSlicingKernels.h"
#include "nbslice_b200.h"
#include "openmm/cuda/CudaContext.h"
#include <array>
#include <string>
#include <vector>

namespace NonbondedSlicing {

class B200CalcSlicedNonbondedForceKernel : public CalcSlicedNonbondedForceKernel {
public:
    B200CalcSlicedNonbondedForceKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu,
                                       const OpenMM::System& system)
        : CalcSlicedNonbondedForceKernel(name, platform), cu(cu), handle(nullptr) {}
    ~B200CalcSlicedNonbondedForceKernel();
    void initialize(const OpenMM::System& system, const SlicedNonbondedForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal);
    void copyParametersToContext(OpenMM::ContextImpl& context, const SlicedNonbondedForce& force);
    void getPMEParameters(double& alpha, int& nx, int& ny, int& nz) const;
    void getLJPMEParameters(double& alpha, int& nx, int& ny, int& nz) const;
private:
    struct ScalingParameterInfo { std::string name; bool hasDerivative = false; };
    struct Description;                         // owns the arrays an nbs_system_desc points into
    void describe(const OpenMM::System& system, const SlicedNonbondedForce& force, Description& out) const;
    static int findLegalFFTDimension(int minimum);
    void check(int status) const;
    OpenMM::CudaContext& cu;
    nbs_context* handle;
    int numParticles, numSlices;
    NonbondedMethod nonbondedMethod;
    double ewaldAlpha;
    int gridSize[3];
    std::vector<std::array<ScalingParameterInfo, 2>> sliceScalingParams;   // [slice][Coul=0, vdW=1]
    std::vector<std::string> globalNames;
    std::vector<double> lastLambdas, lastGlobals;
};

} // namespace NonbondedSlicing
#endif

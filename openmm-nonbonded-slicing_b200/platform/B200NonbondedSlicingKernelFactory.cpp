// Registration: the symbols OpenMM's Platform::loadPluginsFromDirectory looks up, exactly as
// platforms/cuda/src/CudaNonbondedSlicingKernelFactory.cpp:19-54 defines them.  Installing this
// library INSTEAD OF libNonbondedSlicingCUDA.so in $OPENMM_DIR/lib/plugins makes the B200 kernel the
// implementation of "CalcSlicedNonbondedForce" on the CUDA platform.
#include "B200NonbondedSlicingKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include <exception>

using namespace NonbondedSlicing;
using namespace OpenMM;

namespace NonbondedSlicing {
class B200NonbondedSlicingKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
        CudaPlatform::PlatformData& data = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData());
        CudaContext& cu = *data.contexts[0];
        if (name == CalcSlicedNonbondedForceKernel::Name())
            return new B200CalcSlicedNonbondedForceKernel(name, platform, cu, context.getSystem());
        throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '")+name+"'").c_str());
    }
};
}

extern "C" void registerPlatforms() {
}

extern "C" void registerKernelFactories() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        platform.registerKernelFactory(CalcSlicedNonbondedForceKernel::Name(), new B200NonbondedSlicingKernelFactory());
    }
    catch (std::exception& ex) {
        // no CUDA platform in this OpenMM: nothing to register
    }
}

extern "C" void registerNonbondedSlicingB200KernelFactories() {
    registerKernelFactories();
}

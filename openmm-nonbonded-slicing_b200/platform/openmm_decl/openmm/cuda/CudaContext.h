#ifndef NBS_DECL_CUDACONTEXT_H_
#define NBS_DECL_CUDACONTEXT_H_
#include "openmm/Platform.h"
#include "openmm/Vec3.h"
#include <cuda.h>
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class CudaArray {
public:
    CUdeviceptr getDevicePointer();
};
class CudaContext;
class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        std::vector<CudaContext*> contexts;
        bool deterministicForces;
    };
};
class CudaContext {
public:
    void setAsCurrent();
    int getDeviceIndex() const;
    int getContextIndex() const;
    int getNumAtoms() const;
    int getPaddedNumAtoms() const;
    bool getUseDoublePrecision() const;
    bool getUseMixedPrecision() const;
    CUstream getCurrentStream();
    CudaArray& getPosq();
    CudaArray& getAtomIndexArray();
    CudaArray& getLongForceBuffer();     // OpenMM 8.x: getForce() on older releases
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const;
    std::map<std::string, double>& getEnergyParamDerivWorkspace();
    CudaPlatform::PlatformData& getPlatformData();
};
}
#endif

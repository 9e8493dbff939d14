#ifndef NBS_DECL_NONBONDEDFORCEIMPL_H_
#define NBS_DECL_NONBONDEDFORCEIMPL_H_
#include "openmm/NonbondedForce.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"
#include <string>
#include <vector>
namespace OpenMM {
class NonbondedForceImpl {
public:
    NonbondedForceImpl(const NonbondedForce& owner);
    virtual ~NonbondedForceImpl();
    static void calcPMEParameters(const System& system, const NonbondedForce& force, double& alpha, int& xsize, int& ysize, int& zsize, bool lj);
    static void calcEwaldParameters(const System& system, const NonbondedForce& force, double& alpha, int& kmaxx, int& kmaxy, int& kmaxz);
};
}
#endif

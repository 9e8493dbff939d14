#ifndef NBS_DECL_SYSTEM_H_
#define NBS_DECL_SYSTEM_H_
#include "openmm/Vec3.h"
namespace OpenMM {
class System {
public:
    int getNumParticles() const;
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const;
};
}
#endif

// Stand-in for openmm/reference/ReferenceForce.h -- TEST INFRASTRUCTURE ONLY.
// getDeltaR / getDeltaRPeriodic restate OpenMM's published semantics (delta = J - I;
// the periodic version removes whole box vectors c, then b, then a, which is exact for
// rectangular boxes and is OpenMM's convention for reduced triclinic ones) -- SURVEY 8c.
#ifndef NBS_STUB_REFERENCE_FORCE_H_
#define NBS_STUB_REFERENCE_FORCE_H_
#include "openmm/Vec3.h"
#include <cmath>
namespace OpenMM {
class ReferenceForce {
public:
    static const int XIndex = 0;
    static const int YIndex = 1;
    static const int ZIndex = 2;
    static const int R2Index = 3;
    static const int RIndex = 4;
    static const int LastDeltaRIndex = 5;
    static void getDeltaR(const Vec3& atomCoordinatesI, const Vec3& atomCoordinatesJ, double* deltaR) {
        store(atomCoordinatesJ - atomCoordinatesI, deltaR);
    }
    static void getDeltaRPeriodic(const Vec3& atomCoordinatesI, const Vec3& atomCoordinatesJ,
                                  const Vec3* boxVectors, double* deltaR) {
        Vec3 diff = atomCoordinatesJ - atomCoordinatesI;
        diff -= boxVectors[2]*std::floor(diff[2]/boxVectors[2][2] + 0.5);
        diff -= boxVectors[1]*std::floor(diff[1]/boxVectors[1][1] + 0.5);
        diff -= boxVectors[0]*std::floor(diff[0]/boxVectors[0][0] + 0.5);
        store(diff, deltaR);
    }
private:
    static void store(const Vec3& diff, double* deltaR) {
        deltaR[XIndex] = diff[0];
        deltaR[YIndex] = diff[1];
        deltaR[ZIndex] = diff[2];
        deltaR[R2Index] = diff.dot(diff);
        deltaR[RIndex] = std::sqrt(deltaR[R2Index]);
    }
};
}
#endif

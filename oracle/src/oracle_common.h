// oracle_common.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Host-side prelude shared by the two CPU oracles:
//   * the restatement (oracle_core.cpp, "port"), and
//   * the bridge to the reference's own unmodified arithmetic TUs (ref_bridge.cpp, "reference").
// It restates what ReferenceCalcSlicedNonbondedForceKernel::initialize / computeParameters do
// with the Force description (platforms/reference/src/ReferenceNonbondedSlicingKernels.cpp:59-185,
// 339-392) and the neighbour-list contract of the [external] OpenMM call at :197.
#ifndef NBS_ORACLE_COMMON_H_
#define NBS_ORACLE_COMMON_H_

#include "nbslice_b200.h"
#include <array>
#include <cmath>
#include <cstdint>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace nbs_oracle {

// CODATA-2018 values used by OpenMM 8.x (openmm/reference/SimTKOpenMMRealType.h [external]).
constexpr double kPi = 3.14159265358979323846;
constexpr double kECharge = 1.602176634e-19;
constexpr double kAvogadro = 6.02214076e23;
constexpr double kEpsilon0 = 1e-6*8.8541878128e-12/(kECharge*kECharge*kAvogadro);
constexpr double kOne4PiEps0 = 1/(4*kPi*kEpsilon0);

inline int sliceIndex(int i, int j) {                  // openmmapi/include/SlicedNonbondedForce.h:22
    return i > j ? i*(i+1)/2 + j : j*(j+1)/2 + i;
}

struct System {
    int n = 0, numSubsets = 0, numSlices = 0, method = 0;
    std::vector<int> subsets;
    std::vector<std::array<double, 3>> particleParams;      // (sigma/2, 2*sqrt(eps), q)   :364-368
    std::vector<std::set<int>> exclusions;                  // from ALL exceptions          :101-106
    int num14 = 0;
    std::vector<std::array<int, 2>> index14;                // :127-128
    std::vector<std::array<double, 3>> params14;            // (sigma, 4*eps, qq)           :387-391
    std::vector<int> slice14;                               // :129-131
    std::vector<double> dispersionCoefficients;             // [numSlices]                  :181-184
    double cutoff = 0, switchingDistance = 0, rfDielectric = 78.3, alpha = 0;
    int grid[3] = {0, 0, 0};
    int kmax[3] = {0, 0, 0};                                // Ewald: numRx, numRy, numRz   :158-162
    double dispersionAlpha = 0;                             // LJPME                        :168-175
    int dispersionGrid[3] = {0, 0, 0};
    bool useSwitch = false, exceptionsPeriodic = false;
};

// initialize() + computeParameters() for the current global parameter values.
// Returns an empty string on success, else the error text.
std::string buildSystem(const nbs_system_desc& d, const double* globalValues, System& out);

struct Box {
    double v[3][3];
    bool rectangular() const { return v[1][0] == 0 && v[2][0] == 0 && v[2][1] == 0; }
};

// delta = xI - xJ with OpenMM's periodic convention (ReferenceForce::getDeltaRPeriodic [external]).
inline void deltaPeriodic(const double* xI, const double* xJ, const Box& b, double d[3]) {
    for (int k = 0; k < 3; k++) d[k] = xI[k] - xJ[k];
    for (int a = 2; a >= 0; a--) {
        double s = std::floor(d[a]/b.v[a][a] + 0.5);
        for (int k = 0; k < 3; k++) d[k] -= b.v[a][k]*s;
    }
}

typedef std::vector<std::pair<unsigned int, unsigned int>> PairList;   // == OpenMM::NeighborList

// All unordered pairs within `cutoff` (r^2 <= cutoff^2) that are not excluded.
void buildNeighborList(const System& s, const double* pos, const Box& box, bool periodic, PairList& out);

// SlicedNonbondedForceImpl::calcDispersionCorrections (openmmapi/src/SlicedNonbondedForceImpl.cpp:263-354),
// including its int arithmetic.
std::vector<double> dispersionCoefficients(const nbs_system_desc& d, const double* globalDefaults);

} // namespace nbs_oracle
#endif

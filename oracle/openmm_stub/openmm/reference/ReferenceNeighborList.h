// Stand-in for openmm/reference/ReferenceNeighborList.h -- TEST INFRASTRUCTURE ONLY.
// Only the container type is needed by the reference TUs; the list itself is built by
// oracle/src/oracle_common.cpp (restated from the call contract at
// platforms/reference/src/ReferenceNonbondedSlicingKernels.cpp:197).
#ifndef NBS_STUB_REFERENCE_NEIGHBOR_LIST_H_
#define NBS_STUB_REFERENCE_NEIGHBOR_LIST_H_
#include "openmm/Vec3.h"
#include <set>
#include <utility>
#include <vector>
namespace OpenMM {
typedef unsigned int AtomIndex;
typedef std::pair<AtomIndex, AtomIndex> AtomPair;
typedef std::vector<AtomPair> NeighborList;
}
#endif

// Stand-in for openmm/reference/ReferencePairIxn.h -- TEST INFRASTRUCTURE ONLY.
#ifndef NBS_STUB_REFERENCE_PAIR_IXN_H_
#define NBS_STUB_REFERENCE_PAIR_IXN_H_
#include "openmm/Vec3.h"
#include <set>
#include <vector>
#endif

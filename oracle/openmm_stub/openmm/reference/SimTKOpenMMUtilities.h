// Stand-in for openmm/reference/SimTKOpenMMUtilities.h -- TEST INFRASTRUCTURE ONLY.
#ifndef NBS_STUB_SIMTK_UTILITIES_H_
#define NBS_STUB_SIMTK_UTILITIES_H_
#include "openmm/reference/SimTKOpenMMRealType.h"
#endif

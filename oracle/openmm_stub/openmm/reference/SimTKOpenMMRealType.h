// Stand-in for openmm/reference/SimTKOpenMMRealType.h -- TEST INFRASTRUCTURE ONLY.
// Constants are the CODATA-2018 values OpenMM 8.x uses.
#ifndef NBS_STUB_SIMTK_REAL_TYPE_H_
#define NBS_STUB_SIMTK_REAL_TYPE_H_
#include <cmath>
#define PI_M          3.14159265358979323846
#define SQRT_TWO      1.41421356237309504
#define E_CHARGE      (1.602176634e-19)
#define AVOGADRO      (6.02214076e23)
#define EPSILON0      (1e-6*8.8541878128e-12/(E_CHARGE*E_CHARGE*AVOGADRO))
#define ONE_4PI_EPS0  (1/(4*PI_M*EPSILON0))
#define EXP   exp
#define SQRT  sqrt
#define POW   pow
#endif

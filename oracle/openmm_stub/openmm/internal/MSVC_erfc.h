// Stand-in for openmm/internal/MSVC_erfc.h: nothing to do on a C++11 libm.

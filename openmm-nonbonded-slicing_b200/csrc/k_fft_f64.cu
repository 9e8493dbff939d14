// double-precision instantiation of the plane-fused FFT kernels (evaluations that deliver slice energies)
#define NBS_FFT_REAL double
#include "k_fft.inl"

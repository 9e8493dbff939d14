// single-precision instantiation of the plane-fused FFT kernels (forces-only evaluations)
#define NBS_FFT_REAL float
#include "k_fft.inl"

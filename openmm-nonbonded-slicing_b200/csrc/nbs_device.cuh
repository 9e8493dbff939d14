// nbs_device.cuh -- small device-side helpers shared by the kernels.
#ifndef NBS_DEVICE_CUH_
#define NBS_DEVICE_CUH_
#include <cuda_runtime.h>
#include <cstdint>

namespace nbs {

constexpr unsigned FULL_MASK = 0xffffffffu;

// forces are accumulated as 64-bit fixed point (value * 2^32), like OpenMM's CUDA force buffer
// (pme.cc:382-388): integer adds commute, so the sums are independent of execution order.
__device__ __forceinline__ unsigned long long toFixed(float v) {
    return (unsigned long long) __float2ll_rn(v*4294967296.0f);
}
__device__ __forceinline__ unsigned long long toFixed(double v) {
    return (unsigned long long) __double2ll_rn(v*4294967296.0);
}
__device__ __forceinline__ double fromFixed(unsigned long long v) {
    return (double) (long long) v*(1.0/4294967296.0);
}

__device__ __forceinline__ int triSlice(int a, int b) {          // SlicedNonbondedForce.h:22
    return a > b ? a*(a+1)/2 + b : b*(b+1)/2 + a;
}

__device__ __forceinline__ float warpSum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// splitmix64 finaliser -- must equal nbs_pair_hash() in include/nbslice_b200.h
__device__ __forceinline__ unsigned long long pairHash(unsigned first, unsigned second) {
    unsigned long long x = ((unsigned long long) first << 32) | second;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30))*0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27))*0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Order-5 cardinal B-spline weights and derivatives for fractional offset w, following the Darden
// recursion the reference uses (ReferencePME.cpp:280-314), in single precision.
template <typename T>
__device__ __forceinline__ void bspline5(T dr, T* data, T* ddata) {
    data[4] = (T) 0; data[1] = dr; data[0] = (T) 1 - dr;
    // k = 3
    data[2] = (T) 0.5*dr*data[1];
    data[1] = (T) 0.5*((dr+(T) 1)*data[0] + ((T) 2-dr)*data[1]);
    data[0] = (T) 0.5*((T) 1-dr)*data[0];
    // k = 4
    const T third = (T) (1.0/3.0);
    data[3] = third*dr*data[2];
    data[2] = third*((dr+(T) 1)*data[1] + ((T) 3-dr)*data[2]);
    data[1] = third*((dr+(T) 2)*data[0] + ((T) 2-dr)*data[1]);
    data[0] = third*((T) 1-dr)*data[0];
    // differentiate
    ddata[0] = -data[0];
    ddata[1] = data[0] - data[1];
    ddata[2] = data[1] - data[2];
    ddata[3] = data[2] - data[3];
    ddata[4] = data[3] - data[4];
    // k = 5
    data[4] = (T) 0.25*dr*data[3];
    data[3] = (T) 0.25*((dr+(T) 1)*data[2] + ((T) 4-dr)*data[3]);
    data[2] = (T) 0.25*((dr+(T) 2)*data[1] + ((T) 3-dr)*data[2]);
    data[1] = (T) 0.25*((dr+(T) 3)*data[0] + ((T) 2-dr)*data[1]);
    data[0] = (T) 0.25*((T) 1-dr)*data[0];
}

} // namespace nbs
#endif

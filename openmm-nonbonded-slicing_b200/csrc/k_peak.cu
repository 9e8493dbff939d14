// k_peak.cu -- measured instruction-rate ceilings for the pair kernel's roofline (diagnostics only).
//
// The pair kernel is bound by the FP32 pipe / the issue rate, and MEASURED_PEAKS.json carries no FP32 figure,
// so the denominator of its roofline fraction is measured here, on the device the benchmark runs on:
//   * FFMA: every thread runs 16 independent fused-multiply-add chains (nothing depends on the previous
//     instruction of its own chain for 16 issue slots), 8 warps x 4 CTAs per SM, all SMs -- the dense-FMA peak;
//   * MUFU: the same shape with rsqrt.approx (the special-function unit the pair loop uses 3-4 times per pair).
#include "nbs_internal.h"

namespace nbs {

constexpr int PEAK_CHAINS = 16;
constexpr int PEAK_INNER = 64;

__global__ void __launch_bounds__(256) k_peak_ffma(int rounds, float seed, float* sink) {
    float acc[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) acc[k] = seed + (float) (threadIdx.x + k);
    const float a = 1.0f + 1e-7f*seed, b = 1e-3f*seed;
    for (int r = 0; r < rounds; r++) {
#pragma unroll
        for (int u = 0; u < PEAK_INNER; u++) {
#pragma unroll
            for (int k = 0; k < PEAK_CHAINS; k++) acc[k] = fmaf(acc[k], a, b);
        }
    }
    float total = 0.f;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) total += acc[k];
    if (total == 12345.678f) sink[0] = total;            // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) k_peak_mufu(int rounds, float seed, float* sink) {
    float acc[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) acc[k] = 1.0f + seed + 0.01f*(float) (threadIdx.x + k);
    for (int r = 0; r < rounds; r++) {
#pragma unroll
        for (int u = 0; u < PEAK_INNER/4; u++) {
#pragma unroll
            for (int k = 0; k < PEAK_CHAINS; k++) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(acc[k]));
        }
    }
    float total = 0.f;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) total += acc[k];
    if (total == 12345.678f) sink[0] = total;
}

// double-precision and conversion rates (the energy passes of the pair kernel): WHICH = 0 DFMA, 1 int32 -> double,
// 2 float -> double, 3 double -> float, 4 int64 -> double
template <int WHICH>
__global__ void __launch_bounds__(256) k_peak_dp(int rounds, double seed, double* sink) {
    double acc[8];
    int iv[8];
    float fv[8];
    long long lv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = seed + threadIdx.x + k; iv[k] = threadIdx.x*7 + k; fv[k] = (float) (seed + k); lv[k] = (long long) threadIdx.x*77777 + k; }
    const double a = 1.0 + 1e-9*seed, b = 1e-3*seed;
    for (int r = 0; r < rounds; r++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (WHICH == 0) acc[k] = fma(acc[k], a, b);
                else if (WHICH == 1) { double d = (double) iv[k]; iv[k] = __double2hiint(d) ^ (iv[k] + u); }
                else if (WHICH == 2) { double d = (double) fv[k]; fv[k] = __int_as_float((__double2hiint(d) & 0x007fffff) | 0x3f800000); }
                else if (WHICH == 3) { float f = (float) acc[k]; acc[k] = __hiloint2double(__float_as_int(f), u); }
                else { double d = (double) lv[k]; lv[k] = (long long) __double2hiint(d)*3 + lv[k]; }
            }
        }
    }
    double total = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) total += acc[k] + iv[k] + fv[k] + (double) lv[k];
    if (total == 12345.678) sink[0] = total;
}

} // namespace nbs

using namespace nbs;

// out[0..4]: warp-instructions per clock per SM of DFMA, I2F.F64.S32, F2F.F64.F32, F2F.F32.F64, I2F.F64.S64 (an upper
// bound for the conversions, whose loops carry a couple of integer instructions per conversion)
extern "C" int nbs_measure_dp_rates(int32_t device, double out[8]) {
    if (!out) { setError("null argument"); return NBS_ERR_INVALID; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        setError("no CUDA device available (this library has no CPU fallback)");
        return NBS_ERR_CUDA;
    }
    NBS_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    NBS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    double* sink = nullptr;
    NBS_CUDA_CHECK(cudaMalloc((void**) &sink, sizeof(double)));
    cudaEvent_t e0, e1;
    NBS_CUDA_CHECK(cudaEventCreate(&e0));
    NBS_CUDA_CHECK(cudaEventCreate(&e1));
    const int ctas = prop.multiProcessorCount*4, rounds = 64;
    for (int which = 0; which < 5; which++) {
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0, 0);
            switch (which) {
                case 0: k_peak_dp<0><<<ctas, 256>>>(rounds, 0.5, sink); break;
                case 1: k_peak_dp<1><<<ctas, 256>>>(rounds, 0.5, sink); break;
                case 2: k_peak_dp<2><<<ctas, 256>>>(rounds, 0.5, sink); break;
                case 3: k_peak_dp<3><<<ctas, 256>>>(rounds, 0.5, sink); break;
                default: k_peak_dp<4><<<ctas, 256>>>(rounds, 0.5, sink); break;
            }
            cudaEventRecord(e1, 0);
            NBS_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double warpInstr = (double) rounds*16*8*8*ctas;            // 8 warps per CTA
            const double perClockPerSM = warpInstr/(ms*1e-3)/(prop.clockRate*1e3)/prop.multiProcessorCount;
            if (rep >= 1 && perClockPerSM > best) best = perClockPerSM;
        }
        out[which] = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    return NBS_OK;
}

extern "C" int nbs_measure_peaks(int32_t device, double out[4]) {
    if (!out) { setError("null argument"); return NBS_ERR_INVALID; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        setError("no CUDA device available (this library has no CPU fallback)");
        return NBS_ERR_CUDA;
    }
    NBS_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    NBS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    float* sink = nullptr;
    NBS_CUDA_CHECK(cudaMalloc((void**) &sink, sizeof(float)));
    cudaEvent_t e0, e1;
    NBS_CUDA_CHECK(cudaEventCreate(&e0));
    NBS_CUDA_CHECK(cudaEventCreate(&e1));
    const int ctas = prop.multiProcessorCount*4, rounds = 512;
    double best[2] = {0, 0};
    for (int which = 0; which < 2; which++)
        for (int rep = 0; rep < 6; rep++) {            // first repetitions warm the clocks up
            cudaEventRecord(e0, 0);
            if (which == 0) k_peak_ffma<<<ctas, 256>>>(rounds, 0.5f, sink);
            else k_peak_mufu<<<ctas, 256>>>(rounds, 0.5f, sink);
            cudaEventRecord(e1, 0);
            NBS_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double perThread = (double) rounds*(which == 0 ? PEAK_INNER : PEAK_INNER/4)*PEAK_CHAINS;
            const double rate = perThread*256.0*ctas/(ms*1e-3);
            if (rep >= 2 && rate > best[which]) best[which] = rate;
        }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    out[0] = 2.0*best[0]*1e-12;      // TFLOP/s (an FMA counts as two flops)
    out[1] = best[1]*1e-9;           // G rsqrt/s
    out[2] = prop.multiProcessorCount;
    out[3] = prop.clockRate*1e-3;    // MHz (the device's nominal maximum)
    return NBS_OK;
}

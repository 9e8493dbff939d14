// k_final.cu -- convert the fixed-point force accumulators (sorted order) into the caller's layout.
#include "nbs_internal.h"
#include "nbs_device.cuh"

namespace nbs {

__global__ void k_finalize_f64(int N, int Npad, const unsigned long long* __restrict__ force,
                               const int* __restrict__ origToSorted, const int* __restrict__ atomIndex,
                               double* __restrict__ out, int accumulate) {
    const int slot = blockIdx.x*blockDim.x + threadIdx.x;
    if (slot >= N) return;
    const int s = origToSorted[atomIndex ? atomIndex[slot] : slot];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double f = fromFixed(force[(size_t) c*Npad + s]);
        out[3*(size_t) slot + c] = accumulate ? out[3*(size_t) slot + c] + f : f;
    }
}

// OpenMM CUDA's long-long force buffer: [3][paddedAtoms], value * 2^32, always accumulated.
__global__ void k_finalize_i64(int N, int Npad, const unsigned long long* __restrict__ force,
                               const int* __restrict__ origToSorted, const int* __restrict__ atomIndex,
                               unsigned long long* __restrict__ out, long long paddedAtoms) {
    const int slot = blockIdx.x*blockDim.x + threadIdx.x;
    if (slot >= N) return;
    const int s = origToSorted[atomIndex ? atomIndex[slot] : slot];
#pragma unroll
    for (int c = 0; c < 3; c++) out[(size_t) c*paddedAtoms + slot] += force[(size_t) c*Npad + s];
}

int launchFinalize(Context& c, void* dOut, int format, long long paddedAtoms, int accumulate, const int* atomIndex) {
    const int T = 256;
    if (format == NBS_FORCE_F64_XYZ)
        k_finalize_f64<<<(c.N+T-1)/T, T, 0, c.stream>>>(c.N, c.Npad, c.dForce.d, c.dOrigToSorted.d, atomIndex, (double*) dOut, accumulate);
    else
        k_finalize_i64<<<(c.N+T-1)/T, T, 0, c.stream>>>(c.N, c.Npad, c.dForce.d, c.dOrigToSorted.d, atomIndex,
                                                         (unsigned long long*) dOut, paddedAtoms);
    c.launches++;
    timerMark(c, "finalize");
    return NBS_OK;
}

} // namespace nbs

// k_final.cu -- convert the fixed-point force accumulators (sorted order) into the caller's layout.
#include "nbs_internal.h"
#include "nbs_device.cuh"

namespace nbs {

// `force` holds [3][Npad] in cell-sorted order followed (when `both`) by [3][Npad] in particle order.
__global__ void k_finalize_f64(int N, int Npad, const unsigned long long* __restrict__ force,
                               const int* __restrict__ origToSorted, const int* __restrict__ atomIndex,
                               double* __restrict__ out, int accumulate, int both) {
    const int slot = blockIdx.x*blockDim.x + threadIdx.x;
    if (slot >= N) return;
    const int particle = atomIndex ? atomIndex[slot] : slot;
    const int s = origToSorted[particle];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        unsigned long long v = force[(size_t) c*Npad + s];
        if (both) v += force[(size_t) (3 + c)*Npad + particle];
        const double f = fromFixed(v);
        out[3*(size_t) slot + c] = accumulate ? out[3*(size_t) slot + c] + f : f;
    }
}

// OpenMM CUDA's long-long force buffer: [3][paddedAtoms], value * 2^32, always accumulated.
__global__ void k_finalize_i64(int N, int Npad, const unsigned long long* __restrict__ force,
                               const int* __restrict__ origToSorted, const int* __restrict__ atomIndex,
                               unsigned long long* __restrict__ out, long long paddedAtoms, int both) {
    const int slot = blockIdx.x*blockDim.x + threadIdx.x;
    if (slot >= N) return;
    const int particle = atomIndex ? atomIndex[slot] : slot;
    const int s = origToSorted[particle];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        unsigned long long v = force[(size_t) c*Npad + s];
        if (both) v += force[(size_t) (3 + c)*Npad + particle];
        out[(size_t) c*paddedAtoms + slot] += v;
    }
}

int launchFinalize(Context& c, void* dOut, int format, long long paddedAtoms, int accumulate, const int* atomIndex) {
    const int T = 256;
    if (format == NBS_FORCE_F64_XYZ)
        k_finalize_f64<<<(c.N+T-1)/T, T, 0, c.stream>>>(c.N, c.Npad, c.dForce.d, c.dOrigToSorted.d, atomIndex, (double*) dOut, accumulate, c.pmeUnsorted ? 1 : 0);
    else
        k_finalize_i64<<<(c.N+T-1)/T, T, 0, c.stream>>>(c.N, c.Npad, c.dForce.d, c.dOrigToSorted.d, atomIndex,
                                                         (unsigned long long*) dOut, paddedAtoms, c.pmeUnsorted ? 1 : 0);
    c.launches++;
    timerMark(c, "finalize");
    return NBS_OK;
}

} // namespace nbs

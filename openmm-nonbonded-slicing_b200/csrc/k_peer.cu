// k_peer.cu -- the cross-GPU pieces of the peer-memory sharding (include/nbslice_b200.h, nbs_set_slab_shard): a barrier
// over flags in the ranks' mailboxes and the force reduction over NVLink peer memory.  No collective library is
// involved: every rank maps the other ranks' spectra, force accumulators and mailboxes (CUDA IPC) and the kernels load
// and store through those pointers.  The reference has nothing comparable -- its multi-device path leaves reciprocal
// space on device 0 and sums forces on the host side of OpenMM's CudaParallelKernels
// (platforms/cuda/src/CudaParallelNonbondedSlicingKernels.cpp:35-53).
#include "nbs_internal.h"
#include "nbs_device.cuh"
#include <algorithm>

namespace nbs {

struct PeerBoxes { PeerMailbox* box[NBS_MAX_RANKS]; };
struct PeerForces { unsigned long long* f[NBS_MAX_RANKS]; };

__device__ __forceinline__ void storeReleaseSys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long loadAcquireSys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One CTA.  Everything this rank's earlier kernels wrote (to its own memory or a peer's) is made visible, then thread p
// tells rank p "rank `rank` has reached barrier number `epoch`" and waits until rank p has said the same here.  The
// epoch is counted on the device, so a CUDA graph can replay the kernel.  With `energy` the rank first publishes its
// slice-energy table in every mailbox (k_peer_reduce sums them in rank order, so all ranks get identical totals).
// A barrier that waits longer than `timeoutCycles` gives up and flags the mailbox: the host turns that into an error
// instead of a hung GPU.
__global__ void k_peer_barrier(PeerBoxes t, int rank, int R, int wait, const double* __restrict__ energy, long long timeoutCycles) {
    __shared__ unsigned long long epochS;
    PeerMailbox* mine = t.box[rank];
    if (energy != nullptr)
        for (int k = threadIdx.x; k < R*ENERGY_WORDS; k += blockDim.x) {
            const int p = k/ENERGY_WORDS, w = k - p*ENERGY_WORDS;
            t.box[p]->energies[rank][w] = energy[w];
        }
    __threadfence_system();
    __syncthreads();
    if (!wait) return;                       // the caller orders the ranks' steps itself: publishing was all there is to do
    if (threadIdx.x == 0) epochS = ++mine->epoch;
    __syncthreads();
    const unsigned long long epoch = epochS;
    if (threadIdx.x < R) {
        __threadfence_system();
        storeReleaseSys(&t.box[threadIdx.x]->arrive[rank], epoch);
        const long long start = clock64();
        while (loadAcquireSys(&mine->arrive[threadIdx.x]) < epoch) {
            if (clock64() - start > timeoutCycles) { mine->timedOut = epoch; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    __threadfence_system();
}

// Force reduction: this rank owns the 16-byte words [lo, hi) of the fixed-point accumulators (1/R of them): it reads
// them from every rank, adds (64-bit integer adds commute: the totals are bit-identical whatever the order) and writes
// the sums back to every rank -- reduce-scatter and all-gather in one pass, loads and stores straight over NVLink.
// Only this rank touches words [lo, hi) of any rank's accumulators during the kernel, so the update is in place.
// Block 0 also forms the slice-energy totals from the mailbox (rank order).
__global__ void __launch_bounds__(256) k_peer_reduce(PeerForces t, int R, long long lo, long long hi, const PeerMailbox* __restrict__ mine,
                                                      double* __restrict__ energy) {
    if (blockIdx.x == 0)
        for (int w = threadIdx.x; w < ENERGY_WORDS; w += blockDim.x) {
            double sum = 0.0;
            for (int p = 0; p < R; p++) sum += mine->energies[p][w];
            energy[w] = sum;
        }
    const long long stride = (long long) gridDim.x*blockDim.x;
    for (long long i = lo + (long long) blockIdx.x*blockDim.x + threadIdx.x; i < hi; i += stride) {
        ulonglong2 v[NBS_MAX_RANKS];
#pragma unroll
        for (int p = 0; p < NBS_MAX_RANKS; p++)
            if (p < R) v[p] = reinterpret_cast<const ulonglong2*>(t.f[p])[i];
        ulonglong2 sum = make_ulonglong2(0ull, 0ull);
#pragma unroll
        for (int p = 0; p < NBS_MAX_RANKS; p++)
            if (p < R) { sum.x += v[p].x; sum.y += v[p].y; }
#pragma unroll
        for (int p = 0; p < NBS_MAX_RANKS; p++)
            if (p < R) reinterpret_cast<ulonglong2*>(t.f[p])[i] = sum;
    }
}

int launchPeerBarrier(Context& c, bool publishEnergies) {
    PeerBoxes t;
    for (int r = 0; r < NBS_MAX_RANKS; r++) t.box[r] = r < c.nRanks ? c.peerMailbox[r] : nullptr;
    if (!c.peerBarrier && !publishEnergies) return NBS_OK;
    const long long timeout = 40000000000LL;          // ~20 s at 2 GHz: far beyond any skew between ranks
    k_peer_barrier<<<1, 128, 0, c.stream>>>(t, c.rank, c.nRanks, c.peerBarrier ? 1 : 0, publishEnergies ? c.dEnergy.d : nullptr, timeout);
    c.launches++;
    timerMark(c, "peer_barrier");
    return NBS_OK;
}

int launchPeerReduce(Context& c) {
    PeerForces t;
    for (int r = 0; r < NBS_MAX_RANKS; r++) t.f[r] = r < c.nRanks ? c.peerForce[r] : nullptr;
    const long long words2 = (long long) (c.pmeUnsorted ? 6 : 3)*c.Npad/2;       // Npad is a multiple of 32
    const long long lo = words2*c.rank/c.nRanks, hi = words2*(c.rank + 1)/c.nRanks;
    const int ctas = (int) std::max(1LL, std::min((long long) 4*c.numSMs, (hi - lo + 255)/256));
    k_peer_reduce<<<ctas, 256, 0, c.stream>>>(t, c.nRanks, lo, hi, c.peerMailbox[c.rank], c.dEnergy.d);
    c.launches++;
    timerMark(c, "peer_reduce");
    return NBS_OK;
}

} // namespace nbs

// k_fft.cuh -- argument blocks of the plane-fused FFT kernels (k_fft.cu), shared with k_pme.cu.
#ifndef NBS_K_FFT_CUH_
#define NBS_K_FFT_CUH_
#include "nbs_internal.h"

namespace nbs {

constexpr int FFT_THREADS = 512;     // upper bound of the plane kernels' CTA size (zy forward, yz inverse): two CTAs per SM
                                     // at 64 registers; the launch picks the size that splits the widest pass of the plan
                                     // into the fewest rounds of equal size (measured: 544 threads = one CTA per SM = slower)
constexpr int FFT_X_THREADS = 256;   // upper bound for the x + convolution kernel (more registers per thread)
constexpr int FFT_UNPACK_Q = 4;      // (row pair, kz) items a thread stages per round of the real<->complex (un)packing

struct PlaneFftPlan {
    unsigned long long factors[3];   // radices of nx, ny, nz: 4 bits each, first pass in the low bits
};

struct PlaneFftArgs {
    int nS, nx, ny, nz, nzh;
    int ownLo, ownHi;                // subsets whose grids this rank transforms / produces
    // slab sharding over ranks (peer memory): the zy / yz kernels work on the planes x in [xLo, xLo + nxOwn) of every
    // own subset, the x kernel on the rows y in [yLo, yLo + nyOwn), reading and writing plane x in the spectra of the
    // rank that owns it (peerSpectra[owner(x)], the layout of gridC).  Unsharded: xLo = yLo = 0, nxOwn = nx, nyOwn = ny,
    // nRanks = 1, peerSpectra[0] = gridC.
    int xLo, nxOwn, yLo, nyOwn, nRanks;
    void* peerSpectra[NBS_MAX_RANKS];
    int rowStride;                   // zy / yz kernels: complex elements per plane row in shared memory
    int chunk;                       // x kernel: kz values per CTA
    int planeThreads, xThreads, colThreads;   // CTA sizes chosen for this plan
    int slabPairs, slabsPerPlane;    // zy / yz kernels: row pairs per CTA; one slab per plane = the y transform is fused in
    int colChunk;                    // split path: kz columns per CTA of the y kernel
    unsigned long long factorsX, factorsY, factorsZ;
    const void* twx; const void* twy; const void* twz;     // exp(-2 pi i k / n), precision T
    const void* grid;                // real charge grids   [nS][nx][ny][nz]      (T)
    void* gridC;                     // half spectra        [nS][nx][ny][nz/2+1]  (complex T)
    const void* eterm;               // influence function  [nx][ny][nz/2+1]      (T)
    float* pot;                      // potential grids     [nS][nx][ny][nz]      (float; double when potDouble)
    int potDouble;
    double* energy;
    int wantEnergy;
    LambdaTable lam;
};

// NBS_OK: launched.  NBS_RETRY: the grid does not fit this path (use the line-at-a-time kernels).
template <typename T>
int launchPlaneFft(Context& c, const PlaneFftPlan& plan, PlaneFftArgs a, int half);

} // namespace nbs
#endif

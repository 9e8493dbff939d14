// k_pme.cu -- sliced smooth-PME reciprocal space: per-subset B-spline spreading, a hand-written
// 3D real FFT, the sliced reciprocal convolution (cross-subset structure-factor products per slice),
// and the force gather.
//
// Arithmetic follows platforms/reference/src/ReferencePME.cpp (order 5, :754-811): index/fraction
// :196-256, splines :264-317, spread :320-396, convolution + slice energies :400-496, gather :598-702.
// Differences in STRUCTURE (not in results), all B200-motivated:
//   * charges carry sqrt(ONE_4PI_EPS0) so the influence function needs no unit factor;
//   * real-to-complex transforms on half spectra instead of the reference's complex-to-complex;
//   * the x-direction forward FFT, the convolution, the slice energies, the lambda mixing of the
//     subset potentials (G_I = eterm * sum_J lambda_IJ S_J, legal because everything is linear) and the
//     x-direction inverse FFT are ONE kernel that keeps the lines of all subsets in shared memory; the
//     gather then reads 125 points from one grid per atom instead of 125 * nSubsets (pme.cc:360-371);
//   * FFT lines live in shared memory, one warp per line, Stockham passes staged through registers
//     (radices 4, 2, 3, 5, 7, 11, 13 -- the factor set the reference's VkFFT path accepts,
//     platforms/common/include/FFT3DFactory.h:45-47).
// Bound: HBM/L2 bandwidth and launch latency (grids of every BASELINE config fit in the 126 MB L2).
#include "nbs_internal.h"
#include "nbs_device.cuh"
#include "k_fft.cuh"

namespace nbs {

constexpr int FFT_MAX_N = 512;

struct FftPlan {           // factorisation of one dimension
    int n, nf, nq;         // nq = butterflies per lane of the widest pass = ceil(max_R (n/R) / 32)
    int f[12];
    unsigned long long packed;   // the factors, 4 bits each, first pass in the low bits, 0-terminated
};

static bool makePlan(int n, FftPlan& p) {
    p.n = n; p.nf = 0; p.nq = 1;
    int m = n;
    while (m % 4 == 0) { p.f[p.nf++] = 4; m /= 4; }
    const int radices[] = {2, 3, 5, 7, 11, 13};
    for (int r : radices)
        while (m % r == 0) { p.f[p.nf++] = r; m /= r; }
    p.packed = 0;
    for (int k = 0; k < p.nf; k++) {
        if (p.f[k] <= 7) p.nq = std::max(p.nq, (n/p.f[k] + 31)/32);      // radix 11/13 passes are rolled loops
        p.packed |= (unsigned long long) p.f[k] << (4*k);
    }
    return m == 1 && n <= FFT_MAX_N;
}

// complex helpers, generic in precision
template <typename T> struct Cx;
template <> struct Cx<float> { typedef float2 type; };
template <> struct Cx<double> { typedef double2 type; };
__device__ __forceinline__ float2 mk(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ double2 mk(double x, double y) { return make_double2(x, y); }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) { return mk(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }

// ---------------------------------------------------------------------------------------------
// One Stockham pass of radix R over a line of length n in shared memory, executed by one warp.
// Inputs are staged through registers, so the pass is in place (read all, sync, write all).
// NQ = butterflies per lane the instantiation is unrolled for.
// ---------------------------------------------------------------------------------------------
template <int R, int NQ, typename C>
__device__ __forceinline__ void fftPass(C* line, int n, int Ns, const C* tw, int lane) {
    const int nb = n/R;
    const int tstep = n/(Ns*R);
    const int rstep = n/R;
    C v[NQ][R];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int j = lane + 32*q;
        if (j < nb) {
#pragma unroll
            for (int t = 0; t < R; t++) v[q][t] = line[j + t*nb];
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int j = lane + 32*q;
        if (j < nb) {
            const int k = j % Ns;
#pragma unroll
            for (int t = 1; t < R; t++) v[q][t] = cmul(v[q][t], tw[t*k*tstep]);
            const int j0 = (j/Ns)*Ns*R + k;
            if (R == 2) {
                line[j0] = mk(v[q][0].x + v[q][1].x, v[q][0].y + v[q][1].y);
                line[j0 + Ns] = mk(v[q][0].x - v[q][1].x, v[q][0].y - v[q][1].y);
            }
            else if (R == 4) {
                const C a0 = mk(v[q][0].x + v[q][2].x, v[q][0].y + v[q][2].y);
                const C a1 = mk(v[q][0].x - v[q][2].x, v[q][0].y - v[q][2].y);
                const C a2 = mk(v[q][1].x + v[q][3].x, v[q][1].y + v[q][3].y);
                const C a3 = mk(v[q][1].x - v[q][3].x, v[q][1].y - v[q][3].y);
                // forward transform: multiply a3 by -i
                line[j0] = mk(a0.x + a2.x, a0.y + a2.y);
                line[j0 + Ns] = mk(a1.x + a3.y, a1.y - a3.x);
                line[j0 + 2*Ns] = mk(a0.x - a2.x, a0.y - a2.y);
                line[j0 + 3*Ns] = mk(a1.x - a3.y, a1.y + a3.x);
            }
            else {
#pragma unroll
                for (int o = 0; o < R; o++) {
                    C acc = v[q][0];
#pragma unroll
                    for (int t = 1; t < R; t++) {
                        const C w = tw[((o*t) % R)*rstep];
                        acc.x += v[q][t].x*w.x - v[q][t].y*w.y;
                        acc.y += v[q][t].x*w.y + v[q][t].y*w.x;
                    }
                    line[j0 + o*Ns] = acc;
                }
            }
        }
    }
    __syncwarp();
}

// Radix 11 / 13 (rare grid sizes): rolled loops over small local arrays -- slow but register-light,
// so the common radices keep their occupancy.  n <= 512 means at most 46 butterflies, i.e. at most two
// per lane; both are read before anything is written (the pass is in place).
template <typename C>
__device__ __noinline__ void fftPassLarge(C* line, int n, int R, int Ns, const C* tw, int lane) {
    const int nb = n/R, tstep = n/(Ns*R), rstep = n/R;
    C v[2][13];
    for (int q = 0; q < 2; q++) {
        const int j = lane + 32*q;
        if (j < nb)
            for (int t = 0; t < R; t++) v[q][t] = line[j + t*nb];
    }
    __syncwarp();
    for (int q = 0; q < 2; q++) {
        const int j = lane + 32*q;
        if (j >= nb) continue;
        const int k = j % Ns;
        for (int t = 1; t < R; t++) v[q][t] = cmul(v[q][t], tw[t*k*tstep]);
        const int j0 = (j/Ns)*Ns*R + k;
        for (int o = 0; o < R; o++) {
            C acc = v[q][0];
            for (int t = 1; t < R; t++) {
                const C z = tw[((o*t) % R)*rstep];
                acc.x += v[q][t].x*z.x - v[q][t].y*z.y;
                acc.y += v[q][t].x*z.y + v[q][t].y*z.x;
            }
            line[j0 + o*Ns] = acc;
        }
    }
    __syncwarp();
}

// Forward (e^{-i...}) unnormalised FFT of one shared-memory line by one warp.
template <int NQ, typename C>
__device__ __forceinline__ void warpFft(C* line, int n, unsigned long long factors, const C* tw, int lane) {
    int Ns = 1;
    for (; factors != 0; factors >>= 4) {
        const int R = (int) (factors & 15);
        switch (R) {
            case 2: fftPass<2, NQ>(line, n, Ns, tw, lane); break;
            case 3: fftPass<3, NQ>(line, n, Ns, tw, lane); break;
            case 4: fftPass<4, NQ>(line, n, Ns, tw, lane); break;
            case 5: fftPass<5, NQ>(line, n, Ns, tw, lane); break;
            case 7: fftPass<7, NQ>(line, n, Ns, tw, lane); break;
            default: fftPassLarge(line, n, R, Ns, tw, lane); break;
        }
        Ns *= R;
    }
}

// ---------------------------------------------------------------------------------------------
// Spreading: one warp per (sorted) atom; lanes 0..24 own an (ix, iy) offset and walk the 5 z points.
// Reference: pme_grid_spread_charge, ReferencePME.cpp:320-396 (forward-only spreading, :375-394).
// T = float for force-only evaluations, double when slice energies are requested (cross-subset
// structure-factor products cancel to ~1e-6 of their terms; see DESIGN.md "Precision").
// ---------------------------------------------------------------------------------------------
struct PmeArgs {
    int N, Npad, nS, nx, ny, nz, nzh;
    int ownLo, ownHi;            // this rank spreads / gathers the atoms of subsets [ownLo, ownHi)
    int xLo, xHi;                // ... into / from the grid planes x in [xLo, xHi) (slab sharding; the whole grid otherwise)
    // slab sharding, cell-sorted atoms, rectangular box: only the atoms of the cell columns that can reach the slab are
    // looked at -- sorted atoms [binStart[rangeBin[0]], binStart[rangeBin[1]]) and [binStart[rangeBin[2]],
    // binStart[rangeBin[3]]) (the second range is the periodic wrap); rangeBin[0] < 0: all atoms
    const int* binStart; int rangeBin[4];
    // ... and once the host knows the two ranges (they come back with the counters of the evaluation that sorted the
    // atoms) the grid covers exactly those atoms: warp j works on sorted atom hostRange[0] + j, or hostRange[2] + (j - length
    // of the first range)
    int useHostRange; int hostRange[4];
    int batch;                   // atoms per warp in k_spread / k_gather (4 .. 32, a power of two)
    const uint4* posq; const float4* par;
    // unsorted mode (small systems): particle-order inputs straight from k_prep, so that the PME chain does not
    // wait for the cell sort; forces then go to the particle-order half of the accumulator
    int unsorted;
    const uint4* fix; const float* chargeF; const int* subsetOf;
    const double* q64; const double* chargeD; double sqrtK;      // double-precision charges for the double-precision grids
    void* grid; const float* pot;
    unsigned long long* gridFixed;   // deterministic mode: 64-bit fixed-point accumulation grid
    unsigned long long* force;
    float fscale[3];             // n_d / L_d
    // triclinic box (a = (ax,0,0), b = (bx,by,0), c = (cx,cy,cz)): lattice fractions from the brick fractions (u, v, w)
    // the fixed-point coordinates hold, t_x = u - beta v + delta w, t_y = v - gammaY w, t_z = w (ReferencePME.cpp:
    // 186-194, 249-251), and the off-diagonal reciprocal-vector terms of the force (:698-700)
    int triclinic;
    double beta, gammaY, delta;  // bx/ax, cy/by, (bx cy - by cx)/(ax by)
    float fr10, fr20, fr21;      // nx r10, nx r20, ny r21
    double fscaleD[3], fr10D, fr20D, fr21D;      // the same in double (NBS_FLAG_DOUBLE)
};

// Brick fractions -> lattice fractions (both 32-bit fixed point); identity for a rectangular box.
__device__ __forceinline__ uint4 latticeFractions(const PmeArgs& a, uint4 p) {
    if (!a.triclinic) return p;
    const double s = 1.0/4294967296.0;
    const double u = p.x*s, v = p.y*s, w = p.z*s;
    double tx = u - a.beta*v + a.delta*w, ty = v - a.gammaY*w;
    tx -= floor(tx); ty -= floor(ty);
    p.x = (unsigned) (__double2ull_rd(tx*4294967296.0) & 0xffffffffull);
    p.y = (unsigned) (__double2ull_rd(ty*4294967296.0) & 0xffffffffull);
    return p;
}

// Slab sharding: is sorted atom j in one of the two ranges of cell columns that can reach this rank's planes?  (Every
// warp of the grid asks; the four bounds are the same words for all of them and stay in L1.)
__device__ __forceinline__ bool inSlabRanges(const PmeArgs& a, int j) {
    if (a.rangeBin[0] < 0) return true;
    const int a0 = __ldg(a.binStart + a.rangeBin[0]), a1 = __ldg(a.binStart + a.rangeBin[1]);
    const int b0 = __ldg(a.binStart + a.rangeBin[2]), b1 = __ldg(a.binStart + a.rangeBin[3]);
    return (j >= a0 && j < a1) || (j >= b0 && j < b1);
}

// The sorted atom warp j of the grid works on, or -1.
__device__ __forceinline__ int slabAtom(const PmeArgs& a, int j) {
    if (j < 0) return -1;
    if (a.useHostRange) {
        const int lenA = a.hostRange[1] - a.hostRange[0];
        if (j < lenA) return a.hostRange[0] + j;
        const int atom = a.hostRange[2] + (j - lenA);
        return atom < a.hostRange[3] ? atom : -1;
    }
    return (j < a.N && inSlabRanges(a, j)) ? j : -1;
}

__global__ void k_slab_ranges(const int* __restrict__ binStart, int b0, int b1, int b2, int b3, int* __restrict__ out) {
    if (threadIdx.x == 0) { out[0] = binStart[b0]; out[1] = binStart[b1]; out[2] = binStart[b2]; out[3] = binStart[b3]; }
}

// Slab sharding: does any of the five x planes of the atom's spline fall into this rank's slab?  (warp-uniform)
__device__ __forceinline__ bool touchesSlab(const PmeArgs& a, const uint4 pBrick) {
    if (a.xHi - a.xLo >= a.nx) return true;
    const uint4 p = latticeFractions(a, pBrick);
    const int ix0 = (int) (((unsigned long long) p.x*(unsigned) a.nx) >> 32);
#pragma unroll
    for (int o = 0; o < PME_ORDER; o++) {
        int x = ix0 + o; x -= x >= a.nx ? a.nx : 0;
        if (x >= a.xLo && x < a.xHi) return true;
    }
    return false;
}

template <typename T>
__device__ __forceinline__ void gridCoord(unsigned fixed, int n, int& index, T& frac) {
    const unsigned long long t = (unsigned long long) fixed*(unsigned) n;     // frac * n in 32.32 fixed point
    index = (int) (t >> 32);
    frac = (T) (unsigned) (t & 0xffffffffull)*(T) (1.0/4294967296.0);
}

// Spreading and gather work in BATCHES of 32 atoms per warp.  Phase 1, lane = atom: coalesced loads, the three order-5
// splines (about 200 arithmetic instructions -- done once per 32 atoms, every lane busy; one warp per atom paid them per
// atom with 15 lanes busy) into the warp's shared-memory table, k-major with a row stride of 33 so that both phases are
// free of bank conflicts.  Phase 2, one atom at a time: the 125 grid points are dealt to the lanes with z fastest
// (point = lane + 32 i, its (ox, oy, oz) fixed per lane for the whole kernel), so one warp-wide atomic / load touches
// runs of 5 consecutive cells of a grid row -- about a third of the L2 transactions of a row-per-lane assignment.
// Per atom that leaves four passes of ~16 instructions: ~75 warp instructions instead of 340 (spread) / 460 (gather).
constexpr int PME_TAB_STRIDE = 33;
template <typename T, bool DERIV>
struct __align__(16) PmeWarpTab {
    T w[15][PME_TAB_STRIDE];           // spline weights: [0..4] x, [5..9] y, [10..14] z
    T dw[DERIV ? 15 : 1][PME_TAB_STRIDE];          // their derivatives (gather)
    T q[32];
    int ix0[32], iy0[32], iz0[32], subset[32], atom[32];
};

// Phase 1 for lane's atom; returns the ballot of the lanes that hold an atom with work to do.
template <typename T, bool DERIV>
__device__ __forceinline__ unsigned pmeBatchSetup(const PmeArgs& a, PmeWarpTab<T, DERIV>& t, int lane, int mapped) {
    const int j = slabAtom(a, mapped);
    bool valid = j >= 0;
    uint4 p = make_uint4(0u, 0u, 0u, 0u); int subset = 0; float q = 0.f;
    if (valid) {
        if (a.unsorted) { p = a.fix[j]; q = a.chargeF[j]; subset = a.subsetOf[j]; }
        else { p = a.posq[j]; q = __uint_as_float(p.w); subset = __float_as_int(a.par[j].z) & 7; }
        valid = q != 0.f && subset >= a.ownLo && subset < a.ownHi && touchesSlab(a, p);
    }
    if (valid) {
        const uint4 f = latticeFractions(a, p);
        int index; T frac, th[5], dth[5];
        gridCoord<T>(f.x, a.nx, index, frac); t.ix0[lane] = index;
        bspline5(frac, th, dth);
#pragma unroll
        for (int k = 0; k < 5; k++) { t.w[k][lane] = th[k]; if (DERIV) t.dw[k][lane] = dth[k]; }
        gridCoord<T>(f.y, a.ny, index, frac); t.iy0[lane] = index;
        bspline5(frac, th, dth);
#pragma unroll
        for (int k = 0; k < 5; k++) { t.w[5 + k][lane] = th[k]; if (DERIV) t.dw[5 + k][lane] = dth[k]; }
        gridCoord<T>(f.z, a.nz, index, frac); t.iz0[lane] = index;
        bspline5(frac, th, dth);
#pragma unroll
        for (int k = 0; k < 5; k++) { t.w[10 + k][lane] = th[k]; if (DERIV) t.dw[10 + k][lane] = dth[k]; }
        // (the fp32 charge carries 6e-8 of rounding: visible in cross-subset energies that cancel to 1e-6 of their terms)
        t.q[lane] = sizeof(T) == 8 ? (T) (a.unsorted ? a.chargeD[j]*a.sqrtK : a.q64[j]) : (T) q;
        t.subset[lane] = subset;
        t.atom[lane] = j;
    }
    const unsigned todo = __ballot_sync(FULL_MASK, valid);
    __syncwarp();
    return todo;
}

// FIXED (NBS_FLAG_DETERMINISTIC; the plugin's CudaDeterministicForces, pme.cc:108-109, 124-134): the grid points are
// accumulated as 64-bit fixed point -- integer adds commute, so the grid, and with it every reciprocal-space force,
// is bit-reproducible whatever order the atomics retire in -- and k_fixed_to_real converts the grid afterwards.
template <typename T> struct FixedGridScale { static constexpr double value = 4294967296.0; };            // 2^32: 2e-10 of a unit charge
template <> struct FixedGridScale<double> { static constexpr double value = 1099511627776.0; };          // 2^40 for the double-precision grids

// (blockIdx.y = grid of a subset; `n` consecutive cells of each, `stride` cells apart)
template <typename T>
__global__ void k_fixed_to_real(size_t n, size_t stride, const long long* __restrict__ fixed, T* __restrict__ grid) {
    const size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    const size_t at = blockIdx.y*stride + i;
    if (i < n) grid[at] = (T) ((double) fixed[at]*(1.0/FixedGridScale<T>::value));
}

template <typename T, bool FIXED>
__global__ void __launch_bounds__(256) k_spread(const PmeArgs a) {
    extern __shared__ __align__(16) unsigned char pmeSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PmeWarpTab<T, false>& t = reinterpret_cast<PmeWarpTab<T, false>*>(pmeSmem)[warp];
    unsigned todo = pmeBatchSetup<T, false>(a, t, lane, lane < a.batch ? ((int) blockIdx.x*((int) blockDim.x >> 5) + warp)*a.batch + lane : -1);
    if (todo == 0u) return;
    int ox[4], oy[4], oz[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int pt = min(lane + 32*i, 124), row = pt/5;
        oz[i] = pt - row*5; ox[i] = row/5; oy[i] = row - ox[i]*5;
    }
    const size_t G = (size_t) a.nx*a.ny*a.nz;
    while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        const int ix0 = t.ix0[b], iy0 = t.iy0[b], iz0 = t.iz0[b];
        const T qT = t.q[b];
        T* grid = (T*) a.grid + (size_t) t.subset[b]*G;
        unsigned long long* gridFixed = a.gridFixed + (size_t) t.subset[b]*G;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (lane + 32*i < 125) {
                int x = ix0 + ox[i]; x -= x >= a.nx ? a.nx : 0;
                int y = iy0 + oy[i]; y -= y >= a.ny ? a.ny : 0;
                int z = iz0 + oz[i]; z -= z >= a.nz ? a.nz : 0;
                const T value = qT*t.w[ox[i]][b]*t.w[5 + oy[i]][b]*t.w[10 + oz[i]][b];
                const size_t cell = ((size_t) x*a.ny + y)*a.nz + z;
                if (x < a.xLo || x >= a.xHi) continue;                        // another rank's plane
                if (FIXED) atomicAdd(gridFixed + cell, (unsigned long long) __double2ll_rn((double) value*FixedGridScale<T>::value));
                else atomicAdd(grid + cell, value);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// z transform, real -> half complex: two real lines (y, y+1) ride one complex FFT.
// ---------------------------------------------------------------------------------------------
struct FftArgs {
    int nS, nx, ny, nz, nzh;
    int ownLo, ownHi;            // x_conv: subsets whose mixed potential this rank produces
    int n;                       // length of the dimension this launch transforms
    unsigned long long factors;  // its radices, 4 bits each
    const void* tw;              // twiddles of this dimension: exp(-2 pi i k / n), precision T
    void* grid; void* gridC; const void* eterm;
    float* pot;                  // real-space potential grid read by the gather (float; double when potDouble)
    int potDouble;
    double* energy;
    int wantEnergy;
    LambdaTable lam;
};

template <typename T, int NQ>
__global__ void __launch_bounds__(256) k_fft_z_fwd(const FftArgs a) {
    typedef typename Cx<T>::type C;
    extern __shared__ double2 smRaw[];
    C* sm = (C*) smRaw;
    const int n = a.nz, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    C* tw = sm;
    C* line = sm + n + (size_t) warp*n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = ((const C*) a.tw)[k];
    __syncthreads();
    const int halfY = (a.ny + 1) >> 1;
    const int pair = blockIdx.x*8 + warp;
    if (pair >= a.nS*a.nx*halfY) return;
    const int sx = pair/halfY, y0 = 2*(pair - sx*halfY), y1 = y0 + 1;
    const T* r0 = (const T*) a.grid + ((size_t) sx*a.ny + y0)*n;
    const T* r1 = r0 + n;
    for (int z = lane; z < n; z += 32) line[z] = mk(r0[z], y1 < a.ny ? r1[z] : (T) 0);
    __syncwarp();
    warpFft<NQ>(line, a.n, a.factors, tw, lane);
    C* o0 = (C*) a.gridC + ((size_t) sx*a.ny + y0)*a.nzh;
    C* o1 = o0 + a.nzh;
    const T half = (T) 0.5;
    for (int k = lane; k < a.nzh; k += 32) {
        const C zk = line[k], zn = line[k == 0 ? 0 : n - k];
        o0[k] = mk(half*(zk.x + zn.x), half*(zk.y - zn.y));
        if (y1 < a.ny) o1[k] = mk(half*(zk.y + zn.y), -half*(zk.x - zn.x));
    }
}

// z transform, half complex -> real (inverse, unnormalised), two lines per complex FFT; writes the
// float potential grid the gather reads.
template <typename T, int NQ>
__global__ void __launch_bounds__(256) k_fft_z_inv(const FftArgs a) {
    typedef typename Cx<T>::type C;
    extern __shared__ double2 smRaw[];
    C* sm = (C*) smRaw;
    const int n = a.nz, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    C* tw = sm;
    C* line = sm + n + (size_t) warp*n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = ((const C*) a.tw)[k];
    __syncthreads();
    const int halfY = (a.ny + 1) >> 1;
    const int pair = blockIdx.x*8 + warp;
    if (pair >= a.nS*a.nx*halfY) return;
    const int sx = pair/halfY, y0 = 2*(pair - sx*halfY), y1 = y0 + 1;
    const C* i0 = (const C*) a.gridC + ((size_t) sx*a.ny + y0)*a.nzh;
    const C* i1 = i0 + a.nzh;
    for (int k = lane; k < a.nzh; k += 32) {
        const C A = i0[k], B = y1 < a.ny ? i1[k] : mk((T) 0, (T) 0);
        line[k] = mk(A.x - B.y, -(A.y + B.x));                          // conj(A + iB)
        if (k > 0 && 2*k < n) line[n - k] = mk(A.x + B.y, A.y - B.x);   // conj(conj(A) + i conj(B))
    }
    __syncwarp();
    warpFft<NQ>(line, a.n, a.factors, tw, lane);
    float* r0 = a.pot + ((size_t) sx*a.ny + y0)*n;
    float* r1 = r0 + n;
    double* d0 = (double*) a.pot + ((size_t) sx*a.ny + y0)*n;
    for (int z = lane; z < n; z += 32) {
        const C w = line[z];
        if (a.potDouble) { d0[z] = (double) w.x; if (y1 < a.ny) d0[n + z] = (double) -w.y; }
        else { r0[z] = (float) w.x; if (y1 < a.ny) r1[z] = (float) -w.y; }
    }
}

// y transform on the half spectrum, in place.  CTA = one (subset, x) plane x 16 consecutive kz.
template <typename T, int NQ, bool INVERSE>
__global__ void __launch_bounds__(512) k_fft_y(const FftArgs a) {
    typedef typename Cx<T>::type C;
    extern __shared__ double2 smRaw[];
    C* sm = (C*) smRaw;
    const int n = a.ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = n + 1;
    C* tw = sm;
    C* lines = sm + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = ((const C*) a.tw)[k];
    const int chunks = (a.nzh + 15) >> 4;
    const int sx = blockIdx.x/chunks, k0 = (blockIdx.x - sx*chunks)*16;
    C* base = (C*) a.gridC + (size_t) sx*n*a.nzh;
    for (int idx = threadIdx.x; idx < n*16; idx += blockDim.x) {
        const int l = idx & 15, y = idx >> 4;
        if (k0 + l < a.nzh) {
            C v = base[(size_t) y*a.nzh + k0 + l];
            if (INVERSE) v.y = -v.y;
            lines[l*stride + y] = v;
        }
    }
    __syncthreads();
    if (k0 + warp < a.nzh) warpFft<NQ>(lines + warp*stride, a.n, a.factors, tw, lane);
    __syncthreads();
    for (int idx = threadIdx.x; idx < n*16; idx += blockDim.x) {
        const int l = idx & 15, y = idx >> 4;
        if (k0 + l < a.nzh) {
            C v = lines[l*stride + y];
            if (INVERSE) v.y = -v.y;
            base[(size_t) y*a.nzh + k0 + l] = v;
        }
    }
}

// x transform + sliced convolution + x inverse.  CTA = one y x 8 consecutive kz, ALL subsets.
// Convolution and energies: pme_reciprocal_convolution, ReferencePME.cpp:400-496 -- eterm per k,
// E[slice(I,I)] += 1/2 eterm |S_I|^2, E[slice(I,J)] += eterm Re(S_I conj S_J) over the FULL grid (the
// half spectrum counts twice except on the kz = 0 and kz = nz/2 planes).  The reference then scales
// every subset grid by eterm and lets the gather mix subsets with lambda (:681-687); here the mix
// happens in k space.
template <typename T, int NQ, int NS>
__global__ void __launch_bounds__(256) k_fft_x_conv(const FftArgs a) {
    typedef typename Cx<T>::type C;
    extern __shared__ double2 smRaw[];
    C* sm = (C*) smRaw;
    __shared__ double shE[MAX_SLICES];
    const int n = a.nx, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = n + 1;
    C* tw = sm;
    C* lines = sm + n;                       // [s][l][x]
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = ((const C*) a.tw)[k];
    if (threadIdx.x < MAX_SLICES) shE[threadIdx.x] = 0.0;
    const int chunks = (a.nzh + 7) >> 3;
    const int y = blockIdx.x/chunks, k0 = (blockIdx.x - y*chunks)*8;
    const int nS = a.nS;
    C* gridC = (C*) a.gridC;
    for (int idx = threadIdx.x; idx < nS*n*8; idx += blockDim.x) {
        const int l = idx & 7, x = (idx >> 3) % n, s = idx/(8*n);
        if (k0 + l < a.nzh)
            lines[(s*8 + l)*stride + x] = gridC[(((size_t) s*n + x)*a.ny + y)*a.nzh + k0 + l];
    }
    __syncthreads();
    for (int L = warp; L < nS*8; L += 8)
        if (k0 + (L & 7) < a.nzh) warpFft<NQ>(lines + L*stride, a.n, a.factors, tw, lane);
    __syncthreads();
    double e[NS*(NS+1)/2];
#pragma unroll
    for (int s = 0; s < NS*(NS+1)/2; s++) e[s] = 0.0;
    for (int idx = threadIdx.x; idx < n*8; idx += blockDim.x) {
        const int l = idx & 7, x = idx >> 3, k = k0 + l;
        if (k >= a.nzh) continue;
        const T et = ((const T*) a.eterm)[((size_t) x*a.ny + y)*a.nzh + k];
        C S[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) S[s] = s < nS ? lines[(s*8 + l)*stride + x] : mk((T) 0, (T) 0);
        if (a.wantEnergy) {
            const T w = (k == 0 || 2*k == a.nz) ? (T) 1 : (T) 2;
#pragma unroll
            for (int sb = 0; sb < NS; sb++)
#pragma unroll
                for (int sa = 0; sa <= sb; sa++) {
                    if (sa < a.ownLo || sa >= a.ownHi) continue;     // a slice belongs to the owner of its lower subset
                    const T prod = S[sa].x*S[sb].x + S[sa].y*S[sb].y;
                    e[sb*(sb+1)/2 + sa] += (double) ((sa == sb ? (T) 0.5 : (T) 1)*w*et*prod);
                }
        }
#pragma unroll
        for (int si = 0; si < NS; si++) {
            if (si >= a.ownHi) break;
            if (si < a.ownLo) continue;
            T gx = 0, gy = 0;
#pragma unroll
            for (int sj = 0; sj < NS; sj++) {
                const T lam = (T) a.lam.c[triSlice(si, sj)];
                gx += lam*S[sj].x;
                gy += lam*S[sj].y;
            }
            lines[(si*8 + l)*stride + x] = mk(et*gx, -et*gy);      // conjugated for the inverse pass
        }
    }
    __syncthreads();
    const int nOwn = a.ownHi - a.ownLo;
    for (int L = a.ownLo*8 + warp; L < a.ownHi*8; L += 8)
        if (k0 + (L & 7) < a.nzh) warpFft<NQ>(lines + L*stride, a.n, a.factors, tw, lane);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nOwn*n*8; idx += blockDim.x) {
        const int l = idx & 7, x = (idx >> 3) % n, s = a.ownLo + idx/(8*n);
        if (k0 + l < a.nzh) {
            C v = lines[(s*8 + l)*stride + x];
            v.y = -v.y;
            gridC[(((size_t) s*n + x)*a.ny + y)*a.nzh + k0 + l] = v;
        }
    }
    if (a.wantEnergy) {
#pragma unroll
        for (int s = 0; s < NS*(NS+1)/2; s++) {
            const double v = warpSum(e[s]);
            if (lane == 0 && v != 0.0) atomicAdd(&shE[s], v);
        }
        __syncthreads();
        if (threadIdx.x < NS*(NS+1)/2 && shE[threadIdx.x] != 0.0)
            atomicAdd(a.energy + 2*threadIdx.x, shE[threadIdx.x]);       // Coulomb term of the slice
    }
}

// ---------------------------------------------------------------------------------------------
// Gather: one warp per atom, 125 points of the atom's own (lambda-mixed) potential grid.
// Reference: pme_grid_interpolate_force, ReferencePME.cpp:598-702.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_gather(const PmeArgs a) {
    extern __shared__ __align__(16) unsigned char pmeSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PmeWarpTab<T, true>& t = reinterpret_cast<PmeWarpTab<T, true>*>(pmeSmem)[warp];
    unsigned todo = pmeBatchSetup<T, true>(a, t, lane, lane < a.batch ? ((int) blockIdx.x*((int) blockDim.x >> 5) + warp)*a.batch + lane : -1);
    if (todo == 0u) return;
    int ox[4], oy[4], oz[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int pt = min(lane + 32*i, 124), row = pt/5;
        oz[i] = pt - row*5; ox[i] = row/5; oy[i] = row - ox[i]*5;
    }
    const size_t G = (size_t) a.nx*a.ny*a.nz;
    while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        const int ix0 = t.ix0[b], iy0 = t.iy0[b], iz0 = t.iz0[b];
        const T* pot = (const T*) a.pot + (size_t) t.subset[b]*G;
        T fx = 0, fy = 0, fz = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (lane + 32*i < 125) {
                int x = ix0 + ox[i]; x -= x >= a.nx ? a.nx : 0;
                int y = iy0 + oy[i]; y -= y >= a.ny ? a.ny : 0;
                int z = iz0 + oz[i]; z -= z >= a.nz ? a.nz : 0;
                // (slab sharding: the planes of other ranks contribute there; the force reduction adds the shares)
                const T g = (x >= a.xLo && x < a.xHi) ? pot[((size_t) x*a.ny + y)*a.nz + z] : (T) 0;
                const T tx = t.w[ox[i]][b], ty = t.w[5 + oy[i]][b], tz = t.w[10 + oz[i]][b];
                fx = fma(t.dw[ox[i]][b]*ty*tz, g, fx);
                fy = fma(tx*t.dw[5 + oy[i]][b]*tz, g, fy);
                fz = fma(tx*ty*t.dw[10 + oz[i]][b], g, fz);
            }
        }
        // three sums over the warp in six exchanges: the upper half-warp takes over y (and a zero), the lower keeps x and z;
        // then quarter-warps split those again, so that lanes 0-7 hold partial x, 8-15 z, 16-23 y
        {
            const bool up16 = lane & 16;
            const T k0 = up16 ? fy : fx, s0 = up16 ? fx : fy;
            const T k1 = up16 ? (T) 0 : fz, s1 = up16 ? fz : (T) 0;
            T u = k0 + __shfl_xor_sync(FULL_MASK, s0, 16);          // lower: x total of the pair, upper: y
            T v = k1 + __shfl_xor_sync(FULL_MASK, s1, 16);          // lower: z, upper: nothing
            const bool up8 = lane & 8;
            const T keep = up8 ? v : u, send = up8 ? u : v;
            T r = keep + __shfl_xor_sync(FULL_MASK, send, 8);       // lanes 0-7: x, 8-15: z, 16-23: y, 24-31: nothing
            r += __shfl_xor_sync(FULL_MASK, r, 4);
            r += __shfl_xor_sync(FULL_MASK, r, 2);
            r += __shfl_xor_sync(FULL_MASK, r, 1);
            fx = __shfl_sync(FULL_MASK, r, 0); fz = __shfl_sync(FULL_MASK, r, 8); fy = __shfl_sync(FULL_MASK, r, 16);
        }
        if (lane < 3) {
            // F = -q (fx nx r00, fx nx r10 + fy ny r11, fx nx r20 + fy ny r21 + fz nz r22), ReferencePME.cpp:698-700
            // (the off-diagonal reciprocal-vector terms are zero for a rectangular box)
            const T q = t.q[b];
            const T f = sizeof(T) == 8
                ? (lane == 0 ? fx*(T) a.fscaleD[0] : (lane == 1 ? fx*(T) a.fr10D + fy*(T) a.fscaleD[1] : fx*(T) a.fr20D + fy*(T) a.fr21D + fz*(T) a.fscaleD[2]))
                : (lane == 0 ? fx*(T) a.fscale[0] : (lane == 1 ? fx*(T) a.fr10 + fy*(T) a.fscale[1] : fx*(T) a.fr20 + fy*(T) a.fr21 + fz*(T) a.fscale[2]));
            if (f != (T) 0) atomicAdd(a.force + (size_t) lane*a.Npad + t.atom[b], toFixed(-q*f));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Influence function eterm(k) = exp(-pi^2 m^2 / alpha^2) / (pi V m^2 Bx By Bz), ReferencePME.cpp:426-471
// (without ONE_4PI_EPS0, which the charges carry).  Recomputed only when the box changes.
// ---------------------------------------------------------------------------------------------
// The reference transforms complex-to-complex and applies eterm(k) to every element of the full grid, with the
// frequency of index k taken as k or k - n (:436-466).  On the Nyquist planes of a TRICLINIC box that convention is not
// inversion symmetric (index n/2 is its own partner but maps to frequency -n/2 only), so eterm(k) != eterm(partner of
// k) there and the convolved grid is not Hermitian; the reference then keeps the real part of the inverse transform,
// which is the transform of the Hermitian part.  The half-spectrum pipeline here gets the same result by using
// eterm_eff(k) = (eterm(k) + eterm(partner(k)))/2, partner = ((nx-kx)%nx, (ny-ky)%ny, (nz-kz)%nz) -- identical to
// eterm(k) everywhere for a rectangular box and off the Nyquist planes.
template <bool DISPERSION>
__device__ __forceinline__ double influence(int kx, int ky, int kz, int nx, int ny, int nz, double3 invBox, double3 offRecip,
                                            double volume, double alpha, const double* __restrict__ moduli) {
    if (!DISPERSION && kx == 0 && ky == 0 && kz == 0) return 0.0;            // :457-460
    // m = fx r0 + fy r1 + fz r2 with the reciprocal vectors of ReferencePME.cpp:186-194 (offRecip = r10, r20, r21;
    // zero for a rectangular box)
    const double fx = (kx < (nx+1)/2 ? kx : kx - nx), fy = (ky < (ny+1)/2 ? ky : ky - ny), fz = (kz < (nz+1)/2 ? kz : kz - nz);
    const double mx = fx*invBox.x;
    const double my = fx*offRecip.x + fy*invBox.y;
    const double mz = fx*offRecip.y + fy*offRecip.z + fz*invBox.z;
    const double m2 = mx*mx + my*my + mz*mz;
    const double bsp = moduli[kx]*moduli[nx + ky]*moduli[nx + ny + kz];
    if (!DISPERSION) return exp(-kPi*kPi*m2/(alpha*alpha))/(m2*kPi*volume*bsp);
    // dispersion (LJPME), dpme_reciprocal_convolution, ReferencePME.cpp:520-570: eterm = (2 pi^3 sqrt(pi) erfc(b) m^3 +
    // exp(-b^2) (alpha^3 - 2 alpha pi^2 m^2)) * (-2 pi sqrt(pi) / (6 V Bx By Bz)), b = pi m / alpha; the m = 0 term is kept (:551)
    const double sqrtPi = 1.7724538509055160273;
    const double denom = (-2*kPi*sqrtPi/(6.0*volume))/bsp;
    const double m = sqrt(m2), b = (kPi/alpha)*m;
    const double fac1 = 2.0*kPi*kPi*kPi*sqrtPi, fac2 = alpha*alpha*alpha, fac3 = -2.0*alpha*kPi*kPi;
    return (fac1*erfc(b)*m*m2 + exp(-b*b)*(fac2 + fac3*m2))*denom;
}

template <typename T, bool DISPERSION>
__global__ void k_eterm(int nx, int ny, int nz, int nzh, double3 invBox, double3 offRecip, double volume, double alpha,
                        const double* __restrict__ moduli, T* __restrict__ eterm) {
    const size_t idx = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= (size_t) nx*ny*nzh) return;
    const int kz = (int) (idx % nzh), ky = (int) ((idx/nzh) % ny), kx = (int) (idx/((size_t) nzh*ny));
    const double here = influence<DISPERSION>(kx, ky, kz, nx, ny, nz, invBox, offRecip, volume, alpha, moduli);
    const double partner = influence<DISPERSION>((nx - kx) % nx, (ny - ky) % ny, (nz - kz) % nz, nx, ny, nz, invBox, offRecip,
                                                 volume, alpha, moduli);
    eterm[idx] = (T) (0.5*(here + partner));
}

int prepareEterm(Context& c) {
    const CellGeom& g = c.geom;
    if (c.etermBox[0] == g.box[0] && c.etermBox[1] == g.box[1] && c.etermBox[2] == g.box[2] &&
        c.etermBox[3] == g.tilt[0] && c.etermBox[4] == g.tilt[1] && c.etermBox[5] == g.tilt[2]) return NBS_OK;
    const int nx = c.grid[0], ny = c.grid[1], nz = c.grid[2], nzh = nz/2 + 1;
    const size_t total = (size_t) nx*ny*nzh;
    NBS_CUDA_CHECK(c.dEterm.ensure(total));
    NBS_CUDA_CHECK(c.dEtermD.ensure(total));
    const double3 inv = make_double3(g.invBox[0], g.invBox[1], g.invBox[2]);
    // r10 = -bx/(ax by), r20 = (bx cy - by cx)/(ax by cz), r21 = -cy/(by cz)
    const double3 off = make_double3(-g.tilt[0]*g.invBox[0]*g.invBox[1],
                                     (g.tilt[0]*g.tilt[2] - g.box[1]*g.tilt[1])*g.invBox[0]*g.invBox[1]*g.invBox[2],
                                     -g.tilt[2]*g.invBox[1]*g.invBox[2]);
    const double volume = g.box[0]*g.box[1]*g.box[2];
    if (c.dispersionPass) {
        k_eterm<float, true><<<(unsigned) ((total + 255)/256), 256, 0, c.stream>>>(nx, ny, nz, nzh, inv, off, volume, c.alpha, c.dModuli.d, c.dEterm.d);
        k_eterm<double, true><<<(unsigned) ((total + 255)/256), 256, 0, c.stream>>>(nx, ny, nz, nzh, inv, off, volume, c.alpha, c.dModuli.d, c.dEtermD.d);
    }
    else {
        k_eterm<float, false><<<(unsigned) ((total + 255)/256), 256, 0, c.stream>>>(nx, ny, nz, nzh, inv, off, volume, c.alpha, c.dModuli.d, c.dEterm.d);
        k_eterm<double, false><<<(unsigned) ((total + 255)/256), 256, 0, c.stream>>>(nx, ny, nz, nzh, inv, off, volume, c.alpha, c.dModuli.d, c.dEtermD.d);
    }
    c.launches += 2;
    for (int k = 0; k < 3; k++) { c.etermBox[k] = g.box[k]; c.etermBox[3+k] = g.tilt[k]; }
    return NBS_OK;
}

// The chain is split where a multi-rank evaluation exchanges spectra: `half` 0 = z and y forward
// transforms of this rank's own subsets, `half` 1 = fused x pass (all subsets in, own subsets out),
// inverse y and z of the own subsets.
template <typename T, int NQ>
static void launchFftChain(Context& c, FftArgs f, const FftPlan& px, const FftPlan& py, const FftPlan& pz,
                           size_t smX, size_t smY, size_t smZ, int half) {
    typedef typename Cx<T>::type C;
    cudaStream_t st = c.stream;
    const int nx = c.grid[0], ny = c.grid[1], nz = c.grid[2], nzh = nz/2 + 1;
    const C* tw = (const C*) (sizeof(T) == 8 ? (const void*) c.dTwiddleD.d : (const void*) c.dTwiddle.d);
    const int nOwn = c.ownHi - c.ownLo;
    const int pairs = nOwn*nx*((ny + 1)/2);
    const int yCtas = nOwn*nx*((nzh + 15)/16), xCtas = ny*((nzh + 7)/8);
    static bool attr[64] = {false};
    if (!attr[c.device & 63]) {
        const int big = 200*1024;
        cudaFuncSetAttribute(k_fft_z_fwd<T, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_z_inv<T, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_y<T, NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_y<T, NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv<T, NQ, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv<T, NQ, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv<T, NQ, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv<T, NQ, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv<T, NQ, MAX_SUBSETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        attr[c.device & 63] = true;
    }
    // the per-subset kernels see only the own slabs: offset the base pointers, nS = own count
    FftArgs own = f;
    own.nS = nOwn;
    own.grid = (T*) f.grid + (size_t) c.ownLo*nx*ny*nz;
    own.gridC = (C*) f.gridC + (size_t) c.ownLo*nx*ny*nzh;
    own.pot = f.potDouble ? (float*) ((double*) f.pot + (size_t) c.ownLo*nx*ny*nz) : f.pot + (size_t) c.ownLo*nx*ny*nz;
    if (half == 0) {
        own.n = pz.n; own.factors = pz.packed; own.tw = tw + nx + ny;
        k_fft_z_fwd<T, NQ><<<(pairs + 7)/8, 256, smZ, st>>>(own);
        own.n = py.n; own.factors = py.packed; own.tw = tw + nx;
        k_fft_y<T, NQ, false><<<yCtas, 512, smY, st>>>(own);
        c.launches += 2;
        return;
    }
    f.n = px.n; f.factors = px.packed; f.tw = tw;
    switch (c.nS) {
        case 1: k_fft_x_conv<T, NQ, 1><<<xCtas, 256, smX, st>>>(f); break;
        case 2: k_fft_x_conv<T, NQ, 2><<<xCtas, 256, smX, st>>>(f); break;
        case 3: k_fft_x_conv<T, NQ, 3><<<xCtas, 256, smX, st>>>(f); break;
        case 4: k_fft_x_conv<T, NQ, 4><<<xCtas, 256, smX, st>>>(f); break;
        default: k_fft_x_conv<T, NQ, MAX_SUBSETS><<<xCtas, 256, smX, st>>>(f); break;
    }
    own.n = py.n; own.factors = py.packed; own.tw = tw + nx;
    k_fft_y<T, NQ, true><<<yCtas, 512, smY, st>>>(own);
    own.n = pz.n; own.factors = pz.packed; own.tw = tw + nx + ny;
    k_fft_z_inv<T, NQ><<<(pairs + 7)/8, 256, smZ, st>>>(own);
    c.launches += 3;
}

template <typename T>
static int launchPmeT(Context& c, bool wantEnergy, int half) {
    const int nx = c.grid[0], ny = c.grid[1], nz = c.grid[2], nzh = nz/2 + 1;
    const size_t G = (size_t) nx*ny*nz;
    cudaStream_t st = c.stream;
    if (c.ownHi <= c.ownLo) return NBS_OK;            // this rank owns no subset grid
    FftPlan px, py, pz;
    if (!makePlan(nx, px) || !makePlan(ny, py) || !makePlan(nz, pz)) {
        setError("PME grid dimensions must be <= 512 and factor into 2, 3, 5, 7, 11, 13");
        return NBS_ERR_UNSUPPORTED;
    }
    PmeArgs p;
    p.N = c.N; p.Npad = c.Npad; p.nS = c.nS; p.nx = nx; p.ny = ny; p.nz = nz; p.nzh = nzh;
    p.ownLo = c.ownLo; p.ownHi = c.ownHi;
    p.xLo = c.slabMode ? c.xLo : 0; p.xHi = c.slabMode ? c.xHi : nx;
    p.binStart = c.dBinStart.d;
    p.rangeBin[0] = -1; p.rangeBin[1] = p.rangeBin[2] = p.rangeBin[3] = 0;
    if (c.slabMode && c.nRanks > 1 && !c.pmeUnsorted && !c.dispersionPass && !c.geom.triclinic) {
        // cell columns (x-major in the sorted order) whose atoms can touch the planes [xLo, xHi): an order-5 spline
        // starting at plane xLo - 4 .. xHi - 1, plus what an atom may have moved since the sort that placed it
        const CellGeom& g = c.geom;
        const double margin = (0.5*c.skin + 2.0e-3)*g.invBox[0];
        const double uLo = (double) (p.xLo - (PME_ORDER - 1))/nx - margin, uHi = (double) p.xHi/nx + margin;
        const int colLo = (int) std::floor(uLo*g.ncx), colHi = (int) std::floor(uHi*g.ncx);       // inclusive
        if (colHi - colLo + 1 < g.ncx) {
            int A0 = colLo, A1 = colHi, B0 = 0, B1 = -1;
            if (colLo < 0) { A0 = 0; B0 = colLo + g.ncx; B1 = g.ncx - 1; }
            else if (colHi >= g.ncx) { A1 = g.ncx - 1; B0 = 0; B1 = colHi - g.ncx; }
            const int perCol = g.ncy*g.nzb;
            p.rangeBin[0] = A0*perCol; p.rangeBin[1] = (A1 + 1)*perCol;
            p.rangeBin[2] = B0*perCol; p.rangeBin[3] = (B1 + 1)*perCol;
        }
    }
    // batches of up to 32 atoms per warp: fewer on small systems, so that there still are a few thousand warps (a batch
    // is worked off one atom after the other -- at 32 atoms per warp a DHFR-size system would be 700 warps of latency)
    int batch = 32;
    while (batch > 4 && (long long) c.N < (long long) batch*c.numSMs*64) batch >>= 1;       // measured: 4 at C3, 8 at C4, 32 at C5
    static const int batchEnv = getenv("NBS_PME_BATCH") ? atoi(getenv("NBS_PME_BATCH")) : 0;      // tuning experiments
    if (batchEnv == 4 || batchEnv == 8 || batchEnv == 16 || batchEnv == 32) batch = batchEnv;
    p.batch = batch;
    const int pmeThreads = 256;
    const size_t pmeSmem = sizeof(PmeWarpTab<T, false>)*(pmeThreads/32);
    const int atomsPerCta = (pmeThreads/32)*batch;
    int atomCtas = (c.N + atomsPerCta - 1)/atomsPerCta;
    p.useHostRange = 0;
    for (int k = 0; k < 4; k++) p.hostRange[k] = 0;
    if (p.rangeBin[0] >= 0) {
        if (c.reuseNow && c.slabRangeValid) {
            // the sort order is the one whose ranges came back with the counters: launch exactly those atoms
            p.useHostRange = 1;
            for (int k = 0; k < 4; k++) p.hostRange[k] = c.slabRange[k];
            atomCtas = std::max(1, (c.slabRange[1] - c.slabRange[0] + c.slabRange[3] - c.slabRange[2] + atomsPerCta - 1)/atomsPerCta);
        }
        else if (half == 0) {
            // a fresh sort: every warp checks the ranges on the device; the host gets them with this evaluation's counters
            k_slab_ranges<<<1, 32, 0, st>>>(c.dBinStart.d, p.rangeBin[0], p.rangeBin[1], p.rangeBin[2], p.rangeBin[3], c.dCounters.d + 8);
            c.launches++;
        }
    }
    p.posq = c.dPosq.d; p.par = c.dPar.d; p.grid = c.dGrid.d; p.pot = c.dPot.d; p.gridFixed = nullptr;
    p.unsorted = c.pmeUnsorted ? 1 : 0;
    p.fix = c.dFix.d; p.chargeF = c.dChargeF.d; p.subsetOf = c.dSubset.d;
    p.q64 = c.dQ64.d; p.chargeD = c.dCharge.d; p.sqrtK = sqrt(kOne4PiEps0);
    if (c.dispersionPass) {          // LJPME dispersion chain: the C6 coefficients are the "charges", no unit factor
        p.unsorted = 1; p.chargeF = c.dC6F.d; p.chargeD = c.dC6D.d; p.sqrtK = 1.0;
    }
    p.force = c.pmeUnsorted ? c.dForce.d + 3*(size_t) c.Npad : c.dForce.d;
    for (int k = 0; k < 3; k++) { p.fscale[k] = (float) (c.grid[k]*c.geom.invBox[k]); p.fscaleD[k] = c.grid[k]*c.geom.invBox[k]; }
    {
        const CellGeom& g = c.geom;
        p.triclinic = g.triclinic ? 1 : 0;
        p.beta = g.tilt[0]*g.invBox[0];
        p.gammaY = g.tilt[2]*g.invBox[1];
        p.delta = (g.tilt[0]*g.tilt[2] - g.box[1]*g.tilt[1])*g.invBox[0]*g.invBox[1];
        p.fr10D = -c.grid[0]*g.tilt[0]*g.invBox[0]*g.invBox[1];
        p.fr20D = c.grid[0]*(g.tilt[0]*g.tilt[2] - g.box[1]*g.tilt[1])*g.invBox[0]*g.invBox[1]*g.invBox[2];
        p.fr21D = -c.grid[1]*g.tilt[2]*g.invBox[1]*g.invBox[2];
        p.fr10 = (float) p.fr10D; p.fr20 = (float) p.fr20D; p.fr21 = (float) p.fr21D;
    }
    if (half == 0) {
        int status = prepareEterm(c);
        if (status != NBS_OK) return status;
        // the cells this rank spreads into: the grids of its subsets, or (slab sharding) its planes of every grid
        const size_t planeCells = (size_t) ny*nz, first = (size_t) c.ownLo*G + (size_t) p.xLo*planeCells;
        const size_t cellsPerGrid = c.slabMode ? (size_t) (p.xHi - p.xLo)*planeCells : G*(c.ownHi - c.ownLo);
        const int gridsToClear = c.slabMode ? c.nS : 1;
        if (c.flags & NBS_FLAG_DETERMINISTIC) {
            NBS_CUDA_CHECK(c.dGridFixed.ensure(G*c.nS));
            p.gridFixed = c.dGridFixed.d;
            NBS_CUDA_CHECK(cudaMemset2DAsync(c.dGridFixed.d + first, sizeof(unsigned long long)*G, 0, sizeof(unsigned long long)*cellsPerGrid, gridsToClear, st));
            k_spread<T, true><<<atomCtas, pmeThreads, pmeSmem, st>>>(p);
            k_fixed_to_real<T><<<dim3((unsigned) ((cellsPerGrid + 255)/256), gridsToClear), 256, 0, st>>>(cellsPerGrid, G, (const long long*) c.dGridFixed.d + first,
                                                                                                           (T*) c.dGrid.d + first);
            c.launches++;
        }
        else {
            NBS_CUDA_CHECK(cudaMemset2DAsync((T*) c.dGrid.d + first, sizeof(T)*G, 0, sizeof(T)*cellsPerGrid, gridsToClear, st));
            k_spread<T, false><<<atomCtas, pmeThreads, pmeSmem, st>>>(p);
        }
        c.launches++;
        timerMark(c, "spread");
    }

    FftArgs f;
    f.nS = c.nS; f.nx = nx; f.ny = ny; f.nz = nz; f.nzh = nzh;
    f.ownLo = c.ownLo; f.ownHi = c.ownHi;
    // the convolution kernels add slice s's energy to energy[2 s]: Coulomb term, or (dispersion chain) the vdW term
    double* const energyBase = c.dEnergy.d + (c.dispersionPass ? 1 : 0);
    const bool doubleForces = sizeof(T) == 8 && (c.flags & NBS_FLAG_DOUBLE);
    f.grid = c.dGrid.d; f.gridC = c.dGridC.d; f.pot = c.dPot.d; f.potDouble = doubleForces ? 1 : 0; f.energy = energyBase;
    f.eterm = sizeof(T) == 8 ? (const void*) c.dEtermD.d : (const void*) c.dEterm.d;
    f.wantEnergy = wantEnergy ? 1 : 0;
    for (int s = 0; s < MAX_SLICES; s++) {
        // the kernels mix the subset potentials with lam.c: lambda_Coulomb, or (dispersion chain) lambda_vdW (:868)
        f.lam.c[s] = s < c.nSl ? (float) c.lambdas[2*s + (c.dispersionPass ? 1 : 0)] : 1.f;
        f.lam.v[s] = s < c.nSl ? (float) c.lambdas[2*s+1] : 1.f;
    }
    const size_t cs = 2*sizeof(T);
    const size_t smZ = cs*(size_t) nz*9;
    const size_t smY = cs*((size_t) ny + 16*((size_t) ny + 1));
    const size_t smX = cs*((size_t) nx + (size_t) c.nS*8*((size_t) nx + 1));
    // plane-fused transforms (k_fft.cu) whenever a plane fits in shared memory ...
    {
        PlaneFftPlan plan;
        plan.factors[0] = px.packed; plan.factors[1] = py.packed; plan.factors[2] = pz.packed;
        PlaneFftArgs pa;
        pa.nS = c.nS; pa.nx = nx; pa.ny = ny; pa.nz = nz; pa.nzh = nzh;
        pa.ownLo = c.ownLo; pa.ownHi = c.ownHi;
        pa.xLo = c.slabMode ? c.xLo : 0; pa.nxOwn = c.slabMode ? c.xHi - c.xLo : nx;
        pa.yLo = c.slabMode ? c.yLo : 0; pa.nyOwn = c.slabMode ? c.yHi - c.yLo : ny;
        pa.nRanks = c.slabMode ? c.nRanks : 1;
        for (int r = 0; r < NBS_MAX_RANKS; r++) pa.peerSpectra[r] = c.slabMode && r < c.nRanks ? c.peerSpectra[r] : (void*) c.dGridC.d;
        pa.rowStride = 0; pa.chunk = 0;
        pa.grid = c.dGrid.d; pa.gridC = c.dGridC.d; pa.eterm = f.eterm; pa.pot = c.dPot.d; pa.potDouble = f.potDouble;
        pa.energy = energyBase; pa.wantEnergy = f.wantEnergy; pa.lam = f.lam;
        const int planeStatus = (c.flags & NBS_FLAG_LINE_FFT) ? NBS_RETRY : launchPlaneFft<T>(c, plan, pa, half);
        if (planeStatus < 0) return planeStatus;
        if (planeStatus == NBS_OK) {
            timerMark(c, half == 0 ? "fft_fwd" : (half == 3 ? "fft_inv" : "fft_conv_inv"));
            if (half == 1 || half == 3) {
                {
                    if (doubleForces) {
                        // double-precision gather: half the warps per CTA (its tables are twice the size)
                        PmeArgs pd = p;
                        k_gather<double><<<2*atomCtas, pmeThreads/2, sizeof(PmeWarpTab<double, true>)*(pmeThreads/64), st>>>(pd);
                    }
                    else k_gather<float><<<atomCtas, pmeThreads, sizeof(PmeWarpTab<float, true>)*(pmeThreads/32), st>>>(p);
                }
                c.launches++;
                timerMark(c, "gather");
            }
            return NBS_OK;
        }
    }
    // ... otherwise one line per warp, five kernels
    if (c.slabMode) {
        setError("slab sharding needs the plane-FFT path (the x lines of all subsets must fit in shared memory; NBS_FLAG_LINE_FFT is not supported)");
        return NBS_ERR_UNSUPPORTED;
    }
    if (smX > 200*1024 || smY > 200*1024 || smZ > 200*1024) {
        setError("PME grid too large for the shared-memory FFT");
        return NBS_ERR_UNSUPPORTED;
    }
    const int nq = std::max(px.nq, std::max(py.nq, pz.nq));
    if (nq <= 1) launchFftChain<T, 1>(c, f, px, py, pz, smX, smY, smZ, half);
    else if (nq <= 2) launchFftChain<T, 2>(c, f, px, py, pz, smX, smY, smZ, half);
    else if (nq <= 4) launchFftChain<T, 4>(c, f, px, py, pz, smX, smY, smZ, half);
    else launchFftChain<T, 8>(c, f, px, py, pz, smX, smY, smZ, half);
    timerMark(c, half == 0 ? "fft_fwd" : "fft_conv_inv");
    if (half == 1) {
        {
                    if (doubleForces) {
                        // double-precision gather: half the warps per CTA (its tables are twice the size)
                        PmeArgs pd = p;
                        k_gather<double><<<2*atomCtas, pmeThreads/2, sizeof(PmeWarpTab<double, true>)*(pmeThreads/64), st>>>(pd);
                    }
                    else k_gather<float><<<atomCtas, pmeThreads, sizeof(PmeWarpTab<float, true>)*(pmeThreads/32), st>>>(p);
                }
        c.launches++;
        timerMark(c, "gather");
    }
    return NBS_OK;
}

// Energies requested -> double-precision grids and transforms; forces only -> single precision.
// half 0: spread + forward z/y transforms of the own subsets; half 1: x pass + convolution, inverse, gather
// (slab sharding: half 2 = x pass alone, half 3 = inverse transforms + gather).
int launchPme(Context& c, bool wantEnergy, int half) {
    const bool fp64 = (wantEnergy && !(c.flags & NBS_FLAG_FP32_ENERGY)) || (c.flags & NBS_FLAG_DOUBLE);
    return fp64 ? launchPmeT<double>(c, wantEnergy, half) : launchPmeT<float>(c, wantEnergy, half);
}

} // namespace nbs

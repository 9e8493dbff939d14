// k_sort.cu -- positions -> fixed-point fractional coordinates -> cell sort -> i-blocks.
//
// This stage has no counterpart in the reference's source: the plugin's platforms take their
// neighbour list from OpenMM ([external] computeNeighborListVoxelHash on the Reference platform,
// ReferenceNonbondedSlicingKernels.cpp:197; NonbondedUtilities on CUDA, CommonNonbondedSlicingKernels.cpp:721).
// It is rebuilt from scratch on every evaluation, like the Reference platform does.
//
// All kernels here are HBM/latency bound integer work: one coalesced pass over the atoms each.
#include "nbs_internal.h"
#include "nbs_device.cuh"

namespace nbs {

// ---------------------------------------------------------------------------------------------
// k_prep: one thread per input slot.  Wraps the position into the box (brick), converts to 32-bit
// fixed-point fractional coordinates and counts the atom into its (column, z-bin).
// ---------------------------------------------------------------------------------------------
// Wraps a position into the brick [0, ax) x [0, by) x [0, cz) with the lattice translations -- first along
// c = (cx, cy, cz), then b = (bx, by, 0), then a (tilt = (bx, cx, cy), all zero for a rectangular box) -- and
// converts it to 32-bit fixed-point fractions of the brick.
__device__ __forceinline__ void wrapToFixed(double x, double y, double z, const double3 invBox, const double3 origin,
                                            const double3 tilt, unsigned& ux, unsigned& uy, unsigned& uz) {
    x -= origin.x; y -= origin.y; z -= origin.z;
    double fz = z*invBox.z;
    const double kz = floor(fz);
    fz -= kz;
    x -= kz*tilt.y; y -= kz*tilt.z;
    double fy = y*invBox.y;
    const double ky = floor(fy);
    fy -= ky;
    x -= ky*tilt.x;
    double fx = x*invBox.x;
    fx -= floor(fx);
    ux = (unsigned) (__double2ull_rd(fx*4294967296.0) & 0xffffffffull);
    uy = (unsigned) (__double2ull_rd(fy*4294967296.0) & 0xffffffffull);
    uz = (unsigned) (__double2ull_rd(fz*4294967296.0) & 0xffffffffull);
}

__global__ void k_prep(int N, const double* __restrict__ pos64, const float4* __restrict__ pos32,
                       const double4* __restrict__ pos64w, const int* __restrict__ atomIndex, double3 invBox, double3 origin,
                       double3 tilt, int ncx, int ncy, int nzb, uint4* __restrict__ fix, int* __restrict__ binCount,
                       double* __restrict__ pos64out, double* __restrict__ energy, int* __restrict__ counters) {
    int slot = blockIdx.x*blockDim.x + threadIdx.x;
    // the slice-energy table and the counters are first touched by later kernels: zeroed here instead of by two
    // more memset nodes in front of the chain
    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < ENERGY_WORDS; k += blockDim.x) energy[k] = 0.0;
        if (threadIdx.x < 16) counters[threadIdx.x] = 0;
    }
    if (slot >= N) return;
    double x, y, z;
    if (pos64) { x = pos64[3*slot]; y = pos64[3*slot+1]; z = pos64[3*slot+2]; }
    else if (pos64w) { double4 p = pos64w[slot]; x = p.x; y = p.y; z = p.z; }
    else { float4 p = pos32[slot]; x = p.x; y = p.y; z = p.z; }
    int particle = atomIndex ? atomIndex[slot] : slot;
    if (pos64out) {          // particle-ordered, unwrapped, double: what the exception kernel reads
        pos64out[3*particle] = x; pos64out[3*particle+1] = y; pos64out[3*particle+2] = z;
    }
    unsigned ux, uy, uz;
    wrapToFixed(x, y, z, invBox, origin, tilt, ux, uy, uz);
    int bin = (__umulhi(ux, ncx)*ncy + __umulhi(uy, ncy))*nzb + __umulhi(uz, nzb);
    fix[particle] = make_uint4(ux, uy, uz, (unsigned) bin);
    atomicAdd(&binCount[bin], 1);
}

// ---------------------------------------------------------------------------------------------
// k_reprep: the per-evaluation front end while the neighbour list is RE-USED (the plugin's CUDA platform inherits
// the same policy from OpenMM: a padded list kept until atoms have moved half the padding,
// CommonNonbondedSlicingKernels.cpp:721 `useNeighborList`).  One thread per input slot: new fixed-point coordinates
// go straight to the atom's place in the sort order of the last build; nothing is sorted, no block or list changes.
//   * displacement since the build -> running maximum in counters[4] (float bits of nm^2; zeroed by the build's
//     k_prep).  The host redoes the evaluation with a fresh list if it exceeds half the skin.
//   * an atom that has left the brick since the build is wrapped like any other, and the lattice translation
//     (cross_x, cross_y, cross_z) that brings it back next to its build-time position is packed into par.z above
//     the subset (bits 4-6: cross_x in -2..2, 7-8: cross_y, 9-10: cross_z, two's complement): the pair kernel adds
//     it to the image shift of the list entry, whose image code refers to the build-time coordinates.
// ---------------------------------------------------------------------------------------------
__global__ void k_reprep(int N, const double* __restrict__ pos64, const float4* __restrict__ pos32,
                         const double4* __restrict__ pos64w, const int* __restrict__ atomIndex, double3 invBox, double3 origin,
                         double3 tilt, long long shiftB, long long shiftCx, long long shiftCy, float3 scale,
                         const int* __restrict__ origToSorted, const uint4* __restrict__ fixBuild,
                         const float* __restrict__ chargeF, const int* __restrict__ subset,
                         uint4* __restrict__ fix, uint4* __restrict__ posq, float4* __restrict__ par,
                         double* __restrict__ pos64out, double* __restrict__ energy, int* __restrict__ counters) {
    int slot = blockIdx.x*blockDim.x + threadIdx.x;
    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < ENERGY_WORDS; k += blockDim.x) energy[k] = 0.0;
        if (threadIdx.x == 0) { counters[1] = 0; counters[3] = 0; }      // overflow flag, pair-kernel work cursor
    }
    float d2 = 0.f;
    if (slot < N) {
        double x, y, z;
        if (pos64) { x = pos64[3*slot]; y = pos64[3*slot+1]; z = pos64[3*slot+2]; }
        else if (pos64w) { double4 p = pos64w[slot]; x = p.x; y = p.y; z = p.z; }
        else { float4 p = pos32[slot]; x = p.x; y = p.y; z = p.z; }
        const int particle = atomIndex ? atomIndex[slot] : slot;
        if (pos64out) { pos64out[3*particle] = x; pos64out[3*particle+1] = y; pos64out[3*particle+2] = z; }
        unsigned ux, uy, uz;
        wrapToFixed(x, y, z, invBox, origin, tilt, ux, uy, uz);
        const int s = origToSorted[particle];
        const uint4 fb = fixBuild[s];
        // lattice translation back to the build-time neighbourhood: z first (c tilts into x and y), then y, then x
        const long long half = 1ll << 31;
        long long dz = (long long) uz - (long long) fb.z;
        const int cz = dz > half ? -1 : (dz < -half ? 1 : 0);
        dz += (long long) cz << 32;
        long long dy = (long long) uy - (long long) fb.y + cz*shiftCy;
        int cy = 0;
        while (dy > half) { dy -= 1ll << 32; cy--; }
        while (dy < -half) { dy += 1ll << 32; cy++; }
        long long dx = (long long) ux - (long long) fb.x + cz*shiftCx + cy*shiftB;
        int cx = 0;
        while (dx > half) { dx -= 1ll << 32; cx--; }
        while (dx < -half) { dx += 1ll << 32; cx++; }
        const float ex = (float) dx*scale.x, ey = (float) dy*scale.y, ez = (float) dz*scale.z;
        d2 = ex*ex + ey*ey + ez*ez;
        if (cx < -2 || cx > 2 || cy < -1 || cy > 1) d2 = 1.0e30f;       // not representable: forces a fresh list
        fix[particle] = make_uint4(ux, uy, uz, 0u);
        posq[s] = make_uint4(ux, uy, uz, __float_as_uint(chargeF[particle]));
        reinterpret_cast<int*>(par)[4*s + 2] = subset[particle] | ((cx & 7) << 4) | ((cy & 3) << 7) | ((cz & 3) << 9);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    if ((threadIdx.x & 31) == 0 && d2 > 0.f) atomicMax(reinterpret_cast<unsigned*>(counters) + 4, __float_as_uint(d2));
}

// ---------------------------------------------------------------------------------------------
// Exclusive scan (int), out[n] = total.  Small inputs: one CTA walks the array with a carry.
// Large inputs: three passes with 4096-element tiles.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int blockExclusiveScan(int v, int* warpSums, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? warpSums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warpSums[lane] = winc - w;               // exclusive offsets of the warps
        if (lane == 31) warpSums[32] = winc;     // block total
    }
    __syncthreads();
    int result = warpSums[warp] + inc - v;
    total = warpSums[32];
    __syncthreads();
    return result;
}

__global__ void __launch_bounds__(1024) k_scan_single(const int* __restrict__ in, int* __restrict__ out, int n) {
    __shared__ int warpSums[33];
    // every thread owns a contiguous run of `per` elements: one block-wide scan in total
    const int per = (n + 1023) >> 10;
    const int begin = min(n, (int) threadIdx.x*per), end = min(n, begin + per);
    int sum = 0;
    for (int i = begin; i < end; i++) sum += in[i];
    int total;
    int ex = blockExclusiveScan(sum, warpSums, total);
    for (int i = begin; i < end; i++) { const int v = in[i]; out[i] = ex; ex += v; }
    if (threadIdx.x == 0) out[n] = total;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ tileSums) {
    __shared__ int warpSums[33];
    int base = blockIdx.x*4096 + threadIdx.x*4;
    int v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = base+k < n ? in[base+k] : 0; s += v[k]; }
    int total;
    int ex = blockExclusiveScan(s, warpSums, total);
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base+k < n) out[base+k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_add(int* __restrict__ out, int n, const int* __restrict__ tileOffsets) {
    int base = blockIdx.x*4096 + threadIdx.x*4;
    int off = tileOffsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++) if (base+k < n) out[base+k] += off;
    if (blockIdx.x == gridDim.x-1 && threadIdx.x == 0) out[n] = tileOffsets[gridDim.x];
}

static int scanExclusive(Context& c, const int* in, int* out, int n) {
    if (n <= 32768) {
        k_scan_single<<<1, 1024, 0, c.stream>>>(in, out, n);
        c.launches++;
        return NBS_OK;
    }
    int tiles = (n + 4095)/4096;
    NBS_CUDA_CHECK(c.dScanTmp.ensure(2*(size_t) tiles + 2));
    int* sums = c.dScanTmp.d;
    int* offsets = c.dScanTmp.d + tiles + 1;
    k_scan_tiles<<<tiles, 1024, 0, c.stream>>>(in, out, n, sums);
    k_scan_single<<<1, 1024, 0, c.stream>>>(sums, offsets, tiles);
    k_scan_add<<<tiles, 1024, 0, c.stream>>>(out, n, offsets);
    c.launches += 3;
    return NBS_OK;
}

// ---------------------------------------------------------------------------------------------
// k_scatter: slot of every particle inside its bin (order inside a bin is fixed up next).
// ---------------------------------------------------------------------------------------------
__global__ void k_scatter(int N, const uint4* __restrict__ fix, const int* __restrict__ binStart,
                          int* __restrict__ binCursor, int* __restrict__ sortedToOrig) {
    int p = blockIdx.x*blockDim.x + threadIdx.x;
    if (p >= N) return;
    int bin = (int) fix[p].w;
    int slot = binStart[bin] + atomicAdd(&binCursor[bin], 1);
    sortedToOrig[slot] = p;
}

// ---------------------------------------------------------------------------------------------
// k_place: one thread per scattered slot.  The final position of a particle inside its bin is its
// rank by particle index among the bin's members (so the sorted order -- and with it every
// floating-point sum downstream -- is reproducible whatever order the scatter's atomics ran in);
// the thread then writes the particle's sorted records.  Bins hold a handful of atoms (1/8 of a
// block height), so the rank is a short scan of the bin.
// ---------------------------------------------------------------------------------------------
__global__ void k_place(int N, const int* __restrict__ binStart, const int* __restrict__ scattered,
                        int* __restrict__ origToSorted, const uint4* __restrict__ fix,
                        const float* __restrict__ chargeF, const float2* __restrict__ sigEps,
                        const int* __restrict__ subset, const double* __restrict__ charge, double sqrtK,
                        uint4* __restrict__ posq, float4* __restrict__ par, double* __restrict__ q64,
                        uint4* __restrict__ fixBuild) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int p = scattered[s];
    const uint4 f = fix[p];
    const int begin = binStart[f.w], end = binStart[f.w + 1];
    int rank = 0;
    for (int k = begin; k < end; k++) rank += scattered[k] < p ? 1 : 0;
    const int slot = begin + rank;
    const float2 se = sigEps[p];
    posq[slot] = make_uint4(f.x, f.y, f.z, __float_as_uint(chargeF[p]));
    par[slot] = make_float4(se.x, se.y, __int_as_float(subset[p]), __int_as_float(p));
    q64[slot] = charge[p]*sqrtK;
    fixBuild[slot] = make_uint4(f.x, f.y, f.z, 0u);      // what a re-used list measures displacements from (k_reprep)
    origToSorted[p] = slot;
}

// ---------------------------------------------------------------------------------------------
// i-blocks: <= 32 consecutive sorted atoms of one column.
// ---------------------------------------------------------------------------------------------
__global__ void k_col_blocks(int nCols, int nzb, const int* __restrict__ binStart, int* __restrict__ colBlocks) {
    int col = blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= nCols) return;
    int cnt = binStart[(col+1)*nzb] - binStart[col*nzb];
    colBlocks[col] = (cnt + 31) >> 5;
}

// k_col_blocks + k_scan_single in one single-CTA kernel (the usual case: at most 32768 columns): blocks per column
// straight from the bin offsets, exclusive scan, total in out[nCols].
__global__ void __launch_bounds__(1024) k_col_block_scan(int nCols, int nzb, const int* __restrict__ binStart, int* __restrict__ out) {
    __shared__ int warpSums[33];
    const int per = (nCols + 1023) >> 10;
    const int begin = min(nCols, (int) threadIdx.x*per), end = min(nCols, begin + per);
    int sum = 0;
    for (int col = begin; col < end; col++) sum += (binStart[(col+1)*nzb] - binStart[col*nzb] + 31) >> 5;
    int total;
    int ex = blockExclusiveScan(sum, warpSums, total);
    for (int col = begin; col < end; col++) {
        out[col] = ex;
        ex += (binStart[(col+1)*nzb] - binStart[col*nzb] + 31) >> 5;
    }
    if (threadIdx.x == 0) out[nCols] = total;
}

__global__ void k_blocks(int nCols, int nzb, int maxBlocks, const int* __restrict__ binStart,
                         const int* __restrict__ colBlockStart, const uint4* __restrict__ posq,
                         int* __restrict__ blkFirst, int* __restrict__ blkCount, uint4* __restrict__ blkLo,
                         uint4* __restrict__ blkHi, int* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    int b = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int nBlocks = colBlockStart[nCols];
    if (b == 0 && lane == 0) counters[0] = nBlocks;
    if (b >= nBlocks || b >= maxBlocks) return;
    int lo = 0, hi = nCols;                       // last column with colBlockStart[col] <= b
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (colBlockStart[mid] <= b) lo = mid; else hi = mid;
    }
    int col = lo;
    int colBegin = binStart[col*nzb], colEnd = binStart[(col+1)*nzb];
    int first = colBegin + 32*(b - colBlockStart[col]);
    int count = min(32, colEnd - first);
    uint4 p = posq[first + min(lane, count-1)];
    unsigned xl = p.x, xh = p.x, yl = p.y, yh = p.y, zl = p.z, zh = p.z;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xl = min(xl, __shfl_xor_sync(0xffffffffu, xl, o)); xh = max(xh, __shfl_xor_sync(0xffffffffu, xh, o));
        yl = min(yl, __shfl_xor_sync(0xffffffffu, yl, o)); yh = max(yh, __shfl_xor_sync(0xffffffffu, yh, o));
        zl = min(zl, __shfl_xor_sync(0xffffffffu, zl, o)); zh = max(zh, __shfl_xor_sync(0xffffffffu, zh, o));
    }
    if (lane == 0) {
        // largest block extents (fixed-point units): decide whether this list may be re-used (nbs_api.cu phaseComplete)
        atomicMax(reinterpret_cast<unsigned*>(counters) + 5, xh - xl);
        atomicMax(reinterpret_cast<unsigned*>(counters) + 6, yh - yl);
        atomicMax(reinterpret_cast<unsigned*>(counters) + 7, zh - zl);
        blkFirst[b] = first;
        blkCount[b] = count;
        blkLo[b] = make_uint4(xl, yl, zl, (unsigned) col);
        blkHi[b] = make_uint4(xh, yh, zh, 0u);
    }
}

// For every sorted atom: the range of sorted indices its exclusion partners fall in (lets the list
// builder skip the per-atom exclusion walk for almost every (i-block, j) candidate).
__global__ void k_excl_range(int N, const float4* __restrict__ par, const int* __restrict__ exclStart,
                             const int* __restrict__ exclList, const int* __restrict__ origToSorted,
                             int2* __restrict__ exclRange) {
    int j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= N) return;
    int p = __float_as_int(par[j].w);
    int lo = 0x7fffffff, hi = -1;
    for (int k = exclStart[p]; k < exclStart[p+1]; k++) {
        int s = origToSorted[exclList[k]];
        lo = min(lo, s);
        hi = max(hi, s);
    }
    exclRange[j] = make_int2(lo, hi);
}

// Bounding box of the input positions (non-periodic methods build a virtual box around the system).
__global__ void __launch_bounds__(1024) k_bbox(int N, const double* __restrict__ pos64, const float4* __restrict__ pos32,
                                              const double4* __restrict__ pos64w, float* __restrict__ out) {
    __shared__ float red[6][32];
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float v[3];
        if (pos64) { v[0] = (float) pos64[3*i]; v[1] = (float) pos64[3*i+1]; v[2] = (float) pos64[3*i+2]; }
        else if (pos64w) { double4 p = pos64w[i]; v[0] = (float) p.x; v[1] = (float) p.y; v[2] = (float) p.z; }
        else { float4 p = pos32[i]; v[0] = p.x; v[1] = p.y; v[2] = p.z; }
#pragma unroll
        for (int d = 0; d < 3; d++) { lo[d] = fminf(lo[d], v[d]); hi[d] = fmaxf(hi[d], v[d]); }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 3; d++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if (lane == 0) { red[d][warp] = lo[d]; red[3+d][warp] = hi[d]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float r = red[threadIdx.x][0];
        for (int w = 1; w < (int) (blockDim.x >> 5); w++) r = threadIdx.x < 3 ? fminf(r, red[threadIdx.x][w]) : fmaxf(r, red[threadIdx.x][w]);
        out[threadIdx.x] = r;
    }
}

int launchBBox(Context& c, const PosInput& in, float out[6]) {
    NBS_CUDA_CHECK(c.dBBox.ensure(8));
    k_bbox<<<1, 1024, 0, c.stream>>>(c.N, in.format == NBS_POS_F64_XYZ ? (const double*) in.ptr : nullptr,
                                     in.format == NBS_POS_F32_XYZW ? (const float4*) in.ptr : nullptr,
                                     in.format == NBS_POS_F64_XYZW ? (const double4*) in.ptr : nullptr, c.dBBox.d);
    c.launches++;
    NBS_CUDA_CHECK(cudaMemcpyAsync(out, c.dBBox.d, 6*sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    NBS_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    return NBS_OK;
}

// Fixed-point conversion + bin histogram (everything downstream -- the rest of the sort AND the PME chain of small
// systems -- starts from its output).
int launchPrep(Context& c, const PosInput& in) {
    const CellGeom& g = c.geom;
    const int N = c.N;
    cudaStream_t st = c.stream;
    NBS_CUDA_CHECK(cudaMemsetAsync(c.dBinCount.d, 0, sizeof(int)*(g.nBins+1), st));
    NBS_CUDA_CHECK(cudaMemsetAsync(c.dBinCursor.d, 0, sizeof(int)*(g.nBins+1), st));
    const int T = 256;
    k_prep<<<(N+T-1)/T, T, 0, st>>>(N, in.format == NBS_POS_F64_XYZ ? (const double*) in.ptr : nullptr,
                                    in.format == NBS_POS_F32_XYZW ? (const float4*) in.ptr : nullptr,
                                    in.format == NBS_POS_F64_XYZW ? (const double4*) in.ptr : nullptr, in.atomIndex,
                                    make_double3(g.invBox[0], g.invBox[1], g.invBox[2]),
                                    make_double3(g.origin[0], g.origin[1], g.origin[2]),
                                    make_double3(g.tilt[0], g.tilt[1], g.tilt[2]), g.ncx, g.ncy, g.nzb,
                                    c.dFix.d, c.dBinCount.d, in.pos64out, c.dEnergy.d, c.dCounters.d);
    c.launches++;
    timerMark(c, "prep");
    return NBS_OK;
}

int launchReprep(Context& c, const PosInput& in) {
    const CellGeom& g = c.geom;
    const int N = c.N, T = 256;
    k_reprep<<<(N+T-1)/T, T, 0, c.stream>>>(N, in.format == NBS_POS_F64_XYZ ? (const double*) in.ptr : nullptr,
                                            in.format == NBS_POS_F32_XYZW ? (const float4*) in.ptr : nullptr,
                                            in.format == NBS_POS_F64_XYZW ? (const double4*) in.ptr : nullptr, in.atomIndex,
                                            make_double3(g.invBox[0], g.invBox[1], g.invBox[2]),
                                            make_double3(g.origin[0], g.origin[1], g.origin[2]),
                                            make_double3(g.tilt[0], g.tilt[1], g.tilt[2]), g.shiftB, g.shiftCx, g.shiftCy,
                                            make_float3(g.scale[0], g.scale[1], g.scale[2]),
                                            c.dOrigToSorted.d, c.dFixBuild.d, c.dChargeF.d, c.dSubset.d,
                                            c.dFix.d, c.dPosq.d, c.dPar.d, in.pos64out, c.dEnergy.d, c.dCounters.d);
    c.launches++;
    timerMark(c, "reprep");
    return NBS_OK;
}

// The rest of the cell sort: scan of the bin counts, counting sort, sorted records, i-blocks.
int launchSortRest(Context& c) {
    const CellGeom& g = c.geom;
    const int N = c.N;
    cudaStream_t st = c.stream;
    const int T = 256;
    int status = scanExclusive(c, c.dBinCount.d, c.dBinStart.d, g.nBins);
    if (status != NBS_OK) return status;
    k_scatter<<<(N+T-1)/T, T, 0, st>>>(N, c.dFix.d, c.dBinStart.d, c.dBinCursor.d, c.dSortedToOrig.d);
    k_place<<<(N+T-1)/T, T, 0, st>>>(N, c.dBinStart.d, c.dSortedToOrig.d, c.dOrigToSorted.d, c.dFix.d,
                                     c.dChargeF.d, c.dSigEps.d, c.dSubset.d, c.dCharge.d, sqrt(kOne4PiEps0),
                                     c.dPosq.d, c.dPar.d, c.dQ64.d, c.dFixBuild.d);
    c.launches += 2;
    // The per-atom exclusion ranges only need the sorted order: on the side stream they run beside the block
    // construction instead of in front of the list builder (which waits for them, launchBuildLists).
    c.exclRangeForked = false;
    if (c.sideFork) {
        cudaEventRecord(c.evPlaced, st);
        cudaStreamWaitEvent(c.auxStream, c.evPlaced, 0);
        c.stream = c.auxStream;
        status = launchExclRange(c);
        c.stream = st;
        if (status != NBS_OK) return status;
        cudaEventRecord(c.evExclDone, c.auxStream);
        c.exclRangeForked = true;
    }
    if (g.nCols <= 32768) {
        k_col_block_scan<<<1, 1024, 0, st>>>(g.nCols, g.nzb, c.dBinStart.d, c.dColBlockStart.d);
        c.launches++;
    }
    else {
        // dBinCount is reused for the per-column block counts
        k_col_blocks<<<(g.nCols+T-1)/T, T, 0, st>>>(g.nCols, g.nzb, c.dBinStart.d, c.dBinCount.d);
        c.launches++;
        status = scanExclusive(c, c.dBinCount.d, c.dColBlockStart.d, g.nCols);
        if (status != NBS_OK) return status;
    }
    k_blocks<<<(c.maxBlocks*32+T-1)/T, T, 0, st>>>(g.nCols, g.nzb, c.maxBlocks, c.dBinStart.d, c.dColBlockStart.d, c.dPosq.d,
                                                  c.dBlkFirst.d, c.dBlkCount.d, c.dBlkLo.d, c.dBlkHi.d, c.dCounters.d);
    c.launches++;
    timerMark(c, "sort");
    return NBS_OK;
}

int launchExclRange(Context& c) {
    const int T = 256;
    k_excl_range<<<(c.N+T-1)/T, T, 0, c.stream>>>(c.N, c.dPar.d, c.dExclStart.d, c.dExclList.d, c.dOrigToSorted.d, c.dExclRange.d);
    c.launches++;
    return NBS_OK;
}

} // namespace nbs

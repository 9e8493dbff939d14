// k_nlist.cu -- tile neighbour list with exclusion masks; built with cutoff + skin and re-used until an atom has moved
// half the skin (nbs_set_list_skin; skin 0 = rebuilt on every evaluation).
//
// Replaces what the reference gets from OpenMM: computeNeighborListVoxelHash + the exclusion sets
// on the Reference platform (ReferenceNonbondedSlicingKernels.cpp:101-106, 197) and
// NonbondedUtilities' block list / exclusion tiles on CUDA (CommonNonbondedSlicingKernels.cpp:721).
//
// One CTA per i-block (<= 32 sorted atoms of one column).  Its warps walk the neighbouring columns;
// for every column the interacting atoms are ONE contiguous range of the sorted arrays (z-bin range
// of the column, per periodic image), read in coalesced 32-atom chunks.  A candidate j survives if
// it is within cutoff of the i-block's bounding box.  Survivors are compacted with warp ballots into
// shared memory, then copied out as one dense list per i-block:
//   jlist : atoms that have no masked pair with this block             (fast tiles)
//   xlist : atoms of the block itself (pairs i >= j masked) and atoms that share an exclusion with
//           any atom of the block, each with the 32-bit mask of excluded i lanes
// Half list: a pair is owned by the block of the atom with the LOWER sorted index, whatever the
// periodic image, so columns before the block's own column are skipped outright.
//
// CLUSTER MASKS.  An i-block is 8 clusters of 4 consecutive sorted atoms.  Every list entry carries the 8-bit mask
// of clusters whose own bounding box it is within reach of (the pair kernel evaluates 4 i atoms x 8 j atoms per
// step and skips the clusters a group of 8 entries cannot reach: ~0.54 of the evaluated pairs are inside the
// cutoff instead of ~0.30 for whole 32 x 32 tiles).  jlist entries are stored SORTED BY MASK (stable counting sort
// inside the CTA: equal masks become neighbours, so the union over a group of 8 loses almost nothing) together
// with one mask byte per group of 8 entries (gmask; the union of the group's masks).
#include "nbs_internal.h"
#include "nbs_device.cuh"
#include <algorithm>

namespace nbs {

constexpr int NL_PREFETCH = 4;  // 32-atom chunks of a column range whose loads are in flight together
// Per-warp staging capacities (entries) follow the list capacities (3/8 of them: 768 / 96 at the initial 2048 / 256),
// so a system whose blocks overflow the staging -- tiny boxes seen through many periodic images -- gets more of
// both when the evaluation is repeated with doubled capacities; bounded by the shared memory of an SM.
// (worst case, capJ = 65536: 147 KB of staged entries + 31 KB of cluster-mask bytes + 9 KB of group masks + 9 KB static)
static int stageJ(int capJ) { return std::max(32, std::min(3072, capJ*3/8)); }
static int stageX(int capX) { return std::max(32, std::min(768, capX*3/8)); }

struct BuildArgs {
    int N, maxBlocks, capJ, capX;
    int stageJ, stageX;        // per-warp staging capacities
    int blockPeriod, blockOffset, blockWidth;   // this rank's share of the i-blocks
    int ncx, ncy, nzb;
    int periodic;              // 0: NoCutoff / CutoffNonPeriodic -- no images
    float colWx, colWy, binH;
    float Lx, Ly, Lz;
    float bx, cx, cy;          // triclinic tilt (nm): b = (bx, Ly, 0), c = (cx, cy, Lz)
    long long shiftB, shiftCx, shiftCy;   // the same in fixed-point units of their axis
    float sx, sy, sz;          // nm per fixed-point unit
    float reach;               // cutoff + margin
    const int* counters;
    const int* blkFirst; const int* blkCount; const uint4* blkLo; const uint4* blkHi;
    const int* binStart;
    const uint4* posq; const float4* par;
    const int2* exclRange; const int* exclStart; const int* exclList; const int* origToSorted;
    int* jlist; int* jcount; int* xlist; unsigned* xmask; int* xcount;
    int* overflow;             // counters + 1
    int* itemCount;            // counters + 2
    unsigned* gmJ; unsigned* gmX;   // [nBlocks][capJ/32], [nBlocks][capX/32]: cluster masks of 4 groups (one tile) per word
    int4* items;               // work items of the pair kernel: (local block, first tile, first atom, atom count)
    int chunkTiles, maxItems;
    double* overflowFlag;      // energy[2*MAX_SLICES]: the same flag as a double, so that it all-reduces
};

__global__ void __launch_bounds__(BUILD_WARPS*32, 4) k_build_lists(BuildArgs a) {
    const int lb = blockIdx.x;                  // rank-local block index (lists are stored there)
    const int b = localToGlobalBlock(lb, a.blockPeriod, a.blockOffset, a.blockWidth);
    if (b >= a.counters[0]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // dynamic shared memory: [BUILD_WARPS][stageJ] j entries, [BUILD_WARPS][stageX] x entries, x exclusion masks,
    // group-mask words of the two lists, then the cluster-mask bytes of the staged entries
    extern __shared__ int stage[];
    int* const jbuf = stage + warp*a.stageJ;
    int* const xbuf = stage + BUILD_WARPS*a.stageJ + warp*a.stageX;
    unsigned* const xmbuf = (unsigned*) (stage + BUILD_WARPS*(a.stageJ + a.stageX)) + warp*a.stageX;
    unsigned* const gmJs = (unsigned*) (stage + BUILD_WARPS*(a.stageJ + 2*a.stageX));
    unsigned* const gmXs = gmJs + a.capJ/32;
    unsigned char* const keyBase = (unsigned char*) (gmXs + a.capX/32);
    unsigned char* const jkey = keyBase + warp*a.stageJ;
    unsigned char* const xkey = keyBase + BUILD_WARPS*a.stageJ + warp*a.stageX;
    __shared__ int counts[2][BUILD_WARPS];
    __shared__ int hist[BUILD_WARPS][256];      // per-warp histogram of the cluster masks, then the warps' write cursors
    __shared__ int scanTmp[33];
    __shared__ float4 cbox[16];                 // cluster c: cbox[2c] = (lo.xyz, hi.x), cbox[2c+1] = (hi.y, hi.z, -, -), relative to the block corner

    const int first = a.blkFirst[b], count = a.blkCount[b];
    const uint4 lo = a.blkLo[b], hi = a.blkHi[b];
    const int colI = (int) lo.w;
    const int cx = colI/a.ncy, cy = colI - cx*a.ncy;
    // bounding box in nm (absolute coordinates; only used for the conservative column/range walk)
    const float lox = lo.x*a.sx, loy = lo.y*a.sy, loz = lo.z*a.sz;
    const float hix = hi.x*a.sx, hiy = hi.y*a.sy, hiz = hi.z*a.sz;
    const float ex = (float) (hi.x - lo.x)*a.sx, ey = (float) (hi.y - lo.y)*a.sy, ez = (float) (hi.z - lo.z)*a.sz;
    const float R = a.reach, R2 = R*R;
    // bounding boxes of the block's 8 clusters (4 consecutive atoms each); an empty cluster reaches nothing
    if (warp == 0) {
        const uint4 q = a.posq[first + min(lane, count-1)];
        const bool have = lane < count;
        float cl[3], ch[3];
        cl[0] = ch[0] = (float) (q.x - lo.x)*a.sx; cl[1] = ch[1] = (float) (q.y - lo.y)*a.sy; cl[2] = ch[2] = (float) (q.z - lo.z)*a.sz;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if (!have) { cl[d] = 1.0e30f; ch[d] = -1.0e30f; }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                cl[d] = fminf(cl[d], __shfl_xor_sync(FULL_MASK, cl[d], o));
                ch[d] = fmaxf(ch[d], __shfl_xor_sync(FULL_MASK, ch[d], o));
            }
        }
        if ((lane & 3) == 0) {
            cbox[2*(lane >> 2)] = make_float4(cl[0], cl[1], cl[2], ch[0]);
            cbox[2*(lane >> 2) + 1] = make_float4(ch[1], ch[2], 0.f, 0.f);
        }
    }
    for (int k = threadIdx.x; k < a.capJ/32 + a.capX/32; k += blockDim.x) gmJs[k] = 0u;
    for (int k = threadIdx.x; k < BUILD_WARPS*256; k += blockDim.x) (&hist[0][0])[k] = 0;
    __syncthreads();
    // Unwrapped column indices (ux, uy) that can hold neighbours.  Image (kx, ky, kz) of the brick is displaced by
    // kx a + ky b + kz c, i.e. by (kx ax + ky bx + kz cx, ky by + kz cy, kz cz): seen from the columns of that image
    // the block's box sits at x - ky bx - kz cx, y - kz cy.  The candidate range covers every (ky, kz); the exact gap
    // test per image follows below.  (Rectangular boxes: bx = cx = cy = 0 and this is the plain 3 x 3 x 3 walk.)
    const float tiltX = fabsf(a.bx) + fabsf(a.cx), tiltY = fabsf(a.cy);
    const bool triclinic = tiltX != 0.f || tiltY != 0.f;
    const int uxLo = (int) floorf((lox - R - tiltX)/a.colWx), uxHi = (int) floorf((hix + R + tiltX)/a.colWx);
    const int uyLo = (int) floorf((loy - R - tiltY)/a.colWy), uyHi = (int) floorf((hiy + R + tiltY)/a.colWy);
    const int nuy = uyHi - uyLo + 1;
    const int nCand = (uxHi - uxLo + 1)*nuy;

    int nj = 0, nx = 0;
    bool overflow = false;
    for (int cand = warp; cand < nCand; cand += BUILD_WARPS) {
        const int ux = uxLo + cand/nuy, uy = uyLo + cand % nuy;
        const int kx = (ux + 2*a.ncx)/a.ncx - 2, ky = (uy + 2*a.ncy)/a.ncy - 2;     // floor division for ux >= -2 ncx
        const int wx = ux - kx*a.ncx, wy = uy - ky*a.ncy;
        if (kx < -2 || kx > 2 || ky < -1 || ky > 1 || ux < -2*a.ncx || uy < -2*a.ncy) continue;   // too far to interact
        if (!a.periodic && (kx != 0 || ky != 0)) continue;
        const int colJ = wx*a.ncy + wy;
        if (colJ < colI) continue;                                         // owned by the other block
        // gap between the block's box and the column in x, y (rectangular boxes: the same for every kz, tested once)
        float d2 = 0.f;
        if (!triclinic) {
            const float gx = fmaxf(0.f, fmaxf(ux*a.colWx - hix, lox - (ux+1)*a.colWx));
            const float gy = fmaxf(0.f, fmaxf(uy*a.colWy - hiy, loy - (uy+1)*a.colWy));
            d2 = gx*gx + gy*gy;
            if (d2 > R2) continue;
        }
        for (int kz = -1; kz <= 1; kz++) {
            if (!a.periodic && kz != 0) continue;
            if (triclinic) {
                const float offX = ky*a.bx + kz*a.cx, offY = kz*a.cy;        // the block's box as this image sees it
                const float gx = fmaxf(0.f, fmaxf(ux*a.colWx - (hix - offX), (lox - offX) - (ux+1)*a.colWx));
                const float gy = fmaxf(0.f, fmaxf(uy*a.colWy - (hiy - offY), (loy - offY) - (uy+1)*a.colWy));
                d2 = gx*gx + gy*gy;
                if (d2 > R2) continue;
            }
            const float dz = sqrtf(R2 - d2) + 1e-4f;
            const float zlo = loz - dz, zhi = hiz + dz;
            const float segLo = fmaxf(zlo, kz*a.Lz) - kz*a.Lz, segHi = fminf(zhi, (kz+1)*a.Lz) - kz*a.Lz;
            if (segHi < segLo) continue;
            const int zb0 = max(0, min(a.nzb-1, (int) floorf(segLo/a.binH)));
            const int zb1 = max(0, min(a.nzb-1, (int) floorf(segHi/a.binH)));
            int s = a.binStart[colJ*a.nzb + zb0];
            const int e = a.binStart[colJ*a.nzb + zb1 + 1];
            if (colJ == colI) s = max(s, first);                           // j must not precede the block
            const long long shx = ((long long) kx << 32) + ky*a.shiftB + kz*a.shiftCx;
            const long long shy = ((long long) ky << 32) + kz*a.shiftCy, shz = (long long) kz << 32;
            const int code = ((kx+2) + 5*((ky+1) + 3*(kz+1))) << J_SHIFT_BITS;
            for (int jb = s; jb < e; jb += 32*NL_PREFETCH) {
              // the candidates' positions and exclusion ranges of NL_PREFETCH chunks are requested together: the
              // walk is a chain of L2 round trips, and the ballots below only order the OUTPUT, not the loads
              uint4 qv[NL_PREFETCH];
              int2 rv[NL_PREFETCH];
#pragma unroll
              for (int u = 0; u < NL_PREFETCH; u++) {
                  const int j = jb + 32*u + lane;
                  if (j < e) { qv[u] = a.posq[j]; rv[u] = a.exclRange[j]; }
              }
#pragma unroll
              for (int u = 0; u < NL_PREFETCH; u++) {
                const int j0 = jb + 32*u;
                if (j0 >= e) break;                                        // warp-uniform
                const int j = j0 + lane;
                bool pass = false;
                unsigned imask = 0, cmask = 0;
                if (j < e) {
                    const uint4 q = qv[u];
                    const int2 range = rv[u];
                    const float rx = (float) ((long long) q.x + shx - (long long) lo.x)*a.sx;
                    const float ry = (float) ((long long) q.y + shy - (long long) lo.y)*a.sy;
                    const float rz = (float) ((long long) q.z + shz - (long long) lo.z)*a.sz;
                    const float ddx = fmaxf(0.f, fmaxf(-rx, rx - ex));
                    const float ddy = fmaxf(0.f, fmaxf(-ry, ry - ey));
                    const float ddz = fmaxf(0.f, fmaxf(-rz, rz - ez));
                    pass = ddx*ddx + ddy*ddy + ddz*ddz <= R2;
                    if (pass) {
                        // which clusters can it reach?  (none: the block's box is larger than the union of theirs)
#pragma unroll
                        for (int cl = 0; cl < 8; cl++) {
                            const float4 b0 = cbox[2*cl], b1 = cbox[2*cl+1];
                            const float ux_ = fmaxf(0.f, fmaxf(b0.x - rx, rx - b0.w));
                            const float uy_ = fmaxf(0.f, fmaxf(b0.y - ry, ry - b1.x));
                            const float uz_ = fmaxf(0.f, fmaxf(b0.z - rz, rz - b1.y));
                            if (ux_*ux_ + uy_*uy_ + uz_*uz_ <= R2) cmask |= 1u << cl;
                        }
                        pass = cmask != 0;
                    }
                    if (pass) {
                        const int rel = j - first;
                        if (rel >= 0 && rel < count) imask = 0xffffffffu << rel;      // own block: keep i < j only
                        if (range.y >= first && range.x < first + count) {
                            const int p = __float_as_int(a.par[j].w);
                            for (int k = a.exclStart[p]; k < a.exclStart[p+1]; k++) {
                                const int d = a.origToSorted[a.exclList[k]] - first;
                                if (d >= 0 && d < count) imask |= 1u << d;
                            }
                        }
                    }
                }
                const unsigned mJ = __ballot_sync(FULL_MASK, pass && imask == 0);
                const unsigned mX = __ballot_sync(FULL_MASK, pass && imask != 0);
                const unsigned below = (1u << lane) - 1u;
                if (pass) {
                    if (imask == 0) {
                        const int slot = nj + __popc(mJ & below);
                        if (slot < a.stageJ) { jbuf[slot] = code | j; jkey[slot] = (unsigned char) cmask; }
                    }
                    else {
                        const int slot = nx + __popc(mX & below);
                        if (slot < a.stageX) { xbuf[slot] = code | j; xmbuf[slot] = imask; xkey[slot] = (unsigned char) cmask; }
                    }
                }
                nj += __popc(mJ);
                nx += __popc(mX);
              }
            }
        }
    }
    if (nj > a.stageJ || nx > a.stageX) { overflow = true; nj = min(nj, a.stageJ); nx = min(nx, a.stageX); }
    if (lane == 0) { counts[0][warp] = nj; counts[1][warp] = nx; }
    __syncwarp();
    for (int k = lane; k < nj; k += 32) atomicAdd(&hist[warp][jkey[k]], 1);
    __syncthreads();
    // exclusive offsets of every (mask, warp) in mask-major order: thread t owns mask value t
    {
        const int t = threadIdx.x;
        int run = 0;
#pragma unroll
        for (int w = 0; w < BUILD_WARPS; w++) { const int v = hist[w][t]; hist[w][t] = run; run += v; }
        // block-wide exclusive scan of the 256 totals
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) scanTmp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int wv = lane < BUILD_WARPS ? scanTmp[lane] : 0;
            int winc = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL_MASK, winc, o); if (lane >= o) winc += u; }
            scanTmp[lane] = winc - wv;
        }
        __syncthreads();
        const int base = scanTmp[warp] + inc - run;
#pragma unroll
        for (int w = 0; w < BUILD_WARPS; w++) hist[w][t] += base;
    }
    __syncthreads();
    int offJ = 0, offX = 0, totJ = 0, totX = 0;
#pragma unroll
    for (int w = 0; w < BUILD_WARPS; w++) {
        if (w < warp) { offJ += counts[0][w]; offX += counts[1][w]; }
        totJ += counts[0][w];
        totX += counts[1][w];
    }
    if (totJ > a.capJ - 32 || totX > a.capX - 32) overflow = true;
    totJ = min(totJ, a.capJ - 32);
    totX = min(totX, a.capX - 32);
    int* jl = a.jlist + (size_t) lb*a.capJ;
    int* xl = a.xlist + (size_t) lb*a.capX;
    unsigned* xm = a.xmask + (size_t) lb*a.capX;
    // jlist: stable counting sort by cluster mask -- every warp places its own segment in order behind the
    // same-mask entries of the warps before it (hist[warp][mask] is this warp's cursor for that mask)
    for (int k0 = 0; k0 < nj; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < nj;
        const unsigned key = valid ? (unsigned) jkey[k] : 0x100u + lane;
        const unsigned peers = __match_any_sync(FULL_MASK, key);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int pos = 0;
        if (valid) {
            pos = hist[warp][key] + rank;
            if (pos < totJ) {
                jl[pos] = jbuf[k];
                atomicOr(&gmJs[pos >> 5], key << (8*((pos >> 3) & 3)));
            }
        }
        __syncwarp();
        if (valid && rank == 0) hist[warp][key] += __popc(peers);
        __syncwarp();
    }
    for (int k = lane; k < nx; k += 32)
        if (offX + k < totX) {
            const int pos = offX + k;
            xl[pos] = xbuf[k]; xm[pos] = xmbuf[k];
            atomicOr(&gmXs[pos >> 5], (unsigned) xkey[k] << (8*((pos >> 3) & 3)));
        }
    __syncthreads();
    for (int k = threadIdx.x; k < (totJ + 31) >> 5; k += blockDim.x) a.gmJ[(size_t) lb*(a.capJ/32) + k] = gmJs[k];
    for (int k = threadIdx.x; k < (totX + 31) >> 5; k += blockDim.x) a.gmX[(size_t) lb*(a.capX/32) + k] = gmXs[k];
    // pad the last tile of each list with invalid entries
    const int t = threadIdx.x;
    if (t < 32) {
        int padJ = ((totJ + 31) & ~31) - totJ;
        if (t < padJ) jl[totJ + t] = -1;
    }
    else if (t < 64) {
        int padX = ((totX + 31) & ~31) - totX;
        if (t-32 < padX) { xl[totX + t-32] = -1; xm[totX + t-32] = 0xffffffffu; }
    }
    if (t == 0) {
        a.jcount[lb] = totJ;
        a.xcount[lb] = totX;
        // work items for the pair kernel: the block's tiles (J tiles, then X tiles) in chunks
        const int tiles = ((totJ + 31) >> 5) + ((totX + 31) >> 5);
        const int n = (tiles + a.chunkTiles - 1)/a.chunkTiles;
        const int base = atomicAdd(a.itemCount, n);
        if (base + n > a.maxItems) overflow = true;      // (never with the host's sizing of `items`; redo rather than drop work)
        for (int k = 0; k < n; k++)
            if (base + k < a.maxItems) a.items[base + k] = make_int4(lb, k*a.chunkTiles, first, count);
    }
    if (overflow && lane == 0) { atomicOr(a.overflow, 1); a.overflowFlag[0] = 1.0; }
}

int launchBuildLists(Context& c) {
    const CellGeom& g = c.geom;
    if (c.blockWidth == 0) return NBS_OK;          // this rank has no direct-space share
    if (c.exclRangeForked) NBS_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.evExclDone, 0));     // launched beside the block construction
    else {
        int status = launchExclRange(c);
        if (status != NBS_OK) return status;
    }
    BuildArgs a;
    a.N = c.N; a.maxBlocks = c.maxBlocks; a.capJ = c.capJ; a.capX = c.capX;
    a.blockPeriod = c.blockPeriod; a.blockOffset = c.blockOffset; a.blockWidth = c.blockWidth;
    a.ncx = g.ncx; a.ncy = g.ncy; a.nzb = g.nzb;
    a.periodic = c.periodic ? 1 : 0;
    a.colWx = g.colW[0]; a.colWy = g.colW[1]; a.binH = g.binH;
    a.Lx = (float) g.box[0]; a.Ly = (float) g.box[1]; a.Lz = (float) g.box[2];
    a.bx = (float) g.tilt[0]; a.cx = (float) g.tilt[1]; a.cy = (float) g.tilt[2];
    a.shiftB = g.shiftB; a.shiftCx = g.shiftCx; a.shiftCy = g.shiftCy;
    a.sx = g.scale[0]; a.sy = g.scale[1]; a.sz = g.scale[2];
    a.reach = (float) (c.cutoffEff + c.skin) + 2e-4f;      // + skin: the list stays complete while no atom has moved more than skin/2
    a.counters = c.dCounters.d;
    a.blkFirst = c.dBlkFirst.d; a.blkCount = c.dBlkCount.d; a.blkLo = c.dBlkLo.d; a.blkHi = c.dBlkHi.d;
    a.binStart = c.dBinStart.d;
    a.posq = c.dPosq.d; a.par = c.dPar.d;
    a.exclRange = c.dExclRange.d; a.exclStart = c.dExclStart.d; a.exclList = c.dExclList.d; a.origToSorted = c.dOrigToSorted.d;
    a.jlist = c.dJList.d; a.jcount = c.dJCount.d; a.xlist = c.dXList.d; a.xmask = c.dXMask.d; a.xcount = c.dXCount.d;
    a.gmJ = c.dGmJ.d; a.gmX = c.dGmX.d;
    a.overflow = c.dCounters.d + 1;
    a.overflowFlag = c.dEnergy.d + 2*MAX_SLICES;
    a.itemCount = c.dCounters.d + 2;
    a.items = c.dItems.d;
    a.chunkTiles = c.chunkTiles;
    a.maxItems = (int) std::min<size_t>(c.dItems.cap, 0x7fffffff);
    a.stageJ = stageJ(c.capJ); a.stageX = stageX(c.capX);
    const size_t smem = sizeof(int)*(BUILD_WARPS*((size_t) a.stageJ + 2*(size_t) a.stageX) + c.capJ/32 + c.capX/32) +
                        BUILD_WARPS*((size_t) a.stageJ + (size_t) a.stageX);
    static bool attr[64] = {false};
    if (!attr[c.device & 63]) {
        NBS_CUDA_CHECK(cudaFuncSetAttribute(k_build_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, 216*1024));
        attr[c.device & 63] = true;
    }
    if (smem > 216*1024) return NBS_ERR_CAPACITY;
    k_build_lists<<<c.maxLocalBlocks, BUILD_WARPS*32, smem, c.stream>>>(a);
    NBS_CUDA_CHECK(cudaGetLastError());
    c.launches++;
    timerMark(c, "build_lists");
    return NBS_OK;
}

} // namespace nbs

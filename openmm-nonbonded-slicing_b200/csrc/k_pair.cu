// k_pair.cu -- tiled direct-space LJ + erfc-Coulomb pair kernel with per-slice lambda scaling.
//
// Arithmetic follows the reference's direct-space loop (platforms/reference/src/
// ReferenceSlicedLJCoulombIxn.cpp:367-445 for PME; :571-631 for the reaction-field cutoff) --
// per-slice energies are accumulated UNSCALED, forces are scaled by (lambda_vdW, lambda_Coulomb)
// of the pair's slice.  The structure is not the reference's (which injects a per-pair snippet into
// OpenMM's tile loop, platforms/common/src/kernels/coulombLennardJones.cc):
//   * one CTA per i-block, its warps share the block's tile list round-robin;
//   * a tile = 32 i atoms (one per lane, registers) x 32 j atoms staged in shared memory as
//     float4 (tile-relative position, charge) + float4 (sigma/2, 2 sqrt(eps), subset);
//   * lane l meets j slot (l + k) & 31 at step k, so both the i and the j force accumulate in
//     registers; the j accumulators rotate by one lane per step (3 shuffles);
//   * positions are 32-bit fixed-point fractional coordinates; each tile converts them once into
//     floats relative to the i-block's corner, so the cutoff test sees ~1e-7 nm resolution at any box
//     size, and a pair that lands within 2e-5 nm^2 of the cutoff is re-tested exactly in double from
//     the integers -- this is what makes the interacting-pair set bit-exact against the oracle;
//   * per-slice energies (EMODE 2, the default): the in-cutoff pairs of a tile are compacted with warp
//     ballots into a small shared-memory queue and their energies are evaluated in DOUBLE precision
//     from the exact integer coordinates, 32 real pairs at a time (no lane is wasted on pairs beyond the
//     cutoff).  Slice energies are sums of ~10^4..10^8 pair terms of both signs; single precision
//     cannot deliver 1e-5 of a small net value (DESIGN.md "Precision").  EMODE 1 keeps the cheaper
//     single-precision energies (2*NS float accumulators per lane), EMODE 0 computes forces only.
// Bound: FP32 pipe (no tensor-core shaped work here).
#include "nbs_internal.h"
#include "nbs_device.cuh"

namespace nbs {

struct PairArgs {
    int capJ, capX, Npad;
    int blockPeriod, blockOffset, blockWidth;   // this rank's share of the i-blocks
    float sx, sy, sz;
    double dsx, dsy, dsz;
    float rc2, alpha, krf, crf;
    float rswitch, rcut;
    int useSwitch;
    double rc2d, alphaD, krfD, crfD;
    const double* q64;                   // sorted charges * sqrt(ONE_4PI_EPS0), double
    const int* counters;
    const int* blkFirst; const int* blkCount; const uint4* blkLo;
    const uint4* posq; const float4* par;
    const int* jlist; const int* jcount; const int* xlist; const unsigned* xmask; const int* xcount;
    unsigned long long* force;
    double* energy;                      // [nSl][2]
    unsigned long long* pairStats;       // mode 1/2: [0] count, [1] hash
    int2* pairDump;                      // mode 2
    long long dumpCapacity;
    LambdaTable lam;
};

// erfc(x)*exp(x^2) for x in [0, 6]: degree-9 polynomial in t = 1/(1 + x/2), relative error 2.7e-7 in
// fp32 Horner form (fit and verified against scipy.special.erfcx).
__device__ __forceinline__ float erfcxPoly(float t) {
    float p = -3.701474935e-02f;
    p = fmaf(p, t, 1.652663209e-01f);
    p = fmaf(p, t, -2.075968035e-01f);
    p = fmaf(p, t, -9.388812420e-02f);
    p = fmaf(p, t, 2.824362380e-01f);
    p = fmaf(p, t, 3.656986947e-02f);
    p = fmaf(p, t, 3.008000226e-01f);
    p = fmaf(p, t, 2.698958094e-01f);
    p = fmaf(p, t, 2.836117033e-01f);
    p = fmaf(p, t, -8.030773936e-05f);
    return p;
}

// erfc(x)*exp(x^2), x in [0, 6], in double: degree-14 polynomial in u = (8 t - 5)/3, t = 1/(1 + x/2);
// relative error 1e-11 (fit against scipy.special.erfcx).
__device__ __forceinline__ double erfcxPolyD(double t) {
    const double u = fma(t, 2.6666666666666665, -1.6666666666666667);
    double p = 6.52049290501760624e-09;
    p = fma(p, u, 5.91610931414778049e-08);
    p = fma(p, u, -2.22270786854985114e-07);
    p = fma(p, u, -2.43147204892178188e-07);
    p = fma(p, u, 2.86160851006960621e-06);
    p = fma(p, u, -4.29950666125027918e-06);
    p = fma(p, u, -2.27129828446940741e-05);
    p = fma(p, u, 1.00552837184801405e-04);
    p = fma(p, u, 1.33588067694989746e-04);
    p = fma(p, u, -1.66536651372469141e-03);
    p = fma(p, u, -1.67024495580285893e-03);
    p = fma(p, u, 3.29934296618579967e-02);
    p = fma(p, u, 1.69407590984921058e-01);
    p = fma(p, u, 4.22187583608948647e-01);
    p = fma(p, u, 3.78537416928964254e-01);
    return p;
}

// Energy of one pair in double precision from the exact fixed-point coordinates (the wrapped integer
// difference IS the minimum image for any pair inside the cutoff).  Formulas: ReferenceSlicedLJCoulombIxn.cpp
// :376-396, 443-444 (PME) and :598-624 (reaction field), switch :380-384, 428-431.
template <bool IS_PME>
__device__ __forceinline__ void pairEnergyD(const uint4 fi, const uint4 fj, double qi, double qj, float sigi, float sigj,
                                            float epsi, float epsj, const PairArgs& a, double& ec, double& ev) {
    const double dx = (double) (int) (fj.x - fi.x)*a.dsx;
    const double dy = (double) (int) (fj.y - fi.y)*a.dsy;
    const double dz = (double) (int) (fj.z - fi.z)*a.dsz;
    const double r2 = dx*dx + dy*dy + dz*dz;
    double y = (double) rsqrtf((float) r2);
    y = y*fma(-0.5*r2*y, y, 1.5);
    y = y*fma(-0.5*r2*y, y, 1.5);
    const double r = r2*y;
    double s2 = ((double) sigi + (double) sigj)*y;
    s2 *= s2;
    const double s6 = s2*s2*s2;
    ev = (double) epsi*(double) epsj*(s6 - 1.0)*s6;
    if (a.useSwitch && r > (double) a.rswitch) {
        const double u = (r - (double) a.rswitch)/((double) a.rcut - (double) a.rswitch);
        ev *= 1.0 + u*u*u*(-10.0 + u*(15.0 - u*6.0));
    }
    const double qq = qi*qj;
    if (IS_PME) {
        const double x = a.alphaD*r;
        const double d = fma(0.5, x, 1.0);
        double t = (double) __frcp_rn((float) d);
        t = t*fma(-d, t, 2.0);
        t = t*fma(-d, t, 2.0);
        ec = qq*y*exp(-x*x)*erfcxPolyD(t);
    }
    else
        ec = qq*(y + a.krfD*r2 - a.crfD);
}

template <int NS>
__device__ __forceinline__ float pick(const float (&v)[NS], int s) {
    float r = v[0];
#pragma unroll
    for (int k = 1; k < NS; k++) r = (s == k) ? v[k] : r;
    return r;
}

// MODE 0: forces (+ energies when ENERGY); MODE 1: count + hash the interacting pairs; MODE 2: also dump them.
template <int NS, int EMODE, bool IS_PME, int MODE>
__global__ void __launch_bounds__(PAIR_WARPS*32) k_pair(const PairArgs a) {
    constexpr bool ENERGY = EMODE == 1;          // single-precision energies in the main loop
    constexpr int NE = NS*(NS+1);                // 2 * number of slices
    const int lb = blockIdx.x;                  // rank-local block index
    const int b = localToGlobalBlock(lb, a.blockPeriod, a.blockOffset, a.blockWidth);
    if (b >= a.counters[0]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float4 shPos[PAIR_WARPS][32];
    __shared__ float4 shPar[PAIR_WARPS][32];
    __shared__ float shF[PAIR_WARPS][3][32];
    __shared__ double shE[MAX_SLICES*2];
    // double-precision energy path: the i-block's exact coordinates / charges, the tile's, and the queue
    __shared__ uint4 shIFix[EMODE == 2 ? 32 : 1];
    __shared__ double shIQ[EMODE == 2 ? 32 : 1];
    __shared__ float2 shISE[EMODE == 2 ? 32 : 1];
    __shared__ uint4 shJFix[EMODE == 2 ? PAIR_WARPS : 1][32];
    __shared__ double shJQ[EMODE == 2 ? PAIR_WARPS : 1][32];
    __shared__ unsigned short shQueue[EMODE == 2 ? PAIR_WARPS : 1][64];

    const int first = a.blkFirst[b], cnt = a.blkCount[b];
    const uint4 lo = a.blkLo[b];
    const bool iValid = lane < cnt;
    const uint4 pi = iValid ? a.posq[first + lane] : lo;
    const float4 pari = iValid ? a.par[first + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    float xi = (float) (pi.x - lo.x)*a.sx, yi = (float) (pi.y - lo.y)*a.sy, zi = (float) (pi.z - lo.z)*a.sz;
    if (!iValid) xi = 1.0e8f;
    const float qi = iValid ? __uint_as_float(pi.w) : 0.f;
    const float sigi = pari.x, epsi = pari.y;
    const int si = __float_as_int(pari.z);
    const unsigned origI = (unsigned) __float_as_int(pari.w);
    float lamC[NS], lamV[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) { int sl = triSlice(si, s); lamC[s] = a.lam.c[sl]; lamV[s] = a.lam.v[sl]; }

    double acc[NE];                               // EMODE 2: [slice][term], dynamically indexed
    int qn = 0;
    if (EMODE == 2) {
#pragma unroll
        for (int k = 0; k < NE; k++) acc[k] = 0.0;
        if (warp == 0) {
            shIFix[lane] = make_uint4(pi.x, pi.y, pi.z, (unsigned) si);
            shIQ[lane] = iValid ? a.q64[first + lane] : 0.0;
            shISE[lane] = make_float2(sigi, epsi);
        }
        __syncthreads();
    }
    const unsigned below = (1u << lane) - 1u;
    // evaluate `count` queued pairs (one per lane) in double and add them to the lane's slice table
    auto drainQueue = [&](int count) {
        if (lane < count) {
            const unsigned e = shQueue[warp][lane];
            const int il = e >> 5, js = e & 31;
            const uint4 fi = shIFix[il], fj = shJFix[warp][js];
            const float2 sei = shISE[il];
            const float4 prj = shPar[warp][js];
            double ec, ev;
            pairEnergyD<IS_PME>(fi, fj, shIQ[il], shJQ[warp][js], sei.x, prj.x, sei.y, prj.y, a, ec, ev);
            const int sl = triSlice((int) fi.w, (int) fj.w);
            acc[2*sl] += ec;
            acc[2*sl+1] += ev;
        }
    };

    float fix = 0.f, fiy = 0.f, fiz = 0.f;
    float eC[NS], eV[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) { eC[s] = 0.f; eV[s] = 0.f; }
    unsigned long long nPairs = 0, hPairs = 0;

    const int nJ = a.jcount[lb], nX = a.xcount[lb];
    const int tJ = (nJ + 31) >> 5, tX = (nX + 31) >> 5;
    const int* jl = a.jlist + (size_t) lb*a.capJ;
    const int* xl = a.xlist + (size_t) lb*a.capX;
    const unsigned* xm = a.xmask + (size_t) lb*a.capX;
    const float TWO_OVER_SQRT_PI = 1.1283791670955126f;

    for (int t = warp; t < tJ + tX; t += PAIR_WARPS) {
        const bool isX = t >= tJ;
        const int* list = isX ? xl + (t - tJ)*32 : jl + t*32;
        const int entry = list[lane];
        unsigned imask = isX ? xm[(t - tJ)*32 + lane] : 0u;
        float4 pj = make_float4(-1.0e8f, 0.f, 0.f, 0.f), parj = make_float4(0.f, 0.f, 0.f, 0.f);
        int jIndex = 0;
        if (entry >= 0) {
            jIndex = entry & J_INDEX_MASK;
            const int code = entry >> J_SHIFT_BITS;
            const int kx = code % 3 - 1, ky = (code/3) % 3 - 1, kz = code/9 - 1;
            const uint4 q = a.posq[jIndex];
            parj = a.par[jIndex];
            pj.x = (float) ((long long) q.x + ((long long) kx << 32) - (long long) lo.x)*a.sx;
            pj.y = (float) ((long long) q.y + ((long long) ky << 32) - (long long) lo.y)*a.sy;
            pj.z = (float) ((long long) q.z + ((long long) kz << 32) - (long long) lo.z)*a.sz;
            pj.w = __uint_as_float(q.w);
        }
        __syncwarp();
        shPos[warp][lane] = pj;
        shPar[warp][lane] = parj;
        if (EMODE == 2) {
            uint4 fj = make_uint4(0u, 0u, 0u, 0u);
            double qj = 0.0;
            if (entry >= 0) {
                const uint4 q = a.posq[jIndex];
                fj = make_uint4(q.x, q.y, q.z, (unsigned) __float_as_int(parj.z));
                qj = a.q64[jIndex];
            }
            shJFix[warp][lane] = fj;
            shJQ[warp][lane] = qj;
        }
        __syncwarp();
        unsigned excluded = 0;                 // bit s: the pair (this lane's i, j slot s) is masked
        if (isX) {
#pragma unroll 4
            for (int bit = 0; bit < 32; bit++) {
                unsigned m = __ballot_sync(FULL_MASK, (imask >> bit) & 1u);
                if (lane == bit) excluded = m;
            }
        }
        float fjx = 0.f, fjy = 0.f, fjz = 0.f;
#pragma unroll 2
        for (int k = 0; k < 32; k++) {
            const int js = (lane + k) & 31;
            const float4 p = shPos[warp][js];
            const float4 pr = shPar[warp][js];
            const float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz*dz));
            bool in = r2 <= a.rc2;
            if (fabsf(r2 - a.rc2) < 2.0e-5f) {
                // borderline: decide from the exact integer coordinates, in double
                const int e2 = list[js];
                if (e2 >= 0 && iValid) {
                    const int code = e2 >> J_SHIFT_BITS;
                    const uint4 q = a.posq[e2 & J_INDEX_MASK];
                    const double ex = (double) ((long long) q.x + ((long long) (code % 3 - 1) << 32) - (long long) pi.x)*a.dsx;
                    const double ey = (double) ((long long) q.y + ((long long) ((code/3) % 3 - 1) << 32) - (long long) pi.y)*a.dsy;
                    const double ez = (double) ((long long) q.z + ((long long) (code/9 - 1) << 32) - (long long) pi.z)*a.dsz;
                    in = ex*ex + ey*ey + ez*ez <= a.rc2d;
                }
            }
            if (isX) in = in && !((excluded >> js) & 1u);
            if (EMODE == 2 && MODE == 0) {
                const unsigned m = __ballot_sync(FULL_MASK, in);
                if (m) {
                    if (in) shQueue[warp][qn + __popc(m & below)] = (unsigned short) ((lane << 5) | js);
                    qn += __popc(m);
                    if (qn >= 32) {
                        __syncwarp();
                        drainQueue(32);
                        const int rest = qn - 32;
                        const unsigned short moved = lane < rest ? shQueue[warp][32 + lane] : (unsigned short) 0;
                        __syncwarp();
                        if (lane < rest) shQueue[warp][lane] = moved;
                        qn = rest;
                    }
                }
            }
            if (MODE != 0) {
                if (in) {
                    const unsigned origJ = (unsigned) __float_as_int(pr.w);
                    const unsigned f = min(origI, origJ), s = max(origI, origJ);
                    nPairs++;
                    hPairs += pairHash(f, s);
                    if (MODE == 2) {
                        unsigned long long slot = atomicAdd(a.pairStats + 2, 1ull);
                        if ((long long) slot < a.dumpCapacity) a.pairDump[slot] = make_int2((int) f, (int) s);
                    }
                }
                continue;
            }
            float invR = rsqrtf(r2);
            invR = invR*fmaf(-0.5f*r2*invR, invR, 1.5f);            // one Newton step
            const float r = r2*invR;
            const float qq = qi*p.w;
            const float sig = sigi + pr.x;
            float s2 = sig*invR;
            s2 *= s2;
            const float s6 = s2*s2*s2;
            const float eps = epsi*pr.y;
            const float invR2 = invR*invR;
            float ev = eps*(s6 - 1.f)*s6;
            float fv = eps*fmaf(12.f, s6, -6.f)*s6*invR2;
            float ec, fc;
            if (IS_PME) {
                const float ar = a.alpha*r;
                const float ex = __expf(-ar*ar);
                const float tt = __fdividef(1.f, fmaf(0.5f, ar, 1.f));
                const float erfcv = ex*erfcxPoly(tt);
                const float qr = qq*invR;
                ec = qr*erfcv;
                fc = qr*invR2*fmaf(TWO_OVER_SQRT_PI*ar, ex, erfcv);
            }
            else {
                ec = qq*(invR + a.krf*r2 - a.crf);
                fc = qq*invR2*(invR - 2.f*a.krf*r2);
            }
            if (a.useSwitch && r > a.rswitch) {
                const float w = 1.f/(a.rcut - a.rswitch);
                const float u = (r - a.rswitch)*w;
                const float sv = 1.f + u*u*u*(-10.f + u*(15.f - u*6.f));
                const float sd = u*u*(-30.f + u*(60.f - u*30.f))*w;
                fv = sv*fv - ev*sd*invR;
                ev *= sv;
            }
            const int sj = __float_as_int(pr.z);
            float dEdR = pick<NS>(lamV, sj)*fv + pick<NS>(lamC, sj)*fc;
            dEdR = in ? dEdR : 0.f;
            fix = fmaf(dEdR, dx, fix); fiy = fmaf(dEdR, dy, fiy); fiz = fmaf(dEdR, dz, fiz);
            fjx = fmaf(-dEdR, dx, fjx); fjy = fmaf(-dEdR, dy, fjy); fjz = fmaf(-dEdR, dz, fjz);
            if (ENERGY) {
                ec = in ? ec : 0.f;
                ev = in ? ev : 0.f;
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    eC[s] += (sj == s) ? ec : 0.f;
                    eV[s] += (sj == s) ? ev : 0.f;
                }
            }
            const int src = (lane + 1) & 31;
            fjx = __shfl_sync(FULL_MASK, fjx, src);
            fjy = __shfl_sync(FULL_MASK, fjy, src);
            fjz = __shfl_sync(FULL_MASK, fjz, src);
        }
        if (EMODE == 2 && MODE == 0) {              // the queue refers to this tile's shared-memory slots
            __syncwarp();
            drainQueue(qn);
            qn = 0;
            __syncwarp();
        }
        if (MODE == 0 && entry >= 0) {
            atomicAdd(a.force + jIndex, toFixed(fjx));
            atomicAdd(a.force + a.Npad + jIndex, toFixed(fjy));
            atomicAdd(a.force + 2*(size_t) a.Npad + jIndex, toFixed(fjz));
        }
    }

    if (MODE != 0) {
        nPairs = (unsigned long long) warpSum((double) nPairs);        // exact below 2^53
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hPairs += __shfl_xor_sync(FULL_MASK, hPairs, o);
        if (lane == 0) {
            atomicAdd(a.pairStats, nPairs);
            atomicAdd(a.pairStats + 1, hPairs);
        }
        return;
    }

    // fold the i forces of the CTA's warps in a fixed order, then one fixed-point atomic per atom
    shF[warp][0][lane] = fix; shF[warp][1][lane] = fiy; shF[warp][2][lane] = fiz;
    if (EMODE != 0 && threadIdx.x < MAX_SLICES*2) shE[threadIdx.x] = 0.0;
    __syncthreads();
    if (threadIdx.x < 96) {
        const int comp = threadIdx.x >> 5;
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < PAIR_WARPS; w++) sum += shF[w][comp][lane];
        if (lane < cnt) atomicAdd(a.force + (size_t) comp*a.Npad + first + lane, toFixed(sum));
    }
    if (ENERGY) {
        // slice(sa, sb) gets eX[sb] of lanes whose atom is in sa, and eX[sa] of lanes in sb (sa != sb)
#pragma unroll
        for (int sa = 0; sa < NS; sa++)
#pragma unroll
            for (int sb = sa; sb < NS; sb++) {
                double c = 0.0, v = 0.0;
                if (si == sa) { c += eC[sb]; v += eV[sb]; }
                if (sa != sb && si == sb) { c += eC[sa]; v += eV[sa]; }
                c = warpSum(c);
                v = warpSum(v);
                if (lane == 0) {
                    const int sl = sb*(sb+1)/2 + sa;
                    atomicAdd(&shE[2*sl], c);
                    atomicAdd(&shE[2*sl+1], v);
                }
            }
        __syncthreads();
        if (threadIdx.x < NS*(NS+1)) atomicAdd(a.energy + threadIdx.x, shE[threadIdx.x]);
    }
    if (EMODE == 2) {
#pragma unroll
        for (int k = 0; k < NE; k++) {
            const double v = warpSum(acc[k]);
            if (lane == 0 && v != 0.0) atomicAdd(&shE[k], v);
        }
        __syncthreads();
        if (threadIdx.x < NE && shE[threadIdx.x] != 0.0) atomicAdd(a.energy + threadIdx.x, shE[threadIdx.x]);
    }
}

template <int NS>
static void launchPairNS(Context& c, const PairArgs& a, int emode, bool pme) {
    dim3 grid(c.maxLocalBlocks), block(PAIR_WARPS*32);
    cudaStream_t st = c.stream;
    if (pme) {
        if (emode == 0) k_pair<NS, 0, true, 0><<<grid, block, 0, st>>>(a);
        else if (emode == 1) k_pair<NS, 1, true, 0><<<grid, block, 0, st>>>(a);
        else k_pair<NS, 2, true, 0><<<grid, block, 0, st>>>(a);
    }
    else {
        if (emode == 0) k_pair<NS, 0, false, 0><<<grid, block, 0, st>>>(a);
        else if (emode == 1) k_pair<NS, 1, false, 0><<<grid, block, 0, st>>>(a);
        else k_pair<NS, 2, false, 0><<<grid, block, 0, st>>>(a);
    }
}

int launchPairs(Context& c, bool wantEnergy, int mode) {
    if (c.blockWidth == 0) return NBS_OK;          // this rank has no direct-space share
    const CellGeom& g = c.geom;
    PairArgs a;
    a.capJ = c.capJ; a.capX = c.capX; a.Npad = c.Npad;
    a.blockPeriod = c.blockPeriod; a.blockOffset = c.blockOffset; a.blockWidth = c.blockWidth;
    a.sx = g.scale[0]; a.sy = g.scale[1]; a.sz = g.scale[2];
    a.dsx = g.box[0]/4294967296.0; a.dsy = g.box[1]/4294967296.0; a.dsz = g.box[2]/4294967296.0;
    a.rc2 = (float) (c.cutoff*c.cutoff);
    a.rc2d = c.cutoff*c.cutoff;
    a.alphaD = c.alpha;
    a.krfD = pow(c.cutoff, -3.0)*(c.rfDielectric - 1.0)/(2.0*c.rfDielectric + 1.0);
    a.crfD = (1.0/c.cutoff)*(3.0*c.rfDielectric)/(2.0*c.rfDielectric + 1.0);
    a.q64 = c.dQ64.d;
    a.alpha = (float) c.alpha;
    // ReferenceSlicedLJCoulombIxn::setUseCutoff, ReferenceSlicedLJCoulombIxn.cpp:60-68
    a.krf = (float) (pow(c.cutoff, -3.0)*(c.rfDielectric - 1.0)/(2.0*c.rfDielectric + 1.0));
    a.crf = (float) ((1.0/c.cutoff)*(3.0*c.rfDielectric)/(2.0*c.rfDielectric + 1.0));
    a.useSwitch = c.useSwitch ? 1 : 0;
    a.rswitch = (float) c.switchDist;
    a.rcut = (float) c.cutoff;
    a.counters = c.dCounters.d;
    a.blkFirst = c.dBlkFirst.d; a.blkCount = c.dBlkCount.d; a.blkLo = c.dBlkLo.d;
    a.posq = c.dPosq.d; a.par = c.dPar.d;
    a.jlist = c.dJList.d; a.jcount = c.dJCount.d; a.xlist = c.dXList.d; a.xmask = c.dXMask.d; a.xcount = c.dXCount.d;
    a.force = c.dForce.d;
    a.energy = c.dEnergy.d;
    a.pairStats = c.dPairStats.d;
    a.pairDump = c.dPairDump.d;
    a.dumpCapacity = (long long) c.dPairDump.cap;
    for (int s = 0; s < MAX_SLICES; s++) {
        a.lam.c[s] = s < c.nSl ? (float) c.lambdas[2*s] : 1.f;
        a.lam.v[s] = s < c.nSl ? (float) c.lambdas[2*s+1] : 1.f;
    }
    const bool pme = c.method == NBS_METHOD_PME;
    dim3 grid(c.maxLocalBlocks), block(PAIR_WARPS*32);
    const int emode = !wantEnergy ? 0 : ((c.flags & NBS_FLAG_FP32_ENERGY) ? 1 : 2);
    if (mode == 1) k_pair<1, 0, true, 1><<<grid, block, 0, c.stream>>>(a);
    else if (mode == 2) k_pair<1, 0, true, 2><<<grid, block, 0, c.stream>>>(a);
    else {
        switch (c.nS) {
            case 1: launchPairNS<1>(c, a, emode, pme); break;
            case 2: launchPairNS<2>(c, a, emode, pme); break;
            case 3: launchPairNS<3>(c, a, emode, pme); break;
            case 4: launchPairNS<4>(c, a, emode, pme); break;
            default: launchPairNS<MAX_SUBSETS>(c, a, emode, pme); break;
        }
    }
    c.launches++;
    timerMark(c, mode == 0 ? "pair" : "pair_set");
    return NBS_OK;
}

} // namespace nbs

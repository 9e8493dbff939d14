// k_pair.cu -- tiled direct-space LJ + erfc-Coulomb pair kernel with per-slice lambda scaling.
//
// Arithmetic follows the reference's direct-space loop (platforms/reference/src/
// ReferenceSlicedLJCoulombIxn.cpp:367-445 for PME; :571-631 for the reaction-field cutoff) --
// per-slice energies are accumulated UNSCALED, forces are scaled by (lambda_vdW, lambda_Coulomb)
// of the pair's slice.  The structure is not the reference's (which injects a per-pair snippet into
// OpenMM's 32x32 tile loop, platforms/common/src/kernels/coulombLennardJones.cc, and pays the full
// pair arithmetic for every one of the 1024 pairs of a tile although only ~30% are inside the cutoff):
//
//   * WORK ITEMS, PERSISTENT WARPS.  The list builder emits items (i-block, first tile, <= chunk tiles);
//     every warp of a persistent grid pulls items from a global cursor, so there is no CTA-wide barrier,
//     no tail of idle warps, and small systems (818 i-blocks at DHFR size) still fill 148 SMs.
//   * a tile = 32 i atoms (one per lane, registers) x 32 j atoms staged in shared memory as float4
//     (position relative to the block corner, charge) + float4 (sigma/2, 2 sqrt(eps), subset); lane l
//     meets j slot (l + k) & 31 at step k, so the i AND the j forces accumulate in registers without
//     conflicts -- the j accumulators rotate one lane per step (3 shuffles).  The step is ~60 fp32
//     instructions: approx rsqrt / rcp / ex2 (MUFU) with no denormal handling, erfc(x) = exp(-x^2) P(t),
//     lambda pair from a shared-memory row selected by the lane's own subset.
//   * slice energies (EMODE 2, the default when energies are requested): the in-cutoff pairs of a tile
//     are compacted with warp ballots into a per-warp queue and their energies are evaluated in DOUBLE
//     precision from the exact fixed-point coordinates, 32 REAL pairs per pass (no lane is wasted on pairs
//     beyond the cutoff).  Slice energies are sums of 10^4..10^8 terms of both signs; single precision
//     cannot deliver 1e-5 of a small net value (DESIGN.md "Precision").  EMODE 1 keeps single-precision
//     pair energies (the plugin's "single" mode), EMODE 0 computes forces only.
//   * summation order is fixed (per lane in step order; tiles and items are combined as 64-bit fixed
//     point, whose adds commute), so forces are bit-reproducible.
//   * positions are 32-bit fixed-point fractional coordinates; each tile converts them once into floats
//     relative to the i-block's corner (~1e-7 nm resolution at any box size); a pair whose fp32 r^2 lands
//     within 2e-5 nm^2 of the cutoff is re-tested exactly in double from the integers -- this is what
//     makes the interacting-pair set bit-exact against the oracle.
// Bound: FP32 / issue rate (no tensor-core shaped work here).
#include "nbs_internal.h"
#include "nbs_device.cuh"
#include <algorithm>
#include <cstdlib>

// tuning knobs (profiles/README.md has the measurements behind the defaults)
#ifndef PAIR_MIN_CTAS
#define PAIR_MIN_CTAS 2          // resident CTAs per SM the register allocation is bounded for
#endif
#ifndef PAIR_UNROLL
#define PAIR_UNROLL 2            // steps of the tile loop in flight per warp
#endif

namespace nbs {

constexpr int kPairUnroll = PAIR_UNROLL;

struct PairArgs {
    int capJ, capX, Npad;
    int blockPeriod, blockOffset, blockWidth;   // this rank's share of the i-blocks
    int chunkTiles;                             // tiles per work item
    long long shiftB, shiftCx, shiftCy;         // triclinic image shifts (CellGeom), fixed-point units; 0 for a rectangular box
    int nE;                                     // 2 * number of slices
    float sx, sy, sz;
    double dsx, dsy, dsz;
    float rc2, alpha, krf, crf;
    float rswitch, rcut;
    int useSwitch;
    double rc2d, alphaD, krfD, crfD;
    float dalpha2, invCut6, shiftMult;   // LJPME: alpha_d^2, rc^-6, rc^-6 (1 - exp(-x)(1 + x + x^2/2)) at x = (alpha_d rc)^2
    const double* q64;                   // sorted charges * sqrt(ONE_4PI_EPS0), double
    const double* erfcTab;               // piecewise degree-4 fit of erfc(alpha sqrt(s))/sqrt(s) in s = r^2 (ERFC_TAB_ROW doubles per interval)
    int* counters;                       // [2] number of work items, [3] cursor
    const int4* items;                   // (local block, first tile, first atom of the block, atoms in the block)
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

// per-warp shared memory
struct __align__(16) WarpScratch {
    float4 iPos[32];      // i-block: position relative to the block corner, charge*sqrt(K)
    float4 iPar[32];      // sigma/2, 2 sqrt(eps), subset, particle index
    float4 jPos[32];      // current tile
    float4 jPar[32];
    uint4 iFix[32];       // exact coordinates, w = subset
    uint4 jFix[32];
    double iQ[32];        // charges in double (energy path)
    double jQ[32];
    unsigned jMask[32];   // exclusion-list tiles: bit l set = pair (i lane l, this j) is masked
    unsigned short queue[1024 + 32];   // MODE 1/2: a whole tile's pairs; MODE 0: the energy queue (<= 64 used)
};

__device__ __forceinline__ float rsqrtFast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpFast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2Fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

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
// (coefficients live in constant memory: a 64-bit literal costs two uniform moves per use, a constant-bank
// operand costs nothing)
__constant__ double kErfcxD[15] = {
    6.52049290501760624e-09, 5.91610931414778049e-08, -2.22270786854985114e-07, -2.43147204892178188e-07,
    2.86160851006960621e-06, -4.29950666125027918e-06, -2.27129828446940741e-05, 1.00552837184801405e-04,
    1.33588067694989746e-04, -1.66536651372469141e-03, -1.67024495580285893e-03, 3.29934296618579967e-02,
    1.69407590984921058e-01, 4.22187583608948647e-01, 3.78537416928964254e-01};
__constant__ double kExpD[12] = {
    2.50521083854417188e-08, 2.75573192239858907e-07, 2.75573192239858907e-06, 2.48015873015873016e-05,
    1.98412698412698413e-04, 1.38888888888888894e-03, 8.33333333333333322e-03, 4.16666666666666644e-02,
    1.66666666666666657e-01, 0.5, 1.0, 1.0};
__constant__ double kMiscD[6] = {2.6666666666666665, -1.6666666666666667, 1.4426950408889634074, 6755399441055744.0,
                                 -0.693147180559945286, -2.31904681384629956e-17};

__device__ __forceinline__ double erfcxPolyD(double t) {
    const double u = fma(t, kMiscD[0], kMiscD[1]);
    double p = kErfcxD[0];
#pragma unroll
    for (int k = 1; k < 15; k++) p = fma(p, u, kErfcxD[k]);
    return p;
}

// exp(-z) for z in [0, 60], double, ~3e-16 relative: 2^n * e^h with a degree-11 Taylor polynomial on
// |h| <= ln(2)/2 -- the library exp() minus the special cases this kernel cannot hit.
__device__ __forceinline__ double expNegD(double z) {
    const double u = -z*kMiscD[2];                              // log2(e)
    const double shifter = kMiscD[3];                           // 1.5 * 2^52: rounds to nearest integer
    const double n = (u + shifter) - shifter;
    const double g = fma(n, kMiscD[4], -z);                     // -z - n ln2 (hi part of ln2)
    const double h = fma(n, kMiscD[5], g);                      // ... lo part
    double p = kExpD[0];                                        // 1/11! ... Taylor coefficients of exp
#pragma unroll
    for (int k = 1; k < 12; k++) p = fma(p, h, kExpD[k]);
    const int ni = (int) n;
    return __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
}

// Energy of one pair from the exact fixed-point coordinates (the wrapped integer difference IS the
// minimum image for any pair inside the cutoff).  Formulas: ReferenceSlicedLJCoulombIxn.cpp:376-396, 443-444
// (PME) and :598-624 (reaction field), switch :380-384, 428-431.
//   * Coulomb, PME: K q_i q_j erfc(alpha r)/r in DOUBLE.  Double-precision instructions are ~8x more
//     expensive to issue than fp32 ones here, so instead of rsqrt + exp + erfcx (about 45 of them) the
//     function f(s) = erfc(alpha sqrt(s))/sqrt(s), s = r^2, comes from a table of degree-4 polynomials
//     on 256 intervals per octave of s (built on the host from libm's erfc: relative error ~1e-11;
//     csrc/nbs_api.cu buildErfcTable) -- three 16-byte loads and 5 fused multiply-adds.  Pairs closer than
//     2^-3.5 nm (0.088 nm: none in a physical system) take the analytic path.
//   * Lennard-Jones: fp32 from the same exact r^2 (terms of one sign dominate a slice's vdW sum, so 1e-7
//     per term is far inside the 1e-5 target), accumulated in double.
//   * LJPME (CMODE 2, :398-426): the multiplicative C6 term that the dispersion grid carries is taken out in real
//     space, plus the potential shift at the cutoff; fp32 like the rest of the LJ energy.
// CMODE: 0 = reaction field / no cutoff, 1 = PME or Ewald, 2 = LJPME.
template <int CMODE>
__device__ __forceinline__ void pairEnergyD(const uint4 fi, const uint4 fj, double qi, double qj, float sigi, float sigj,
                                            float epsi, float epsj, const PairArgs& a, double& ec, double& ev) {
    constexpr bool IS_PME = CMODE != 0;
    const double dx = (double) (int) (fj.x - fi.x)*a.dsx;
    const double dy = (double) (int) (fj.y - fi.y)*a.dsy;
    const double dz = (double) (int) (fj.z - fi.z)*a.dsz;
    const double r2 = dx*dx + dy*dy + dz*dz;
    const float r2f = (float) r2;
    float yf = rsqrtFast(r2f);
    yf = yf*fmaf(-0.5f*r2f*yf, yf, 1.5f);              // fp32 Newton step: ~1e-7
    {
        float s2 = (sigi + sigj)*yf;
        s2 *= s2;
        const float s6 = s2*s2*s2;
        float evf = epsi*epsj*(s6 - 1.f)*s6;
        if (CMODE == 2) {
            const float sg = sigi*sigj;
            const float c6 = 64.f*sg*sg*sg*epsi*epsj;                  // c6_i c6_j, c6 = 8 (sigma/2)^3 (2 sqrt(eps))
            const float dar2 = a.dalpha2*r2f;
            const float y2 = yf*yf;
            const float emult = c6*y2*y2*y2*(1.f - ex2Fast(-1.4426950408889634f*dar2)*fmaf(dar2, fmaf(0.5f, dar2, 1.f), 1.f));
            float sc = sigi + sigj;
            sc *= sc;
            const float sc6 = sc*sc*sc*a.invCut6;
            evf += emult + epsi*epsj*(1.f - sc6)*sc6 - c6*a.shiftMult;
        }
        if (a.useSwitch) {
            const float r = r2f*yf;
            if (r > a.rswitch) {
                const float u = (r - a.rswitch)/(a.rcut - a.rswitch);
                evf *= 1.f + u*u*u*(-10.f + u*(15.f - u*6.f));
            }
        }
        ev = (double) evf;
    }
    const double qq = qi*qj;
    if (IS_PME) {
        const unsigned bits = __float_as_uint(r2f);
        const int idx = (int) (bits >> (23 - ERFC_TAB_PER_OCTAVE_LOG2)) - ERFC_TAB_BASE;
        if (idx >= 0) {
            const double2* row = reinterpret_cast<const double2*>(a.erfcTab + (size_t) idx*ERFC_TAB_ROW);
            const double2 c0 = __ldg(row), c1 = __ldg(row + 1), c2 = __ldg(row + 2);
            // 2 / (interval width) = 2^(1 + 8 - e), e = unbiased exponent of r2f: built straight into the exponent field
            const double scale = __hiloint2double((1023 + 1 + ERFC_TAB_PER_OCTAVE_LOG2 + 127 - (int) (bits >> 23)) << 20, 0);
            const double d = fma(r2, scale, c0.x);                  // position inside the interval, [-1, 1]
            double p = fma(c0.y, d, c1.x);
            p = fma(p, d, c1.y);
            p = fma(p, d, c2.x);
            p = fma(p, d, c2.y);
            ec = qq*p;
        }
        else {
            double y = (double) yf;
            y = y*fma(-0.5*r2*y, y, 1.5);               // double Newton step: ~1e-14
            const double x = a.alphaD*r2*y;
            const double dd = fma(0.5, x, 1.0);
            double t = (double) rcpFast((float) dd);
            t = t*fma(-dd, t, 2.0);
            t = t*fma(-dd, t, 2.0);
            ec = qq*y*expNegD(x*x)*erfcxPolyD(t);
        }
    }
    else {
        double y = (double) yf;
        y = y*fma(-0.5*r2*y, y, 1.5);
        ec = qq*(y + a.krfD*r2 - a.crfD);
    }
}

// Exact cutoff test from the fixed-point coordinates (the wrapped integer difference is the minimum image
// for any pair near the cutoff, because the box is at least twice the cutoff).
__device__ __forceinline__ bool exactInRange(const uint4 fj, const uint4 fi, const PairArgs& a) {
    const double ex = (double) (int) (fj.x - fi.x)*a.dsx;
    const double ey = (double) (int) (fj.y - fi.y)*a.dsy;
    const double ez = (double) (int) (fj.z - fi.z)*a.dsz;
    return ex*ex + ey*ey + ez*ez <= a.rc2d;
}

// The one definition of "pair (this lane's i, j slot js) interacts": r^2 <= rc^2 (decided exactly when
// the fp32 value is within 2e-5 nm^2 of the cutoff) and not masked by an exclusion.
__device__ __forceinline__ bool pairInRange(const WarpScratch& w, const PairArgs& a, float xi, float yi, float zi,
                                            const uint4 pi, int js, bool isX, int lane) {
    const float4 p = w.jPos[js];
    const float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
    const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz*dz));
    bool in = r2 <= a.rc2;
    if (fabsf(r2 - a.rc2) < 2.0e-5f) in = exactInRange(w.jFix[js], pi, a);
    if (isX) in = in && !((w.jMask[js] >> lane) & 1u);
    return in;
}

// One 32 x 32 tile of the force kernel: lane l meets j slot (l + k) & 31 at step k; i forces (fix, fiy, fiz)
// and the rotating j forces (fjx, fjy, fjz) stay in registers.  EMODE 2 also compacts the in-cutoff pairs
// into w.queue and evaluates their energies in double precision, 32 real pairs per pass.
template <int EMODE, int CMODE>
__device__ __forceinline__ void energyPass(const WarpScratch& w, const PairArgs& a, int lane, int count, double* acc) {
    if (lane < count) {
        const unsigned e = w.queue[lane];
        const int il = e >> 5, jq = e & 31;
        // (sigma/2, 2 sqrt(eps)) only: 8-byte loads; the subsets ride in the .w of the exact coordinates
        const float2 q1 = *reinterpret_cast<const float2*>(&w.iPar[il]), q2 = *reinterpret_cast<const float2*>(&w.jPar[jq]);
        const uint4 fi = w.iFix[il], fj = w.jFix[jq];
        double ecd, evd;
        pairEnergyD<CMODE>(fi, fj, w.iQ[il], w.jQ[jq], q1.x, q2.x, q1.y, q2.y, a, ecd, evd);
        const int sl = triSlice((int) fi.w, (int) fj.w);
        acc[2*sl] += ecd;
        acc[2*sl+1] += evd;
    }
}

template <int EMODE, int CMODE, bool IS_X, bool SWITCH>
__device__ __forceinline__ void tileLoop(WarpScratch& w, const PairArgs& a, const float2* shLam, int lamOff, int lane,
                                         float xi, float yi, float zi, float qi, float sigi, float epsi, int si, const uint4 pi,
                                         float& fix, float& fiy, float& fiz, float& fjx, float& fjy, float& fjz, double* acc) {
    constexpr bool IS_PME = CMODE != 0;
    const unsigned below = (1u << lane) - 1u;
    const float rc2 = a.rc2, alpha = a.alpha;
    const float TWO_OVER_SQRT_PI = 1.1283791670955126f;
    const int src = (lane + 1) & 31;
    int qn = 0;                                   // pairs waiting in the energy queue
#pragma unroll kPairUnroll
    for (int k = 0; k < 32; k++) {
        const int js = (lane + k) & 31;
        const float4 p = w.jPos[js];
        const float4 pr = w.jPar[js];
        const float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
        const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz*dz));
        bool in = r2 <= rc2;
        if (fabsf(r2 - rc2) < 2.0e-5f) in = exactInRange(w.jFix[js], pi, a);     // borderline: rare
        if (IS_X) in = in && !((w.jMask[js] >> lane) & 1u);
        if (EMODE == 2) {
            const unsigned m = __ballot_sync(FULL_MASK, in);
            if (in) w.queue[qn + __popc(m & below)] = (unsigned short) ((lane << 5) | js);
            qn += __popc(m);
            if (qn >= 32) {                       // warp-uniform
                __syncwarp();
                energyPass<EMODE, CMODE>(w, a, lane, 32, acc);
                const int rest = qn - 32;
                const unsigned short moved = lane < rest ? w.queue[32 + lane] : (unsigned short) 0;
                __syncwarp();
                if (lane < rest) w.queue[lane] = moved;
                qn = rest;
            }
        }
        const float invR = rsqrtFast(r2);
        const float r = r2*invR;
        const float invR2 = invR*invR;
        float s2 = (sigi + pr.x)*invR;
        s2 *= s2;
        const float s6 = s2*s2*s2;
        const float eps = epsi*pr.y;
        float ev = eps*(s6 - 1.f)*s6;
        float fv = eps*fmaf(12.f, s6, -6.f)*s6*invR2;
        const float qr = qi*p.w*invR;
        float ec, fc;
        if (IS_PME) {
            const float ar = alpha*r;
            const float ex = ex2Fast(-1.4426950408889634f*ar*ar);
            const float tt = rcpFast(fmaf(0.5f, ar, 1.f));
            const float erfcv = ex*erfcxPoly(tt);
            ec = qr*erfcv;
            fc = qr*invR2*fmaf(TWO_OVER_SQRT_PI*ar, ex, erfcv);
        }
        else {
            ec = qr*fmaf(a.krf*r2, r, 1.f) - qi*p.w*a.crf;
            fc = qr*invR2*fmaf(-2.f*a.krf*r2, r, 1.f);
        }
        if (CMODE == 2) {
            // LJPME (:398-426): real-space share of the multiplicative C6 term; the shift only enters the energy
            const float sg = sigi*pr.x;
            const float c6 = 64.f*sg*sg*sg*eps;
            const float dar2 = a.dalpha2*r2;
            const float exd = ex2Fast(-1.4426950408889634f*dar2);
            const float p2 = fmaf(dar2, fmaf(0.5f, dar2, 1.f), 1.f);             // 1 + x + x^2/2
            const float c6r6 = c6*invR2*invR2*invR2;
            fv = fmaf(6.f*c6r6*invR2, 1.f - exd*fmaf(dar2*dar2*dar2, 1.f/6.f, p2), fv);
            if (EMODE == 1) {
                float sc = sigi + pr.x;
                sc *= sc;
                const float sc6 = sc*sc*sc*a.invCut6;
                ev += c6r6*(1.f - exd*p2) + eps*(1.f - sc6)*sc6 - c6*a.shiftMult;
            }
        }
        if (SWITCH) {
            if (r > a.rswitch) {
                const float wd = 1.f/(a.rcut - a.rswitch);
                const float u = (r - a.rswitch)*wd;
                const float sv = 1.f + u*u*u*(-10.f + u*(15.f - u*6.f));
                const float sd = u*u*(-30.f + u*(60.f - u*30.f))*wd;
                fv = sv*fv - ev*sd*invR;
                ev *= sv;
            }
        }
        const int sj = __float_as_int(pr.z);
        const float2 lam = shLam[lamOff + sj];
        float dEdR = fmaf(lam.y, fv, lam.x*fc);
        dEdR = in ? dEdR : 0.f;
        fix = fmaf(dEdR, dx, fix); fiy = fmaf(dEdR, dy, fiy); fiz = fmaf(dEdR, dz, fiz);
        fjx = fmaf(-dEdR, dx, fjx); fjy = fmaf(-dEdR, dy, fjy); fjz = fmaf(-dEdR, dz, fjz);
        if (EMODE == 1 && in) {
            const int sl = triSlice(si, sj);
            acc[2*sl] += (double) ec;
            acc[2*sl+1] += (double) ev;
        }
        fjx = __shfl_sync(FULL_MASK, fjx, src);
        fjy = __shfl_sync(FULL_MASK, fjy, src);
        fjz = __shfl_sync(FULL_MASK, fjz, src);
    }
    if (EMODE == 2) {                             // the queue refers to this tile's shared-memory slots
        __syncwarp();
        energyPass<EMODE, CMODE>(w, a, lane, qn, acc);
    }
}

// MODE 0: forces (+ energies per EMODE); MODE 1: count + hash the interacting pairs; MODE 2: also dump them.
// EMODE 0: forces only; 1: single-precision energies; 2: double-precision energies.
// MINCTAS: resident CTAs per SM the register allocation is bounded for (2: 128 registers; 3: 80 registers and a
// few bytes of spill -- no different at DHFR size, 8 % faster at STMV size, where there is always a next item)
template <int EMODE, int CMODE, int MODE, int MINCTAS>
__global__ void __launch_bounds__(PAIR_WARPS*32, MINCTAS) k_pair(const PairArgs a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ float2 shLam[MAX_SUBSETS*MAX_SUBSETS];      // (lambda_Coulomb, lambda_vdW) of subset pair (si, sj)
    __shared__ double shE[MAX_SLICES*2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& w = reinterpret_cast<WarpScratch*>(smemRaw)[warp];
    if (threadIdx.x < MAX_SUBSETS*MAX_SUBSETS) {
        const int sl = triSlice(threadIdx.x / MAX_SUBSETS, threadIdx.x % MAX_SUBSETS);
        shLam[threadIdx.x] = make_float2(a.lam.c[sl], a.lam.v[sl]);
    }
    if (threadIdx.x < MAX_SLICES*2) shE[threadIdx.x] = 0.0;
    __syncthreads();

    const unsigned below = (1u << lane) - 1u;
    const int nItems = a.counters[2];
    double acc[MAX_SLICES*2];                     // [slice][term], dynamically indexed (local memory, L1-resident)
    if (EMODE != 0) {
#pragma unroll
        for (int k = 0; k < MAX_SLICES*2; k++) acc[k] = 0.0;
    }
    unsigned long long nPairs = 0, hPairs = 0;

    // Everything a warp needs from global memory is requested one step ahead of its use -- the next work
    // item's index while the current item runs, the next tile's atoms while the current tile runs, the list
    // entries two tiles ahead -- because a warp has only a few tiles of work and an L2 round trip per tile
    // would otherwise be fully exposed.
    int nextItem = 0;
    if (lane == 0) nextItem = atomicAdd(a.counters + 3, 1);
    for (;;) {
        const int item = __shfl_sync(FULL_MASK, nextItem, 0);
        if (item >= nItems) break;
        if (lane == 0) nextItem = atomicAdd(a.counters + 3, 1);
        const int4 it = a.items[item];                 // (local block, first tile, first atom, atom count)
        const int lb = it.x;
        const int b = localToGlobalBlock(lb, a.blockPeriod, a.blockOffset, a.blockWidth);
        const int first = it.z, cnt = it.w;
        const int* jl = a.jlist + (size_t) lb*a.capJ;
        const int* xl = a.xlist + (size_t) lb*a.capX;
        const unsigned* xm = a.xmask + (size_t) lb*a.capX;
        const int nJ = a.jcount[lb], nX = a.xcount[lb];
        const uint4 lo = a.blkLo[b];
        const bool iValid = lane < cnt;
        const uint4 pi = iValid ? a.posq[first + lane] : lo;
        const float4 pari = iValid ? a.par[first + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        const double qi64 = (EMODE == 2 && iValid) ? a.q64[first + lane] : 0.0;
        const int tJ = (nJ + 31) >> 5, tX = (nX + 31) >> 5;
        const int tEnd = min(it.y + a.chunkTiles, tJ + tX);
        // list entry (and exclusion mask) of this lane in tile t; -1 beyond the item
        auto loadEntry = [&](int t, unsigned& mask) -> int {
            mask = 0u;
            if (t >= tEnd) return -1;
            if (t >= tJ) { mask = xm[(t - tJ)*32 + lane]; return xl[(t - tJ)*32 + lane]; }
            return jl[t*32 + lane];
        };
        unsigned maskCur, maskNext;
        int entryCur = loadEntry(it.y, maskCur);
        int entryNext = loadEntry(it.y + 1, maskNext);
        uint4 qCur = make_uint4(0u, 0u, 0u, 0u);
        float4 parCur = make_float4(0.f, 0.f, 0.f, 0.f);
        double q64Cur = 0.0;
        if (entryCur >= 0) {
            const int j = entryCur & J_INDEX_MASK;
            qCur = a.posq[j]; parCur = a.par[j];
            if (EMODE == 2) q64Cur = a.q64[j];
        }
        float xi = (float) (pi.x - lo.x)*a.sx;
        const float yi = (float) (pi.y - lo.y)*a.sy, zi = (float) (pi.z - lo.z)*a.sz;
        if (!iValid) xi = 1.0e8f;
        const float qi = iValid ? __uint_as_float(pi.w) : 0.f;
        const float sigi = pari.x, epsi = pari.y;
        const int si = __float_as_int(pari.z);
        const int lamOff = si*MAX_SUBSETS;
        float fix = 0.f, fiy = 0.f, fiz = 0.f;
        __syncwarp();
        w.iPos[lane] = make_float4(xi, yi, zi, qi);
        w.iPar[lane] = pari;
        w.iFix[lane] = make_uint4(pi.x, pi.y, pi.z, (unsigned) __float_as_int(pari.z));
        if (EMODE == 2) w.iQ[lane] = qi64;

        for (int t = it.y; t < tEnd; t++) {
            const bool isX = t >= tJ;
            const int entry = entryCur;
            float4 pj = make_float4(-1.0e8f, 0.f, 0.f, 0.f);
            const float4 parj = parCur;
            uint4 fixj = make_uint4(0u, 0u, 0u, 0u);
            int jIndex = 0;
            if (entry >= 0) {
                jIndex = entry & J_INDEX_MASK;
                const int code = entry >> J_SHIFT_BITS;
                const int kx = code % 5 - 2, ky = (code/5) % 3 - 1, kz = code/15 - 1;
                const uint4 q = qCur;
                // image (kx, ky, kz) is displaced by kx a + ky b + kz c (b and c tilt into x, c into y; zero for a
                // rectangular box), in fixed-point units of each axis
                const long long shx = ((long long) kx << 32) + ky*a.shiftB + kz*a.shiftCx;
                const long long shy = ((long long) ky << 32) + kz*a.shiftCy;
                pj.x = (float) ((long long) q.x + shx - (long long) lo.x)*a.sx;
                pj.y = (float) ((long long) q.y + shy - (long long) lo.y)*a.sy;
                pj.z = (float) ((long long) q.z + ((long long) kz << 32) - (long long) lo.z)*a.sz;
                pj.w = __uint_as_float(q.w);
                // exact coordinates of THIS image modulo the box: the wrapped 32-bit difference to an i atom is then
                // the displacement to this image for any pair inside the cutoff (<= half the box along every axis)
                fixj = make_uint4(q.x + (unsigned) shx, q.y + (unsigned) shy, q.z, (unsigned) __float_as_int(parj.z));
            }
            __syncwarp();
            w.jPos[lane] = pj;
            w.jPar[lane] = parj;
            w.jFix[lane] = fixj;
            if (EMODE == 2) w.jQ[lane] = q64Cur;
            if (isX) w.jMask[lane] = maskCur;
            __syncwarp();
            // requests for the next tile (atoms) and the one after (list entry) go out before this tile's arithmetic
            entryCur = entryNext; maskCur = maskNext;
            qCur = make_uint4(0u, 0u, 0u, 0u); parCur = make_float4(0.f, 0.f, 0.f, 0.f); q64Cur = 0.0;
            if (entryCur >= 0) {
                const int j = entryCur & J_INDEX_MASK;
                qCur = a.posq[j]; parCur = a.par[j];
                if (EMODE == 2) q64Cur = a.q64[j];
            }
            entryNext = loadEntry(t + 2, maskNext);

            if (MODE != 0) {
                // ---- the interacting-pair set itself (parity diagnostics): cull, then hash / dump ----
                int qn = 0;
#pragma unroll 4
                for (int k = 0; k < 32; k++) {
                    const int js = (lane + k) & 31;
                    const bool in = pairInRange(w, a, xi, yi, zi, pi, js, isX, lane);
                    const unsigned m = __ballot_sync(FULL_MASK, in);
                    if (in) w.queue[qn + __popc(m & below)] = (unsigned short) ((lane << 5) | js);
                    qn += __popc(m);
                }
                __syncwarp();
                for (int base = 0; base < qn; base += 32) {
                    if (base + lane < qn) {
                        const unsigned e = w.queue[base + lane];
                        const unsigned oi = (unsigned) __float_as_int(w.iPar[e >> 5].w), oj = (unsigned) __float_as_int(w.jPar[e & 31].w);
                        const unsigned f = min(oi, oj), s = max(oi, oj);
                        nPairs++;
                        hPairs += pairHash(f, s);
                        if (MODE == 2) {
                            unsigned long long slot = atomicAdd(a.pairStats + 2, 1ull);
                            if ((long long) slot < a.dumpCapacity) a.pairDump[slot] = make_int2((int) f, (int) s);
                        }
                    }
                }
                continue;
            }

            // ---- forces: dense 32 x 32 tile, lane l meets j slot (l + k) & 31 at step k ----
            float fjx = 0.f, fjy = 0.f, fjz = 0.f;
            // four copies of the loop so that the exclusion-mask test and the switching function cost nothing
            // in the tiles that do not have them (both conditions are warp-uniform)
            if (isX) {
                if (a.useSwitch) tileLoop<EMODE, CMODE, true, true>(w, a, shLam, lamOff, lane, xi, yi, zi, qi, sigi, epsi, si, pi, fix, fiy, fiz, fjx, fjy, fjz, acc);
                else tileLoop<EMODE, CMODE, true, false>(w, a, shLam, lamOff, lane, xi, yi, zi, qi, sigi, epsi, si, pi, fix, fiy, fiz, fjx, fjy, fjz, acc);
            }
            else {
                if (a.useSwitch) tileLoop<EMODE, CMODE, false, true>(w, a, shLam, lamOff, lane, xi, yi, zi, qi, sigi, epsi, si, pi, fix, fiy, fiz, fjx, fjy, fjz, acc);
                else tileLoop<EMODE, CMODE, false, false>(w, a, shLam, lamOff, lane, xi, yi, zi, qi, sigi, epsi, si, pi, fix, fiy, fiz, fjx, fjy, fjz, acc);
            }
            // j forces of this tile (lane l ends up with slot l) -> global fixed point
            if (entry >= 0 && (fjx != 0.f || fjy != 0.f || fjz != 0.f)) {
                atomicAdd(a.force + jIndex, toFixed(fjx));
                atomicAdd(a.force + a.Npad + jIndex, toFixed(fjy));
                atomicAdd(a.force + 2*(size_t) a.Npad + jIndex, toFixed(fjz));
            }
        }
        // i forces of this item -> global fixed point
        if (MODE == 0) {
            __syncwarp();
            if (iValid) {
                atomicAdd(a.force + first + lane, toFixed(fix));
                atomicAdd(a.force + a.Npad + first + lane, toFixed(fiy));
                atomicAdd(a.force + 2*(size_t) a.Npad + first + lane, toFixed(fiz));
            }
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
    if (EMODE != 0) {
        for (int k = 0; k < a.nE; k++) {
            const double v = warpSum(acc[k]);
            if (lane == 0 && v != 0.0) atomicAdd(&shE[k], v);
        }
        __syncthreads();
        if (threadIdx.x < MAX_SLICES*2 && shE[threadIdx.x] != 0.0) atomicAdd(a.energy + threadIdx.x, shE[threadIdx.x]);
    }
}

template <int EMODE, int CMODE, int MODE, int MINCTAS>
static int launchPairK(Context& c, const PairArgs& a) {
    static bool attr[64] = {false};
    const size_t smem = sizeof(WarpScratch)*PAIR_WARPS;
    if (!attr[c.device & 63]) {
        NBS_CUDA_CHECK(cudaFuncSetAttribute(k_pair<EMODE, CMODE, MODE, MINCTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr[c.device & 63] = true;
    }
    // persistent grid: as many CTAs as are resident at once
    static int perSM[64] = {0};
    if (perSM[c.device & 63] == 0) {
        int n = 0;
        NBS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pair<EMODE, CMODE, MODE, MINCTAS>, PAIR_WARPS*32, smem));
        perSM[c.device & 63] = std::max(1, n);
    }
    k_pair<EMODE, CMODE, MODE, MINCTAS><<<perSM[c.device & 63]*c.numSMs, PAIR_WARPS*32, smem, c.stream>>>(a);
    return NBS_OK;
}

template <int EMODE, int CMODE, int MODE>
static int launchPairT(Context& c, const PairArgs& a) {
    if constexpr (MODE == 0 && CMODE != 2) {
        if (c.N >= 300000) return launchPairK<EMODE, CMODE, MODE, 3>(c, a);
    }
    return launchPairK<EMODE, CMODE, MODE, PAIR_MIN_CTAS>(c, a);
}

template <int CMODE>
static int launchPairE(Context& c, const PairArgs& a, int emode) {
    return emode == 0 ? launchPairT<0, CMODE, 0>(c, a) : (emode == 1 ? launchPairT<1, CMODE, 0>(c, a) : launchPairT<2, CMODE, 0>(c, a));
}

int launchPairs(Context& c, bool wantEnergy, int mode) {
    if (c.blockWidth == 0) return NBS_OK;          // this rank has no direct-space share
    const CellGeom& g = c.geom;
    PairArgs a;
    a.capJ = c.capJ; a.capX = c.capX; a.Npad = c.Npad;
    a.blockPeriod = c.blockPeriod; a.blockOffset = c.blockOffset; a.blockWidth = c.blockWidth;
    a.chunkTiles = c.chunkTiles;
    a.shiftB = g.shiftB; a.shiftCx = g.shiftCx; a.shiftCy = g.shiftCy;
    a.nE = 2*c.nSl;
    a.sx = g.scale[0]; a.sy = g.scale[1]; a.sz = g.scale[2];
    a.dsx = g.box[0]/4294967296.0; a.dsy = g.box[1]/4294967296.0; a.dsz = g.box[2]/4294967296.0;
    // NoCutoff: every pair interacts; the bound only has to exclude the padding lanes (parked at 1e8 nm)
    const bool noCutoff = c.method == NBS_METHOD_NOCUTOFF;
    a.rc2 = noCutoff ? 1.0e12f : (float) (c.cutoff*c.cutoff);
    a.rc2d = noCutoff ? 1.0e12 : c.cutoff*c.cutoff;
    a.alphaD = c.alpha;
    // reaction field only with a cutoff: ReferenceSlicedLJCoulombIxn::setUseCutoff, ReferenceSlicedLJCoulombIxn.cpp:60-68
    a.krfD = noCutoff ? 0.0 : pow(c.cutoff, -3.0)*(c.rfDielectric - 1.0)/(2.0*c.rfDielectric + 1.0);
    a.crfD = noCutoff ? 0.0 : (1.0/c.cutoff)*(3.0*c.rfDielectric)/(2.0*c.rfDielectric + 1.0);
    a.q64 = c.dQ64.d;
    a.erfcTab = c.dErfcTab.d;
    a.alpha = (float) c.alpha;
    a.krf = (float) a.krfD;
    a.crf = (float) a.crfD;
    a.useSwitch = c.useSwitch ? 1 : 0;
    {
        const double dac2 = c.dispAlpha*c.dispAlpha*c.cutoff*c.cutoff, invCut6 = noCutoff ? 0.0 : std::pow(c.cutoff, -6.0);
        a.dalpha2 = (float) (c.dispAlpha*c.dispAlpha);
        a.invCut6 = (float) invCut6;
        a.shiftMult = (float) (invCut6*(1.0 - std::exp(-dac2)*(1.0 + dac2 + 0.5*dac2*dac2)));
    }
    a.rswitch = (float) c.switchDist;
    a.rcut = (float) c.cutoff;
    a.counters = c.dCounters.d;
    a.items = c.dItems.d;
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
    const bool pme = c.ewaldDirect();
    const int emode = !wantEnergy ? 0 : ((c.flags & NBS_FLAG_FP32_ENERGY) ? 1 : 2);
    int status;
    if (mode != 0) {
        NBS_CUDA_CHECK(cudaMemsetAsync(c.dCounters.d + 3, 0, sizeof(int), c.stream));      // rewind the work cursor
        status = mode == 1 ? launchPairT<0, 1, 1>(c, a) : launchPairT<0, 1, 2>(c, a);
    }
    else if (c.ljpme()) status = launchPairE<2>(c, a, emode);
    else if (pme) status = launchPairE<1>(c, a, emode);
    else status = launchPairE<0>(c, a, emode);
    if (status != NBS_OK) return status;
    c.launches++;
    timerMark(c, mode == 0 ? "pair" : "pair_set");
    return NBS_OK;
}

} // namespace nbs

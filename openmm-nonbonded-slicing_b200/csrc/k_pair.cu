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

namespace nbs {


struct PairArgs {
    int capJ, capX, Npad;
    int blockPeriod, blockOffset, blockWidth;   // this rank's share of the i-blocks
    int chunkTiles;                             // tiles per work item
    long long shiftB, shiftCx, shiftCy;         // triclinic image shifts (CellGeom), fixed-point units; 0 for a rectangular box
    unsigned guardX, guardY, guardZ;            // re-used lists: how far (fixed-point units) an i atom may have moved below its block's build-time corner
    int nE;                                     // 2 * number of slices
    float sx, sy, sz;
    double dsx, dsy, dsz;
    float rc2, alpha, krf, crf;
    float rswitch, rcut;
    int useSwitch;
    double rc2d, alphaD, krfD, crfD;
    float dalpha2, invCut6, shiftMult;   // LJPME: alpha_d^2, rc^-6, rc^-6 (1 - exp(-x)(1 + x + x^2/2)) at x = (alpha_d rc)^2
    const double* q64;                   // sorted charges * sqrt(ONE_4PI_EPS0), double
    const double* erfcTab;               // piecewise degree-7 fit of erfc(alpha sqrt(s))/sqrt(s) in s = r^2, [coefficient][tabRows]
    int tabRows;
    int* counters;                       // [2] number of work items, [3] cursor
    const int4* items;                   // (local block, first tile, first atom of the block, atoms in the block)
    const int* blkFirst; const int* blkCount; const uint4* blkLo;
    const uint4* posq; const float4* par;
    const int* jlist; const int* jcount; const int* xlist; const unsigned* xmask; const int* xcount;
    const unsigned* gmJ; const unsigned* gmX;      // cluster masks of the lists' groups of 8 entries, one word per tile
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
    float4 iPar[32];      // sigma/2, 2 sqrt(eps), subset * MAX_SUBSETS, particle index
    float4 jPos[32];      // current tile
    float4 jPar[32];
    // Everything the exact (integer-coordinate) paths gather by slot number -- the double-precision energy passes
    // and the borderline cutoff test -- is stored one 32-bit word per array: 32 words are 32 banks, so a warp's
    // gather with arbitrary slot numbers is conflict-free whatever the pattern.
    unsigned iX[32], iY[32], iZ[32];      // exact fixed-point coordinates
    unsigned jX[32], jY[32], jZ[32];
    unsigned iQlo[32], iQhi[32];          // charge * sqrt(K) in double
    unsigned jQlo[32], jQhi[32];
    float iSig[32], iEps[32];
    float jSig[32], jEps[32];
    // i forces of the work item: cluster c's partial sums of lane l (= 4 jl + il: i atom 4 c + il, as seen by the
    // lane's j slots) -- one private word per (cluster, lane), so the read-modify-write of a step needs no atomics
    float fiX[8][32], fiY[8][32], fiZ[8][32];
    unsigned jMask[32];   // exclusion-list tiles: bit l set = pair (i lane l, this j) is masked
    unsigned short queue[1024 + 32];      // the tile's in-cutoff pairs: subset_i << 13 | subset_j << 10 | i slot << 5 | j slot
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
__device__ __forceinline__ void pairEnergyD(unsigned ix, unsigned iy, unsigned iz, unsigned jx, unsigned jy, unsigned jz,
                                            double qi, double qj, float sigi, float sigj, float epsi, float epsj,
                                            const PairArgs& a, const double* tab, int tabRows, double& ec, double& ev) {
    constexpr bool IS_PME = CMODE != 0;
    const double dx = (double) (int) (jx - ix)*a.dsx;
    const double dy = (double) (int) (jy - iy)*a.dsy;
    const double dz = (double) (int) (jz - iz)*a.dsz;
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
        if (idx >= 0 && idx < tabRows) {
            // d = s 2^(5-e) - (33 + 2 m): position inside the interval, [-1, 1]; the power of two goes straight into
            // the exponent field (e = unbiased exponent of r2f), 33 + 2 m = 2 (16 + m) + 1 from the index bits
            const double scale = __hiloint2double((1023 + 1 + ERFC_TAB_PER_OCTAVE_LOG2 + 127 - (int) (bits >> 23)) << 20, 0);
            const int m2 = 2*(int) ((bits >> (23 - ERFC_TAB_PER_OCTAVE_LOG2)) & ((1u << ERFC_TAB_PER_OCTAVE_LOG2) - 1u)) + (2 << ERFC_TAB_PER_OCTAVE_LOG2) + 1;
            const double d = fma(r2, scale, -(double) m2);
            double p = tab[idx];
#pragma unroll
            for (int k = 1; k <= ERFC_TAB_DEGREE; k++) p = fma(p, d, tab[k*ERFC_TAB_MAX_ROWS + idx]);
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
__device__ __noinline__ bool exactInRange(unsigned ix, unsigned iy, unsigned iz, unsigned jx, unsigned jy, unsigned jz,
                                          double dsx, double dsy, double dsz, double rc2d) {
    const double ex = (double) (int) (jx - ix)*dsx;
    const double ey = (double) (int) (jy - iy)*dsy;
    const double ez = (double) (int) (jz - iz)*dsz;
    return ex*ex + ey*ey + ez*ez <= rc2d;
}

// Lattice translation packed into par.z above the subset by k_reprep (all zero unless the atom has left the brick
// since the neighbour list was built), in 64-bit fixed-point units of each axis.
struct LatticeShift { long long x, y, z; };
__device__ __forceinline__ LatticeShift crossShift(int parz, const PairArgs& a) {
    const int cx = (parz << 25) >> 29, cy = (parz << 23) >> 30, cz = (parz << 21) >> 30;
    LatticeShift s;
    s.x = ((long long) cx << 32) + cy*a.shiftB + cz*a.shiftCx;
    s.y = ((long long) cy << 32) + cz*a.shiftCy;
    s.z = (long long) cz << 32;
    return s;
}

// ---- the tile loop ----------------------------------------------------------------------------------------------
// A tile is 32 list entries (j atoms staged in shared memory) against the item's 32 i atoms.  It is walked as
// 4 GROUPS of 8 entries x up to 8 CLUSTERS of 4 i atoms: lane = 4*jl + il meets j slot 8*g + jl and i slot 4*c + il
// in the step of (group g, cluster c), and the step only exists if the group's cluster mask (from the list
// builder) has bit c -- a warp-uniform test.  The j forces of a group stay in registers over its cluster steps and
// are reduced over the four il lanes once per group; the i forces accumulate in lane-private shared-memory words
// (fiX/Y/Z[c][lane]) -- the cluster index is a run-time value, so ONE copy of the step serves all clusters (eight
// unrolled copies with register accumulators overflowed the instruction cache: 27 % of the kernel's stall samples
// were "no instruction", profiles/r02_ncu_k_pair_a_summary.txt).
struct StepCtx {
    float rc2, alpha;
    float jx, jy, jz, jq, jsig, jeps;     // this lane's j atom of the current group
    int sj;                               // its subset
    unsigned jm;                          // its exclusion mask (exclusion-list tiles)
    int js;                               // its slot in the staged tile
    unsigned qbits;                       // queue entry without the i part: subset_j << 10 | js (the i subset is added per step)
    float fjx, fjy, fjz;
    int qn;                               // pairs waiting in the energy / pair-set queue
};

template <int EMODE, int CMODE, int MODE>
__device__ __forceinline__ void pairStep(WarpScratch& w, const PairArgs& a, const float2* shLam, int lane, int il, int c,
                                         bool isX, StepCtx& s, double* acc) {
    constexpr bool IS_PME = CMODE != 0;
    const float TWO_OVER_SQRT_PI = 1.1283791670955126f;
    const int is = c*4 + il;
    const float4 ip = w.iPos[is];
    const float4 ipar = w.iPar[is];               // sigma/2, 2 sqrt(eps), subset * MAX_SUBSETS (int bits), particle index
    const float dx = ip.x - s.jx, dy = ip.y - s.jy, dz = ip.z - s.jz;
    const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz*dz));
    bool in = r2 <= s.rc2;
    if (fabsf(r2 - s.rc2) < 2.0e-5f) in = exactInRange(w.iX[is], w.iY[is], w.iZ[is], w.jX[s.js], w.jY[s.js], w.jZ[s.js], a.dsx, a.dsy, a.dsz, a.rc2d);     // borderline: rare
    if (isX) in = in && !((s.jm >> is) & 1u);
    if (EMODE == 2 || MODE != 0) {
        const unsigned m = __ballot_sync(FULL_MASK, in);
        if (in) w.queue[s.qn + __popc(m & ((1u << lane) - 1u))] = (unsigned short) (s.qbits | (is << 5) | ((unsigned) __float_as_int(ipar.z) << 10));
        s.qn += __popc(m);
    }
    if (MODE != 0) return;
    const float invR = rsqrtFast(r2);
    const float r = r2*invR;
    const float invR2 = invR*invR;
    float s2 = (ipar.x + s.jsig)*invR;
    s2 *= s2;
    const float s6 = s2*s2*s2;
    const float eps = ipar.y*s.jeps;
    float ev = eps*(s6 - 1.f)*s6;
    float fv = eps*fmaf(12.f, s6, -6.f)*s6*invR2;
    const float qr = ip.w*s.jq*invR;
    float ec, fc;
    if (IS_PME) {
        const float ar = s.alpha*r;
        const float ex = ex2Fast(-1.4426950408889634f*ar*ar);
        const float tt = rcpFast(fmaf(0.5f, ar, 1.f));
        const float erfcv = ex*erfcxPoly(tt);
        ec = qr*erfcv;
        fc = qr*invR2*fmaf(TWO_OVER_SQRT_PI*ar, ex, erfcv);
    }
    else {
        ec = qr*fmaf(a.krf*r2, r, 1.f) - ip.w*s.jq*a.crf;
        fc = qr*invR2*fmaf(-2.f*a.krf*r2, r, 1.f);
    }
    if (CMODE == 2) {
        // LJPME (:398-426): real-space share of the multiplicative C6 term; the shift only enters the energy
        const float sg = ipar.x*s.jsig;
        const float c6 = 64.f*sg*sg*sg*eps;
        const float dar2 = a.dalpha2*r2;
        const float exd = ex2Fast(-1.4426950408889634f*dar2);
        const float p2 = fmaf(dar2, fmaf(0.5f, dar2, 1.f), 1.f);             // 1 + x + x^2/2
        const float c6r6 = c6*invR2*invR2*invR2;
        fv = fmaf(6.f*c6r6*invR2, 1.f - exd*fmaf(dar2*dar2*dar2, 1.f/6.f, p2), fv);
        if (EMODE == 1) {
            float sc = ipar.x + s.jsig;
            sc *= sc;
            const float sc6 = sc*sc*sc*a.invCut6;
            ev += c6r6*(1.f - exd*p2) + eps*(1.f - sc6)*sc6 - c6*a.shiftMult;
        }
    }
    if (a.useSwitch) {                                            // warp-uniform
        if (r > a.rswitch) {
            const float wd = 1.f/(a.rcut - a.rswitch);
            const float u = (r - a.rswitch)*wd;
            const float sv = 1.f + u*u*u*(-10.f + u*(15.f - u*6.f));
            const float sd = u*u*(-30.f + u*(60.f - u*30.f))*wd;
            fv = sv*fv - ev*sd*invR;
            ev *= sv;
        }
    }
    const int siOff = __float_as_int(ipar.z);
    const float2 lam = shLam[siOff + s.sj];
    float dEdR = fmaf(lam.y, fv, lam.x*fc);
    dEdR = in ? dEdR : 0.f;
    w.fiX[c][lane] = fmaf(dEdR, dx, w.fiX[c][lane]);
    w.fiY[c][lane] = fmaf(dEdR, dy, w.fiY[c][lane]);
    w.fiZ[c][lane] = fmaf(dEdR, dz, w.fiZ[c][lane]);
    s.fjx = fmaf(-dEdR, dx, s.fjx); s.fjy = fmaf(-dEdR, dy, s.fjy); s.fjz = fmaf(-dEdR, dz, s.fjz);
    if (EMODE == 1 && in) {
        const int sl = triSlice(siOff/MAX_SUBSETS, s.sj);
        acc[2*sl] += (double) ec;
        acc[2*sl+1] += (double) ev;
    }
}

// Double-precision energies of the queued (in-cutoff) pairs of the current tile, 32 real pairs per pass, two passes
// in flight (their loads and dependent FMA chains interleave).  Operands are gathered word by word from the
// conflict-free per-slot arrays; the erfc table is the CTA's shared-memory copy.  The common case -- every pair of a
// pass in the same slice -- accumulates in two registers; the per-lane table in local memory is only touched when
// the slice changes.
template <int CMODE>
__device__ __forceinline__ void energyPasses(const WarpScratch& w, const PairArgs& a, const double* tab, int lane, int count,
                                             double* acc, int& curSl, double& regC, double& regV) {
    for (int base = 0; base < count; base += 64) {
        double ecd[2] = {0.0, 0.0}, evd[2] = {0.0, 0.0};
        int sl[2] = {-1, -1};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int k = base + 32*u + lane;
            if (k < count) {
                const unsigned e = w.queue[k];
                const int iq = (e >> 5) & 31, jq = e & 31;
                const double qi = __hiloint2double((int) w.iQhi[iq], (int) w.iQlo[iq]);
                const double qj = __hiloint2double((int) w.jQhi[jq], (int) w.jQlo[jq]);
                pairEnergyD<CMODE>(w.iX[iq], w.iY[iq], w.iZ[iq], w.jX[jq], w.jY[jq], w.jZ[jq], qi, qj,
                                   w.iSig[iq], w.jSig[jq], w.iEps[iq], w.jEps[jq], a, tab, a.tabRows, ecd[u], evd[u]);
                sl[u] = triSlice((int) (e >> 13), (int) ((e >> 10) & 7u));
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int lead = __shfl_sync(FULL_MASK, sl[u], 0);      // lane 0 is active whenever the pass has any pair
            const bool uniform = __all_sync(FULL_MASK, sl[u] == lead || sl[u] < 0);
            if (uniform) {
                if (lead < 0) continue;
                if (lead != curSl) {
                    if (curSl >= 0) { acc[2*curSl] += regC; acc[2*curSl+1] += regV; }
                    regC = 0.0; regV = 0.0; curSl = lead;
                }
                regC += ecd[u]; regV += evd[u];
            }
            else if (sl[u] >= 0) { acc[2*sl[u]] += ecd[u]; acc[2*sl[u]+1] += evd[u]; }
        }
    }
}

template <int EMODE, int CMODE, int MODE>
__device__ __forceinline__ void tileLoop(WarpScratch& w, const PairArgs& a, const float2* shLam, int lane, unsigned gmWord, bool isX,
                                         int jIndexMine, bool jValidMine, double* acc, int& qnOut) {
    const int il = lane & 3, jl = lane >> 2;
    StepCtx s;
    s.rc2 = a.rc2; s.alpha = a.alpha;
    s.qn = 0;
#pragma unroll 1
    for (int g = 0; g < 4; g++) {
        const unsigned m = (gmWord >> (8*g)) & 0xffu;
        if (m == 0u) continue;                         // warp-uniform
        s.js = 8*g + jl;
        {
            const float4 p = w.jPos[s.js];
            const float4 pr = w.jPar[s.js];
            s.jx = p.x; s.jy = p.y; s.jz = p.z; s.jq = p.w;
            s.jsig = pr.x; s.jeps = pr.y; s.sj = __float_as_int(pr.z);
            s.jm = isX ? w.jMask[s.js] : 0u;
            s.qbits = ((unsigned) s.sj << 10) | (unsigned) s.js;
        }
        s.fjx = 0.f; s.fjy = 0.f; s.fjz = 0.f;
#pragma unroll 1
        for (unsigned mm = m; mm != 0u; mm &= mm - 1u)
            pairStep<EMODE, CMODE, MODE>(w, a, shLam, lane, il, __ffs((int) mm) - 1, isX, s, acc);
        if (MODE == 0) {
            // j forces of the group: sum over the four il lanes (x and y share the first exchange: odd lanes end
            // up owning y, even lanes x), then lanes il = 0, 1, 2 add x, y, z to the 64-bit fixed-point accumulators
            const bool odd = il & 1;
            float keep = odd ? s.fjy : s.fjx;
            const float send = odd ? s.fjx : s.fjy;
            keep += __shfl_xor_sync(FULL_MASK, send, 1);
            s.fjz += __shfl_xor_sync(FULL_MASK, s.fjz, 1);
            keep += __shfl_xor_sync(FULL_MASK, keep, 2);
            s.fjz += __shfl_xor_sync(FULL_MASK, s.fjz, 2);
            const int jIndex = __shfl_sync(FULL_MASK, jIndexMine, s.js);
            const bool jValid = __shfl_sync(FULL_MASK, (int) jValidMine, s.js) != 0;
            const float v = il == 2 ? s.fjz : keep;
            if (il < 3 && jValid && v != 0.f) atomicAdd(a.force + (size_t) il*a.Npad + jIndex, toFixed(v));
        }
    }
    qnOut = s.qn;
}

// MODE 0: forces (+ energies per EMODE); MODE 1: count + hash the interacting pairs; MODE 2: also dump them.
// EMODE 0: forces only; 1: single-precision energies; 2: double-precision energies.
// MINCTAS: resident CTAs per SM the register allocation is bounded for.
template <int EMODE, int CMODE, int MODE, int MINCTAS>
__global__ void __launch_bounds__(PAIR_WARPS*32, MINCTAS) k_pair(const PairArgs a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ float2 shLam[MAX_SUBSETS*MAX_SUBSETS];      // (lambda_Coulomb, lambda_vdW) of subset pair (si, sj)
    __shared__ double shE[MAX_SLICES*2];
    __shared__ int shArrived;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& w = reinterpret_cast<WarpScratch*>(smemRaw)[warp];
    // EMODE 2: the CTA's copy of the erfc table, behind the warps' scratch areas ([coefficient][ERFC_TAB_MAX_ROWS])
    double* const shTab = reinterpret_cast<double*>(smemRaw + sizeof(WarpScratch)*PAIR_WARPS);
    if (EMODE == 2 && MODE == 0 && CMODE != 0) {
        for (int k = threadIdx.x; k < (ERFC_TAB_DEGREE + 1)*a.tabRows; k += blockDim.x)
            shTab[(k / a.tabRows)*ERFC_TAB_MAX_ROWS + k % a.tabRows] = a.erfcTab[k];
    }
    if (threadIdx.x < MAX_SUBSETS*MAX_SUBSETS) {
        const int sl = triSlice(threadIdx.x / MAX_SUBSETS, threadIdx.x % MAX_SUBSETS);
        shLam[threadIdx.x] = make_float2(a.lam.c[sl], a.lam.v[sl]);
    }
    if (threadIdx.x < MAX_SLICES*2) shE[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) shArrived = 0;
    __syncthreads();

    const int nItems = a.counters[2];
    double acc[MAX_SLICES*2];                     // [slice][term], dynamically indexed (local memory, L1-resident)
    if (EMODE != 0) {
#pragma unroll
        for (int k = 0; k < MAX_SLICES*2; k++) acc[k] = 0.0;
    }
    int curSl = -1;                               // EMODE 2: slice whose energies currently accumulate in registers
    double regC = 0.0, regV = 0.0;
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
        const unsigned* gj = a.gmJ + (size_t) lb*(a.capJ >> 5);
        const unsigned* gx = a.gmX + (size_t) lb*(a.capX >> 5);
        const int nJ = a.jcount[lb], nX = a.xcount[lb];
        const uint4 lo = a.blkLo[b];
        const bool iValid = lane < cnt;
        uint4 pi = iValid ? a.posq[first + lane] : lo;
        const float4 pari = iValid ? a.par[first + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        {
            // an atom that left the brick since the list was built: back next to its build-time position (k_reprep)
            const LatticeShift ls = crossShift(__float_as_int(pari.z), a);
            pi.x += (unsigned) ls.x; pi.y += (unsigned) ls.y;
        }
        const double qi64 = (EMODE == 2 && iValid) ? a.q64[first + lane] : 0.0;
        const int tJ = (nJ + 31) >> 5, tX = (nX + 31) >> 5;
        const int tEnd = min(it.y + a.chunkTiles, tJ + tX);
        // list entry (and exclusion mask) of this lane in tile t; -1 beyond the item
        auto loadEntry = [&](int t, unsigned& mask, unsigned& gm) -> int {
            mask = 0u; gm = 0u;
            if (t >= tEnd) return -1;
            if (t >= tJ) { mask = xm[(t - tJ)*32 + lane]; gm = gx[t - tJ]; return xl[(t - tJ)*32 + lane]; }
            gm = gj[t];
            return jl[t*32 + lane];
        };
        unsigned maskCur, maskNext, gmCur, gmNext;
        int entryCur = loadEntry(it.y, maskCur, gmCur);
        int entryNext = loadEntry(it.y + 1, maskNext, gmNext);
        uint4 qCur = make_uint4(0u, 0u, 0u, 0u);
        float4 parCur = make_float4(0.f, 0.f, 0.f, 0.f);
        double q64Cur = 0.0;
        if (entryCur >= 0) {
            const int j = entryCur & J_INDEX_MASK;
            qCur = a.posq[j]; parCur = a.par[j];
            if (EMODE == 2) q64Cur = a.q64[j];
        }
        // (window arithmetic: the atom sits in [corner - guard, corner + box - guard) whichever side of the brick's
        // faces it and the corner are on; guard = 0 unless the list is being re-used)
        float xi = (float) ((long long) (pi.x - lo.x + a.guardX) - (long long) a.guardX)*a.sx;
        const float yi = (float) ((long long) (pi.y - lo.y + a.guardY) - (long long) a.guardY)*a.sy;
        const float zi = (float) ((long long) (pi.z - lo.z + a.guardZ) - (long long) a.guardZ)*a.sz;
        if (!iValid) xi = 1.0e8f;
        const float qi = iValid ? __uint_as_float(pi.w) : 0.f;
        const int si = __float_as_int(pari.z) & 7;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; c++) { w.fiX[c][lane] = 0.f; w.fiY[c][lane] = 0.f; w.fiZ[c][lane] = 0.f; }
        w.iPos[lane] = make_float4(xi, yi, zi, qi);
        w.iPar[lane] = make_float4(pari.x, pari.y, __int_as_float(si*MAX_SUBSETS), pari.w);
        w.iX[lane] = pi.x; w.iY[lane] = pi.y; w.iZ[lane] = pi.z;
        if (EMODE == 2) {
            w.iQlo[lane] = (unsigned) __double2loint(qi64); w.iQhi[lane] = (unsigned) __double2hiint(qi64);
            w.iSig[lane] = pari.x; w.iEps[lane] = pari.y;
        }

        for (int t = it.y; t < tEnd; t++) {
            const bool isX = t >= tJ;
            const int entry = entryCur;
            const unsigned gmWord = gmCur;
            float4 pj = make_float4(-1.0e8f, 0.f, 0.f, 0.f);
            const float4 parj = parCur;
            unsigned fjx_ = 0u, fjy_ = 0u, fjz_ = 0u;
            int jIndex = 0;
            if (entry >= 0) {
                jIndex = entry & J_INDEX_MASK;
                const int code = entry >> J_SHIFT_BITS;
                const LatticeShift ls = crossShift(__float_as_int(parj.z), a);
                const int kz = code/15 - 1;
                const uint4 q = qCur;
                // image (kx, ky, kz) is displaced by kx a + ky b + kz c (b and c tilt into x, c into y; zero for a
                // rectangular box), in fixed-point units of each axis
                // (+ the lattice translation of an atom that left the brick since the list was built)
                const long long shx = ((long long) (code % 5 - 2) << 32) + ((code/5) % 3 - 1)*a.shiftB + kz*a.shiftCx + ls.x;
                const long long shy = ((long long) ((code/5) % 3 - 1) << 32) + kz*a.shiftCy + ls.y;
                pj.x = (float) ((long long) q.x + shx - (long long) lo.x)*a.sx;
                pj.y = (float) ((long long) q.y + shy - (long long) lo.y)*a.sy;
                pj.z = (float) ((long long) q.z + ((long long) kz << 32) + ls.z - (long long) lo.z)*a.sz;
                pj.w = __uint_as_float(q.w);
                // exact coordinates of THIS image modulo the box: the wrapped 32-bit difference to an i atom is then
                // the displacement to this image for any pair inside the cutoff (<= half the box along every axis)
                fjx_ = q.x + (unsigned) shx; fjy_ = q.y + (unsigned) shy; fjz_ = q.z;
            }
            __syncwarp();
            w.jPos[lane] = pj;
            w.jPar[lane] = make_float4(parj.x, parj.y, __int_as_float(__float_as_int(parj.z) & 7), parj.w);
            w.jX[lane] = fjx_; w.jY[lane] = fjy_; w.jZ[lane] = fjz_;
            if (EMODE == 2) {
                w.jQlo[lane] = (unsigned) __double2loint(q64Cur); w.jQhi[lane] = (unsigned) __double2hiint(q64Cur);
                w.jSig[lane] = parj.x; w.jEps[lane] = parj.y;
            }
            if (isX) w.jMask[lane] = maskCur;
            __syncwarp();
            // requests for the next tile (atoms) and the one after (list entry) go out before this tile's arithmetic
            entryCur = entryNext; maskCur = maskNext; gmCur = gmNext;
            qCur = make_uint4(0u, 0u, 0u, 0u); parCur = make_float4(0.f, 0.f, 0.f, 0.f); q64Cur = 0.0;
            if (entryCur >= 0) {
                const int j = entryCur & J_INDEX_MASK;
                qCur = a.posq[j]; parCur = a.par[j];
                if (EMODE == 2) q64Cur = a.q64[j];
            }
            entryNext = loadEntry(t + 2, maskNext, gmNext);

            int qn = 0;
            tileLoop<EMODE, CMODE, MODE>(w, a, shLam, lane, gmWord, isX, jIndex, entry >= 0, acc, qn);
            if (MODE != 0) {
                // ---- the interacting-pair set itself (parity diagnostics): hash / dump what the loop queued ----
                __syncwarp();
                for (int base = 0; base < qn; base += 32) {
                    if (base + lane < qn) {
                        const unsigned e = w.queue[base + lane];
                        const unsigned oi = (unsigned) __float_as_int(w.iPar[(e >> 5) & 31].w), oj = (unsigned) __float_as_int(w.jPar[e & 31].w);
                        const unsigned f = min(oi, oj), sd = max(oi, oj);
                        nPairs++;
                        hPairs += pairHash(f, sd);
                        if (MODE == 2) {
                            unsigned long long slot = atomicAdd(a.pairStats + 2, 1ull);
                            if ((long long) slot < a.dumpCapacity) a.pairDump[slot] = make_int2((int) f, (int) sd);
                        }
                    }
                }
            }
            else if (EMODE == 2) {                 // the queue refers to this tile's shared-memory slots
                __syncwarp();
                energyPasses<CMODE>(w, a, shTab, lane, qn, acc, curSl, regC, regV);
            }
        }
        // i forces of this item: fi[c] summed over the eight jl lanes.  Three exchange stages, each halving the
        // number of clusters a lane still carries, leave lane 4 c + il with the total of cluster c, atom il --
        // i.e. lane l with the force on the block's atom l.
        if (MODE == 0) {
            __syncwarp();
#pragma unroll
            for (int d = 0; d < 3; d++) {
                float v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = d == 0 ? w.fiX[c][lane] : (d == 1 ? w.fiY[c][lane] : w.fiZ[c][lane]);
#pragma unroll
                for (int st = 2; st >= 0; st--) {
                    const int half = 1 << st;
                    const bool upper = (lane >> (st + 2)) & 1;
#pragma unroll
                    for (int k = 0; k < half; k++) {
                        const float send = upper ? v[k] : v[k + half];
                        const float keep = upper ? v[k + half] : v[k];
                        v[k] = keep + __shfl_xor_sync(FULL_MASK, send, 4 << st);
                    }
                }
                if (iValid && v[0] != 0.f) atomicAdd(a.force + (size_t) d*a.Npad + first + lane, toFixed(v[0]));
            }
            __syncwarp();
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
        if (EMODE == 2 && curSl >= 0) { acc[2*curSl] += regC; acc[2*curSl+1] += regV; }
        // CTA-level sum without a barrier (a warp that has run out of work items retires; the last one to arrive
        // adds the CTA's totals to the global table)
        for (int k = 0; k < a.nE; k++) {
            const double v = warpSum(acc[k]);
            if (lane == 0 && v != 0.0) atomicAdd(&shE[k], v);
        }
        __threadfence_block();
        int arrived = 0;
        if (lane == 0) arrived = atomicAdd(&shArrived, 1);
        arrived = __shfl_sync(FULL_MASK, arrived, 0);
        if (arrived == PAIR_WARPS - 1) {
            __threadfence_block();
            for (int k = lane; k < a.nE; k += 32) {
                const double v = *((volatile double*) &shE[k]);
                if (v != 0.0) atomicAdd(a.energy + k, v);
            }
        }
    }
}

template <int EMODE, int CMODE, int MODE, int MINCTAS>
static int launchPairK(Context& c, const PairArgs& a) {
    static bool attr[64] = {false};
    const size_t smem = sizeof(WarpScratch)*PAIR_WARPS + (EMODE == 2 && MODE == 0 && CMODE != 0 ? sizeof(double)*(ERFC_TAB_DEGREE + 1)*ERFC_TAB_MAX_ROWS : 0);
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
        static const int forced = getenv("NBS_PAIR_MINCTAS") ? atoi(getenv("NBS_PAIR_MINCTAS")) : 0;     // tuning experiments
        if (forced == 3 || (forced == 0 && c.N >= 300000)) return launchPairK<EMODE, CMODE, MODE, 3>(c, a);
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
    {
        const double room = c.skin > 0 ? 0.5*c.skin + 1.0e-3 : 0.0;      // nm an atom may have moved since the build
        a.guardX = (unsigned) (room/g.box[0]*4294967296.0); a.guardY = (unsigned) (room/g.box[1]*4294967296.0);
        a.guardZ = (unsigned) (room/g.box[2]*4294967296.0);
    }
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
    a.tabRows = c.erfcRows;
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
    a.gmJ = c.dGmJ.d; a.gmX = c.dGmX.d;
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

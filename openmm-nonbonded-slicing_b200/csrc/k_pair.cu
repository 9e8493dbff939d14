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
    double ds2;                                 // dsx^2 when the three units are equal (cubic box: r^2 in integer arithmetic)
    float rc2, alpha, krf, crf;
    float rswitch, rcut;
    int useSwitch;
    double rc2d, alphaD, krfD, crfD;
    float dalpha2, invCut6, shiftMult;   // LJPME: alpha_d^2, rc^-6, rc^-6 (1 - exp(-x)(1 + x + x^2/2)) at x = (alpha_d rc)^2
    double dalpha2D, invCut6D, shiftMultD, rswitchD, rcutD;     // the same (and the switching range) for the double-precision kernel
    const double2* sigEpsD;              // double-precision kernel: (sigma/2, 2 sqrt(eps)) by PARTICLE index
    const double* lamD;                  // double-precision kernel: [nSl][2] (lambda_Coulomb, lambda_vdW)
    const double* q64;                   // sorted charges * sqrt(ONE_4PI_EPS0), double
    const double* erfcTab;               // piecewise fit of erfc(alpha sqrt(s))/sqrt(s) in s = r^2: c0[tabRows], float4 (c1..c4)[tabRows]
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
    float4 jPar[32];      // sigma/2, 2 sqrt(eps), subset, particle index
    // exact fixed-point coordinates and double-precision charges (charge * sqrt(K)), one 32-bit word per array:
    // 32 words are 32 banks, so any gather by slot number is conflict-free
    unsigned iX[32], iY[32], iZ[32], iQlo[32], iQhi[32];
    unsigned jX[32], jY[32], jZ[32], jQlo[32], jQhi[32];
    unsigned jMask[32];   // exclusion-list tiles: bit l set = pair (i lane l, this j) is masked
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

// erfc(alpha r)/r in double WITHOUT the table, for the pairs the table does not cover (closer than 0.088 nm -- none
// in a physical system -- or beyond its last row): rsqrt + Newton, exp(-x^2) as 2^n e^h with a degree-11 Taylor
// polynomial, erfcx as a degree-14 polynomial.  Out of line: it is never on the hot path.
__device__ __noinline__ double erfcOverRAnalytic(double r2, double alphaD) {
    double y = (double) rsqrtFast((float) r2);
    y = y*fma(-0.5*r2*y, y, 1.5);
    y = y*fma(-0.5*r2*y, y, 1.5);
    const double x = alphaD*r2*y;
    const double dd = fma(0.5, x, 1.0);
    double t = (double) rcpFast((float) dd);
    t = t*fma(-dd, t, 2.0);
    t = t*fma(-dd, t, 2.0);
    const double z = x*x;
    const double u = -z*kMiscD[2];                              // log2(e)
    const double n = (u + kMiscD[3]) - kMiscD[3];               // 1.5 * 2^52: rounds to the nearest integer
    const double h = fma(n, kMiscD[5], fma(n, kMiscD[4], -z));  // -z - n ln2, ln2 in two parts
    double e = kExpD[0];
#pragma unroll
    for (int k = 1; k < 12; k++) e = fma(e, h, kExpD[k]);
    e = __hiloint2double(__double2hiint(e) + ((int) n << 20), __double2loint(e));
    const double v = fma(t, kMiscD[0], kMiscD[1]);
    double p = kErfcxD[0];
#pragma unroll
    for (int k = 1; k < 15; k++) p = fma(p, v, kErfcxD[k]);
    return y*e*p;
}

// Exact cutoff test from the fixed-point coordinates (the wrapped integer difference is the minimum image
// for any pair near the cutoff, because the box is at least twice the cutoff).  Out of line: only pairs whose
// fp32 r^2 lands within 2e-5 nm^2 of the cutoff get here.
#ifdef PAIR_EXACT_INLINE
__device__ __forceinline__ bool exactInRange(
#else
__device__ __noinline__ bool exactInRange(
#endif
unsigned ix, unsigned iy, unsigned iz, unsigned jx, unsigned jy, unsigned jz,
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
// are reduced over the four il lanes once per group; the i forces of all 8 clusters stay in registers (fi[c]) for
// the whole work item.
//
// Energies are evaluated IN the step, for every lane, and counted for the lanes whose pair is inside the cutoff:
//   * Coulomb (EMODE 2): K q_i q_j erfc(alpha r)/r in DOUBLE from the exact fixed-point coordinates -- slice energies
//     are sums of 10^4..10^8 terms of both signs and single precision cannot deliver 1e-5 of a small net value
//     (DESIGN.md "Precision").  r^2 is formed in 64-bit INTEGER arithmetic when the box is cubic (one conversion to
//     double; conversions are the scarce resource: 0.45 warp instructions per clock and SM against 1.6 for DFMA,
//     profiles/r02_peaks.json), f(s) = erfc(alpha sqrt(s))/sqrt(s) comes from the CTA's shared-memory table: per row
//     (128 per octave of s; row and position inside it straight from the bits of the double) c0 in double and a
//     degree-4 remainder in SINGLE precision -- the remainder is at most 3 % of f, so its single-precision rounding is
//     3e-9 of f (ERFC_TAB_* in nbs_internal.h).  qq c0 is summed in double, qq (f - c0) in single precision per tile:
//     two double-precision instructions and two shared-memory loads per pair (the first table was a degree-7 double
//     Horner chain: eight loads, twelve double-precision instructions, 60 % of the kernel's shared-memory wavefronts).
//     With the exact r^2 at hand the cutoff decision is exact too: no borderline branch.
//   * Lennard-Jones: the step's fp32 value, summed in fp32 over a tile and in double across tiles.
//   * a tile whose i atoms share one subset and whose j atoms share one subset -- almost all of them -- adds into two
//     registers; mixed tiles go through the per-lane slice table in local memory.
struct StepCtx {
    float rc2, alpha;
    float jx, jy, jz, jq, jsig, jeps;     // this lane's j atom of the current group
    int sj;                               // its subset
    unsigned jm;                          // its exclusion mask (exclusion-list tiles)
    int js;                               // its slot in the staged tile
    unsigned jxe, jye, jze;               // EMODE 2: its exact coordinates and double-precision charge
    double jq64;
    unsigned jOrig;                       // MODE 1/2: its particle index
    float fjx, fjy, fjz;
    // energy accumulators of the current tile
    double ecTile;                        // EMODE 2, uniform tile
    float ecTileF, evTile;                // EMODE 1 Coulomb; Lennard-Jones (both modes), uniform tile
    bool uniform;                         // the tile's pairs all belong to one slice
    unsigned long long nPairs, hPairs;    // MODE 1/2
};

template <int EMODE, int CMODE, int MODE, bool CUBIC, int C>
__device__ __forceinline__ void pairStep(WarpScratch& w, const PairArgs& a, const float2* shLam, const double* tab, int lane, int il,
                                         bool isX, StepCtx& s, float (&fi)[8][3], double* acc) {
    constexpr bool IS_PME = CMODE != 0;
    const float TWO_OVER_SQRT_PI = 1.1283791670955126f;
    const int is = C*4 + il;
    const float4 ip = w.iPos[is];
    const float4 ipar = w.iPar[is];               // sigma/2, 2 sqrt(eps), subset * MAX_SUBSETS (int bits), particle index
    const float dx = ip.x - s.jx, dy = ip.y - s.jy, dz = ip.z - s.jz;
    const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz*dz));
    bool in;
    double r2d = 0.0;
    if (EMODE == 2 && MODE == 0) {
        // exact r^2 (needed for the energy anyway) decides; the fp32 value only gates out the parked padding lanes
        const int ex = (int) (s.jxe - w.iX[is]), ey = (int) (s.jye - w.iY[is]), ez = (int) (s.jze - w.iZ[is]);
        if (CUBIC) {
            const unsigned long long r2i = (unsigned long long) ((long long) ex*ex) + (unsigned long long) ((long long) ey*ey) +
                                           (unsigned long long) ((long long) ez*ez);
            r2d = __ull2double_rn(r2i)*a.ds2;
        }
        else {
            const double fx = (double) ex*a.dsx, fy = (double) ey*a.dsy, fz = (double) ez*a.dsz;
            r2d = fma(fx, fx, fma(fy, fy, fz*fz));
        }
        in = r2 < s.rc2 + 2.0e-5f && r2d <= a.rc2d;
    }
    else {
        in = r2 <= s.rc2;
        if (fabsf(r2 - s.rc2) < 2.0e-5f)          // borderline: rare
            in = exactInRange(w.iX[is], w.iY[is], w.iZ[is], w.jX[s.js], w.jY[s.js], w.jZ[s.js], a.dsx, a.dsy, a.dsz, a.rc2d);
    }
    if (isX) in = in && !((s.jm >> is) & 1u);
    if (MODE != 0) {
        // the interacting-pair set itself (parity diagnostics): count + hash (+ dump)
        if (in) {
            const unsigned oi = (unsigned) __float_as_int(ipar.w);
            const unsigned f = min(oi, s.jOrig), sd = max(oi, s.jOrig);
            s.nPairs++;
            s.hPairs += pairHash(f, sd);
            if (MODE == 2) {
                const unsigned long long slot = atomicAdd(a.pairStats + 2, 1ull);
                if ((long long) slot < a.dumpCapacity) a.pairDump[slot] = make_int2((int) f, (int) sd);
            }
        }
        return;
    }
    const float invR = rsqrtFast(r2);
    const float r = r2*invR;
    const float invR2 = invR*invR;
    float s2 = (ipar.x + s.jsig)*invR;
    s2 *= s2;
    const float s6 = s2*s2*s2;
    const float eps = ipar.y*s.jeps;
    float ev = eps*(s6 - 1.f)*s6;
    float fv = eps*fmaf(12.f, s6, -6.f)*s6*invR2;
    float invE2 = invR2;                          // 1/r^2 for the energy terms
    if (EMODE == 2) {
        // Lennard-Jones ENERGY at the exact r^2: the force path evaluated E at the distance rho with 1/rho = invR, which
        // differs from the exact one by the approximation error of rsqrt and the ~6e-7 of the tile-relative fp32
        // coordinates (r^-12 would turn that into 4e-6 per close contact: visible in small slices such as
        // protein-ligand).  First order in the difference: E(r^2) = E(rho^2) + dE/d(r^2) (r^2 - rho^2), with
        // dE/d(r^2) = -fv/2 and r^2 - rho^2 = -(1 - r^2 invR^2)/invR^2 -- three instructions instead of a second evaluation
        // (the second-order term is 1e-12 of E).
        const float r2e = (float) r2d;
        const float t = fmaf(-r2e, invR2, 1.f);
        ev = fmaf(0.5f*eps*fmaf(12.f, s6, -6.f)*s6, t, ev);
        if (CMODE == 2) invE2 = invR2*(1.f + t);  // LJPME's extra r^-6 term below, to the same order
    }
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
        if (EMODE != 0) {
            float sc = ipar.x + s.jsig;
            sc *= sc;
            const float sc6 = sc*sc*sc*a.invCut6;
            ev += c6*invE2*invE2*invE2*(1.f - exd*p2) + eps*(1.f - sc6)*sc6 - c6*a.shiftMult;
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
    fi[C][0] = fmaf(dEdR, dx, fi[C][0]); fi[C][1] = fmaf(dEdR, dy, fi[C][1]); fi[C][2] = fmaf(dEdR, dz, fi[C][2]);
    s.fjx = fmaf(-dEdR, dx, s.fjx); s.fjy = fmaf(-dEdR, dy, s.fjy); s.fjz = fmaf(-dEdR, dz, s.fjz);
    if (EMODE == 0) return;

    // ---- energies ----
    // EMODE 2: K q_i q_j f(r^2) = qq c0 (double) + qq (f - c0) (single: the remainder is a few per cent of f)
    double ecd = 0.0;
    float ecRem = 0.f;
    if (EMODE == 2) {
        const double qq = __hiloint2double((int) w.iQhi[is], (int) w.iQlo[is])*s.jq64;
        if (IS_PME) {
            // table row and position inside its interval straight from the bits of r^2: exponent and top seven mantissa
            // bits are the row, the next 23 mantissa bits the position d in [-1/2, 1/2) (centred: + half a unit)
            constexpr int L = ERFC_TAB_PER_OCTAVE_LOG2;
            const int hi = __double2hiint(r2d);
            const unsigned lo = (unsigned) __double2loint(r2d);
            const int idx = (hi >> (20 - L)) - ((1023 - 7) << L);
            // (a pair beyond the table -- far outside the cutoff, or closer than 0.088 nm -- reads the nearest row; its
            // value is only used, and then replaced by the analytic form, if the pair is inside the cutoff)
            const int row = min(max(idx, 0), a.tabRows - 1);
            const unsigned frac = (((unsigned) hi & ((1u << (20 - L)) - 1u)) << (3 + L)) | (lo >> (29 - L));
            const float d = (__uint_as_float(0x3f800000u | frac) - 1.5f) + 5.9604645e-8f;
            double c0 = tab[row];
            const float4 cf = reinterpret_cast<const float4*>(tab + ERFC_TAB_MAX_ROWS)[row];
            ecRem = d*fmaf(d, fmaf(d, fmaf(d, cf.w, cf.z), cf.y), cf.x);
            if (in && row != idx) { c0 = erfcOverRAnalytic(r2d, a.alphaD); ecRem = 0.f; }           // rare
            ecd = qq*c0;
            ecRem *= ip.w*s.jq;
        }
        else {
            double y = (double) invR;                              // reaction field (:598-624): 1/r + krf r^2 - crf
            y = y*fma(-0.5*r2d*y, y, 1.5);
            y = y*fma(-0.5*r2d*y, y, 1.5);
            ecd = qq*(y + a.krfD*r2d - a.crfD);
        }
    }
    if (s.uniform) {
        if (EMODE == 2) { s.ecTile += in ? ecd : 0.0; s.ecTileF += in ? ecRem : 0.f; }
        else s.ecTileF += in ? ec : 0.f;
        s.evTile += in ? ev : 0.f;
    }
    else if (in) {
        const int sl = triSlice(siOff/MAX_SUBSETS, s.sj);
        acc[2*sl] += EMODE == 2 ? ecd + (double) ecRem : (double) ec;
        acc[2*sl+1] += (double) ev;
    }
}

template <int EMODE, int CMODE, int MODE, bool CUBIC>
__device__ __forceinline__ void tileLoop(WarpScratch& w, const PairArgs& a, const float2* shLam, const double* tab, int lane, unsigned gmWord,
                                         bool isX, int jIndexMine, bool jValidMine, StepCtx& s, float (&fi)[8][3], double* acc) {
    const int il = lane & 3, jl = lane >> 2;
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
            if (MODE != 0) s.jOrig = (unsigned) __float_as_int(pr.w);
            if (EMODE == 2 && MODE == 0) {
                s.jxe = w.jX[s.js]; s.jye = w.jY[s.js]; s.jze = w.jZ[s.js];
                s.jq64 = __hiloint2double((int) w.jQhi[s.js], (int) w.jQlo[s.js]);
            }
        }
        s.fjx = 0.f; s.fjy = 0.f; s.fjz = 0.f;
        if (m & 0x01u) pairStep<EMODE, CMODE, MODE, CUBIC, 0>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
        if (m & 0x02u) pairStep<EMODE, CMODE, MODE, CUBIC, 1>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
        if (m & 0x04u) pairStep<EMODE, CMODE, MODE, CUBIC, 2>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
        if (m & 0x08u) pairStep<EMODE, CMODE, MODE, CUBIC, 3>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
        if (m & 0x10u) pairStep<EMODE, CMODE, MODE, CUBIC, 4>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
        if (m & 0x20u) pairStep<EMODE, CMODE, MODE, CUBIC, 5>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
        if (m & 0x40u) pairStep<EMODE, CMODE, MODE, CUBIC, 6>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
        if (m & 0x80u) pairStep<EMODE, CMODE, MODE, CUBIC, 7>(w, a, shLam, tab, lane, il, isX, s, fi, acc);
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
}

// MODE 0: forces (+ energies per EMODE); MODE 1: count + hash the interacting pairs; MODE 2: also dump them.
// EMODE 0: forces only; 1: single-precision energies; 2: double-precision Coulomb energies.
// CUBIC: the three fixed-point units are equal (r^2 in integer arithmetic).
// MINCTAS: resident CTAs per SM the register allocation is bounded for.
template <int EMODE, int CMODE, int MODE, bool CUBIC, int MINCTAS>
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
        // (c0[tabRows] doubles, then the rows' float4 coefficient sets as pairs of 8-byte words)
        for (int k = threadIdx.x; k < 3*a.tabRows; k += blockDim.x)
            shTab[k < a.tabRows ? k : ERFC_TAB_MAX_ROWS + (k - a.tabRows)] = a.erfcTab[k];
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
    int curSl = -1;                               // slice whose energies currently accumulate in registers
    double regC = 0.0, regV = 0.0;
    StepCtx s;
    s.rc2 = a.rc2; s.alpha = a.alpha;
    s.nPairs = 0; s.hPairs = 0;
    s.ecTile = 0.0; s.ecTileF = 0.f; s.evTile = 0.f; s.uniform = false;
    s.jm = 0u; s.jOrig = 0u; s.jxe = s.jye = s.jze = 0u; s.jq64 = 0.0;

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
        // do the i atoms share one subset?  (then a tile whose j atoms do too accumulates its energies in registers)
        const int si0 = __shfl_sync(FULL_MASK, si, 0);
        const bool iUniform = __all_sync(FULL_MASK, si == si0 || !iValid);
        float fi[8][3];
#pragma unroll
        for (int c = 0; c < 8; c++) { fi[c][0] = 0.f; fi[c][1] = 0.f; fi[c][2] = 0.f; }
        __syncwarp();
        w.iPos[lane] = make_float4(xi, yi, zi, qi);
        w.iPar[lane] = make_float4(pari.x, pari.y, __int_as_float(si*MAX_SUBSETS), pari.w);
        w.iX[lane] = pi.x; w.iY[lane] = pi.y; w.iZ[lane] = pi.z;
        if (EMODE == 2) { w.iQlo[lane] = (unsigned) __double2loint(qi64); w.iQhi[lane] = (unsigned) __double2hiint(qi64); }

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
            const int sjMine = __float_as_int(parj.z) & 7;
            __syncwarp();
            w.jPos[lane] = pj;
            w.jPar[lane] = make_float4(parj.x, parj.y, __int_as_float(sjMine), parj.w);
            w.jX[lane] = fjx_; w.jY[lane] = fjy_; w.jZ[lane] = fjz_;
            if (EMODE == 2) { w.jQlo[lane] = (unsigned) __double2loint(q64Cur); w.jQhi[lane] = (unsigned) __double2hiint(q64Cur); }
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

            if (EMODE != 0 && MODE == 0) {
                // one slice for the whole tile?  (lane 0's entry of a tile is never padding)
                const int sj0 = __shfl_sync(FULL_MASK, sjMine, 0);
                s.uniform = iUniform && __all_sync(FULL_MASK, sjMine == sj0 || entry < 0);
                if (s.uniform) {
                    const int sl = triSlice(si0, sj0);
                    if (sl != curSl) {
                        if (curSl >= 0) { acc[2*curSl] += regC; acc[2*curSl+1] += regV; }
                        regC = 0.0; regV = 0.0; curSl = sl;
                    }
                }
                s.ecTile = 0.0; s.ecTileF = 0.f; s.evTile = 0.f;
            }
            tileLoop<EMODE, CMODE, MODE, CUBIC>(w, a, shLam, shTab, lane, gmWord, isX, jIndex, entry >= 0, s, fi, acc);
            if (EMODE != 0 && MODE == 0 && s.uniform) {
                regC += EMODE == 2 ? s.ecTile + (double) s.ecTileF : (double) s.ecTileF;
                regV += (double) s.evTile;
            }
        }
        // i forces of this item: fi[c] summed over the eight jl lanes.  Three exchange stages, each halving the
        // number of clusters a lane still carries, leave lane 4 c + il with the total of cluster c, atom il --
        // i.e. lane l with the force on the block's atom l.
        if (MODE == 0) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                float v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = fi[c][d];
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
        unsigned long long nPairs = (unsigned long long) warpSum((double) s.nPairs), hPairs = s.hPairs;        // exact below 2^53
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hPairs += __shfl_xor_sync(FULL_MASK, hPairs, o);
        if (lane == 0) {
            atomicAdd(a.pairStats, nPairs);
            atomicAdd(a.pairStats + 1, hPairs);
        }
        return;
    }
    if (EMODE != 0) {
        if (curSl >= 0) { acc[2*curSl] += regC; acc[2*curSl+1] += regV; }
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

// ---- double-precision force mode (NBS_FLAG_DOUBLE) ----------------------------------------------------------------
// The plugin's Precision = double (CommonNonbondedSlicingKernels.cpp:297-299).  Same neighbour list and work items; the
// arithmetic of ReferenceSlicedLJCoulombIxn.cpp:367-445 (:571-631 for the reaction field) in double precision from the
// exact fixed-point coordinates, libm-grade erfc / exp.  Not tuned: one lane per i atom, lane l meets j slot (l + k) mod 32
// at step k and the j accumulators rotate one lane per step.
struct __align__(16) WarpScratchD {
    long long jX[32], jY[32], jZ[32];          // relative to the i-block's corner, fixed-point units
    unsigned jMask[32];
    double jQ[32], jSig[32], jEps[32];
    int jSub[32], jIndex[32];                  // jIndex < 0: padding
};

template <int CMODE>
__global__ void __launch_bounds__(256) k_pair_f64(const PairArgs a) {
    __shared__ WarpScratchD scratch[8];
    __shared__ double shE[MAX_SLICES*2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratchD& w = scratch[warp];
    if (threadIdx.x < MAX_SLICES*2) shE[threadIdx.x] = 0.0;
    __syncthreads();
    const int nItems = a.counters[2];
    double acc[MAX_SLICES*2];
#pragma unroll
    for (int k = 0; k < MAX_SLICES*2; k++) acc[k] = 0.0;
    const double TWO_OVER_SQRT_PI = 1.1283791670955125739;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(a.counters + 3, 1);
        item = __shfl_sync(FULL_MASK, item, 0);
        if (item >= nItems) break;
        const int4 it = a.items[item];
        const int lb = it.x, first = it.z, cnt = it.w;
        const int* jl = a.jlist + (size_t) lb*a.capJ;
        const int* xl = a.xlist + (size_t) lb*a.capX;
        const unsigned* xm = a.xmask + (size_t) lb*a.capX;
        const int nJ = a.jcount[lb], nX = a.xcount[lb];
        const int tJ = (nJ + 31) >> 5, tX = (nX + 31) >> 5;
        const int tEnd = min(it.y + a.chunkTiles, tJ + tX);
        const bool iValid = lane < cnt;
        const uint4 lo = a.blkLo[localToGlobalBlock(lb, a.blockPeriod, a.blockOffset, a.blockWidth)];
        uint4 pi = lo;
        double qi = 0.0, sigi = 0.0, epsi = 0.0;
        int si = 0;
        // coordinates relative to the block's corner, in 64-bit fixed-point units: the i atom in the window [corner - guard,
        // corner + box - guard), a list entry at the periodic image its image code names -- exactly the geometry of the
        // single-precision kernel, without its conversion to float (a plain wrapped 32-bit difference would pick the
        // NEAREST image of j, which in a small box can be a second entry of the same list)
        long long wix = 0, wiy = 0, wiz = 0;
        if (iValid) {
            pi = a.posq[first + lane];
            const float4 pr = a.par[first + lane];
            const LatticeShift ls = crossShift(__float_as_int(pr.z), a);
            pi.x += (unsigned) ls.x; pi.y += (unsigned) ls.y;
            wix = (long long) (pi.x - lo.x + a.guardX) - (long long) a.guardX;
            wiy = (long long) (pi.y - lo.y + a.guardY) - (long long) a.guardY;
            wiz = (long long) (pi.z - lo.z + a.guardZ) - (long long) a.guardZ;
            qi = a.q64[first + lane];
            const double2 se = a.sigEpsD[__float_as_int(pr.w)];
            sigi = se.x; epsi = se.y;
            si = __float_as_int(pr.z) & 7;
        }
        double fix = 0.0, fiy = 0.0, fiz = 0.0;
        for (int t = it.y; t < tEnd; t++) {
            const bool isX = t >= tJ;
            const int entry = isX ? xl[(t - tJ)*32 + lane] : jl[t*32 + lane];
            __syncwarp();
            w.jIndex[lane] = -1;
            w.jMask[lane] = isX ? xm[(t - tJ)*32 + lane] : 0u;
            if (entry >= 0) {
                const int j = entry & J_INDEX_MASK, code = entry >> J_SHIFT_BITS;
                const uint4 q = a.posq[j];
                const float4 pr = a.par[j];
                const LatticeShift ls = crossShift(__float_as_int(pr.z), a);
                const int kz = code/15 - 1;
                const long long shx = ((long long) (code % 5 - 2) << 32) + ((code/5) % 3 - 1)*a.shiftB + kz*a.shiftCx + ls.x;
                const long long shy = ((long long) ((code/5) % 3 - 1) << 32) + kz*a.shiftCy + ls.y;
                w.jX[lane] = (long long) q.x + shx - (long long) lo.x;
                w.jY[lane] = (long long) q.y + shy - (long long) lo.y;
                w.jZ[lane] = (long long) q.z + ((long long) kz << 32) + ls.z - (long long) lo.z;
                w.jQ[lane] = a.q64[j];
                const double2 se = a.sigEpsD[__float_as_int(pr.w)];
                w.jSig[lane] = se.x; w.jEps[lane] = se.y;
                w.jSub[lane] = __float_as_int(pr.z) & 7;
                w.jIndex[lane] = j;
            }
            __syncwarp();
            double fjx = 0.0, fjy = 0.0, fjz = 0.0;
#pragma unroll 1
            for (int k = 0; k < 32; k++) {
                const int slot = (lane + k) & 31;
                const int jIndex = w.jIndex[slot];
                bool on = iValid && jIndex >= 0 && !((w.jMask[slot] >> lane) & 1u);
                // j - i
                const double ex = (double) (w.jX[slot] - wix)*a.dsx;
                const double ey = (double) (w.jY[slot] - wiy)*a.dsy;
                const double ez = (double) (w.jZ[slot] - wiz)*a.dsz;
                const double r2 = ex*ex + ey*ey + ez*ez;
                on = on && r2 <= a.rc2d;
                if (on) {
                    const double r = sqrt(r2), invR = 1.0/r, invR2 = invR*invR;
                    const double sig = sigi + w.jSig[slot];
                    double s2 = sig*invR; s2 *= s2;
                    const double s6 = s2*s2*s2;
                    const double eps = epsi*w.jEps[slot];
                    double ev = eps*(s6 - 1.0)*s6;
                    double fv = eps*(12.0*s6 - 6.0)*s6*invR2;
                    const double qq = qi*w.jQ[slot];
                    double ec, fc;
                    if (CMODE != 0) {
                        const double ar = a.alphaD*r;
                        const double erfcv = erfc(ar), ex2 = exp(-ar*ar);
                        ec = qq*invR*erfcv;
                        fc = qq*invR*invR2*(erfcv + TWO_OVER_SQRT_PI*ar*ex2);
                    }
                    else {
                        ec = qq*(invR + a.krfD*r2 - a.crfD);
                        fc = qq*(invR*invR2 - 2.0*a.krfD);
                    }
                    if (CMODE == 2) {                     // LJPME, :398-426
                        const double sg = sigi*w.jSig[slot];
                        const double c6 = 64.0*sg*sg*sg*eps;
                        const double dar2 = a.dalpha2D*r2, exd = exp(-dar2);
                        const double p2 = 1.0 + dar2 + 0.5*dar2*dar2;
                        const double c6r6 = c6*invR2*invR2*invR2;
                        fv += 6.0*c6r6*invR2*(1.0 - exd*(p2 + dar2*dar2*dar2/6.0));
                        double sc = sig*sig;
                        const double sc6 = sc*sc*sc*a.invCut6D;
                        ev += c6r6*(1.0 - exd*p2) + eps*(1.0 - sc6)*sc6 - c6*a.shiftMultD;
                    }
                    if (a.useSwitch && r > a.rswitchD) {
                        const double wd = 1.0/(a.rcutD - a.rswitchD), u = (r - a.rswitchD)*wd;
                        const double sv = 1.0 + u*u*u*(-10.0 + u*(15.0 - u*6.0));
                        const double sd = u*u*(-30.0 + u*(60.0 - u*30.0))*wd;
                        fv = sv*fv - ev*sd*invR;
                        ev *= sv;
                    }
                    const int sl = triSlice(si, w.jSub[slot]);
                    const double dEdR = a.lamD[2*sl + 1]*fv + a.lamD[2*sl]*fc;
                    // force on i = dEdR (x_i - x_j) = -dEdR e;  on j the opposite
                    fix -= dEdR*ex; fiy -= dEdR*ey; fiz -= dEdR*ez;
                    fjx += dEdR*ex; fjy += dEdR*ey; fjz += dEdR*ez;
                    acc[2*sl] += ec;
                    acc[2*sl + 1] += ev;
                }
                // the accumulator of slot (lane + k) moves to the lane that meets that slot next: lane - 1
                fjx = __shfl_sync(FULL_MASK, fjx, (lane + 1) & 31);
                fjy = __shfl_sync(FULL_MASK, fjy, (lane + 1) & 31);
                fjz = __shfl_sync(FULL_MASK, fjz, (lane + 1) & 31);
            }
            // 32 rotations: lane l holds the total of slot l again
            const int jIndex = w.jIndex[lane];
            if (jIndex >= 0) {
                if (fjx != 0.0) atomicAdd(a.force + jIndex, toFixed(fjx));
                if (fjy != 0.0) atomicAdd(a.force + (size_t) a.Npad + jIndex, toFixed(fjy));
                if (fjz != 0.0) atomicAdd(a.force + 2*(size_t) a.Npad + jIndex, toFixed(fjz));
            }
        }
        if (iValid) {
            if (fix != 0.0) atomicAdd(a.force + first + lane, toFixed(fix));
            if (fiy != 0.0) atomicAdd(a.force + (size_t) a.Npad + first + lane, toFixed(fiy));
            if (fiz != 0.0) atomicAdd(a.force + 2*(size_t) a.Npad + first + lane, toFixed(fiz));
        }
    }
    for (int k = 0; k < a.nE; k++) {
        const double v = warpSum(acc[k]);
        if (lane == 0 && v != 0.0) atomicAdd(&shE[k], v);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < a.nE; k += blockDim.x)
        if (shE[k] != 0.0) atomicAdd(a.energy + k, shE[k]);
}

template <int EMODE, int CMODE, int MODE, bool CUBIC, int MINCTAS>
static int launchPairK(Context& c, const PairArgs& a) {
    static bool attr[64] = {false};
    const size_t smem = sizeof(WarpScratch)*PAIR_WARPS + (EMODE == 2 && MODE == 0 && CMODE != 0 ? sizeof(double)*3*ERFC_TAB_MAX_ROWS : 0);
    if (!attr[c.device & 63]) {
        NBS_CUDA_CHECK(cudaFuncSetAttribute(k_pair<EMODE, CMODE, MODE, CUBIC, MINCTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr[c.device & 63] = true;
    }
    // persistent grid: as many CTAs as are resident at once
    static int perSM[64] = {0};
    if (perSM[c.device & 63] == 0) {
        int n = 0;
        NBS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pair<EMODE, CMODE, MODE, CUBIC, MINCTAS>, PAIR_WARPS*32, smem));
        perSM[c.device & 63] = std::max(1, n);
    }
    k_pair<EMODE, CMODE, MODE, CUBIC, MINCTAS><<<perSM[c.device & 63]*c.numSMs, PAIR_WARPS*32, smem, c.stream>>>(a);
    return NBS_OK;
}

template <int EMODE, int CMODE, int MODE, bool CUBIC>
static int launchPairT(Context& c, const PairArgs& a) {
    if constexpr (MODE == 0 && CMODE != 2) {
        static const int forced = getenv("NBS_PAIR_MINCTAS") ? atoi(getenv("NBS_PAIR_MINCTAS")) : 0;     // tuning experiments
        if (forced == 3 || (forced == 0 && c.N >= 300000 && EMODE == 0)) return launchPairK<EMODE, CMODE, MODE, CUBIC, 3>(c, a);
    }
    return launchPairK<EMODE, CMODE, MODE, CUBIC, PAIR_MIN_CTAS>(c, a);
}

template <int CMODE>
static int launchPairE(Context& c, const PairArgs& a, int emode, bool cubic) {
    if (emode == 0) return launchPairT<0, CMODE, 0, false>(c, a);
    if (emode == 1) return launchPairT<1, CMODE, 0, false>(c, a);
    return cubic ? launchPairT<2, CMODE, 0, true>(c, a) : launchPairT<2, CMODE, 0, false>(c, a);
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
    a.ds2 = a.dsx*a.dsx;
    const bool cubic = g.box[0] == g.box[1] && g.box[1] == g.box[2];
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
    if (mode == 0 && (c.flags & NBS_FLAG_DOUBLE)) {
        // Precision = double: its own kernel (energies always computed; they cost nothing extra there)
        a.dalpha2D = c.dispAlpha*c.dispAlpha;
        a.invCut6D = noCutoff ? 0.0 : std::pow(c.cutoff, -6.0);
        { const double dac2 = a.dalpha2D*c.cutoff*c.cutoff; a.shiftMultD = a.invCut6D*(1.0 - std::exp(-dac2)*(1.0 + dac2 + 0.5*dac2*dac2)); }
        a.rswitchD = c.switchDist; a.rcutD = c.cutoff;
        a.sigEpsD = c.dSigEpsD.d;
        NBS_CUDA_CHECK(c.dLamD.ensure(2*MAX_SLICES));
        NBS_CUDA_CHECK(cudaMemcpyAsync(c.dLamD.d, c.lambdas.data(), sizeof(double)*2*c.nSl, cudaMemcpyHostToDevice, c.stream));
        a.lamD = c.dLamD.d;
        const int ctas = 4*c.numSMs;
        if (c.ljpme()) k_pair_f64<2><<<ctas, 256, 0, c.stream>>>(a);
        else if (pme) k_pair_f64<1><<<ctas, 256, 0, c.stream>>>(a);
        else k_pair_f64<0><<<ctas, 256, 0, c.stream>>>(a);
        c.launches++;
        timerMark(c, "pair");
        return NBS_OK;
    }
    if (mode != 0) {
        NBS_CUDA_CHECK(cudaMemsetAsync(c.dCounters.d + 3, 0, sizeof(int), c.stream));      // rewind the work cursor
        status = mode == 1 ? launchPairT<0, 1, 1, false>(c, a) : launchPairT<0, 1, 2, false>(c, a);
    }
    else if (c.ljpme()) status = launchPairE<2>(c, a, emode, cubic);
    else if (pme) status = launchPairE<1>(c, a, emode, cubic);
    else status = launchPairE<0>(c, a, emode, cubic);
    if (status != NBS_OK) return status;
    c.launches++;
    timerMark(c, mode == 0 ? "pair" : "pair_set");
    return NBS_OK;
}

} // namespace nbs

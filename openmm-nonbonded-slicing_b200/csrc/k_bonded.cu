// k_bonded.cu -- exceptions ("1-4" pairs) and Ewald exclusion corrections, one thread per exception.
//
// Follows ReferenceSlicedLJCoulomb14::calculateBondIxn (platforms/reference/src/
// ReferenceSlicedLJCoulomb14.cpp:61-95: plain Coulomb + LJ, no cutoff, lambda-scaled force, unscaled
// slice energy) and the exclusion loop of ReferenceSlicedLJCoulombIxn::calculateEwaldIxn
// (ReferenceSlicedLJCoulombIxn.cpp:449-506: subtract the reciprocal-space part, erf(alpha r)/r, for
// EVERY exception pair using the particles' current charges; r -> 0 limit when erf(alpha r) <= 1e-6).
// The reference's device version splits this over two kernels (nonbondedExceptions.cc, pmeExclusions.cc);
// here it is one kernel in double precision -- O(N) pairs, nowhere near the critical path.
#include "nbs_internal.h"
#include "nbs_device.cuh"

namespace nbs {

struct BondedArgs {
    int nExc, Npad;
    int rank, nRanks;                // exception e belongs to rank e % nRanks
    int doExclusionCorrection;       // Ewald, PME, LJPME
    int periodic;                    // exceptionsUsePeriodic
    double3 box, invBox, tilt;       // tilt = (bx, cx, cy) of a triclinic box
    double alpha;
    int ljpme;                       // LJPME: also back out the dispersion grid's contribution (:487-504)
    double dispAlpha;
    const double* c6;                // particle order
    const int2* pairs; const double4* params; const int* slices;
    const double* charge; const double* pos; const int* origToSorted;
    const int* slotOf;               // particle -> slot of the caller's position array (or NULL)
    unsigned long long* force; double* energy;
    double lamC[MAX_SLICES], lamV[MAX_SLICES];
};

__global__ void k_bonded(const BondedArgs a) {
    const int e = (blockIdx.x*blockDim.x + threadIdx.x)*a.nRanks + a.rank;     // this rank's exceptions, densely packed into the grid
    const int lane = threadIdx.x & 31;
    int slice = -1;
    double eCoul = 0.0, eVdw = 0.0;
    if (e < a.nExc) {
        const int2 pr = a.pairs[e];
        const int s1 = a.slotOf ? a.slotOf[pr.x] : pr.x, s2 = a.slotOf ? a.slotOf[pr.y] : pr.y;
        double dx = a.pos[3*s1] - a.pos[3*s2], dy = a.pos[3*s1+1] - a.pos[3*s2+1], dz = a.pos[3*s1+2] - a.pos[3*s2+2];
        if (a.periodic) {
            // ReferenceForce::getDeltaRPeriodic [external]: subtract whole c, then b, then a vectors (tilt = bx, cx, cy)
            const double kz = floor(dz*a.invBox.z + 0.5);
            dx -= kz*a.tilt.y; dy -= kz*a.tilt.z; dz -= kz*a.box.z;
            const double ky = floor(dy*a.invBox.y + 0.5);
            dx -= ky*a.tilt.x; dy -= ky*a.box.y;
            dx -= a.box.x*floor(dx*a.invBox.x + 0.5);
        }
        const double r2 = dx*dx + dy*dy + dz*dz;
        const double r = sqrt(r2);
        const double inverseR = 1.0/r;
        const double4 p = a.params[e];
        slice = a.slices[e];
        double dEdR = 0.0;                       // force on particle 1 is +dEdR*delta
        if (p.w != 0.0) {                        // a 1-4 interaction
            double sig2 = inverseR*p.x;
            sig2 *= sig2;
            const double sig6 = sig2*sig2*sig2;
            double f = a.lamV[slice]*p.y*(12.0*sig6 - 6.0)*sig6;
            f += a.lamC[slice]*p.z*inverseR;
            dEdR += f*inverseR*inverseR;
            eVdw += p.y*(sig6 - 1.0)*sig6;
            eCoul += p.z*inverseR;
        }
        if (a.doExclusionCorrection) {
            const double qq = kOne4PiEps0*a.charge[pr.x]*a.charge[pr.y];
            const double alphaR = a.alpha*r;
            const double erfAlphaR = erf(alphaR);
            if (erfAlphaR > 1e-6) {
                const double SQRT_PI = 1.7724538509055160273;
                double f = qq*inverseR*inverseR*inverseR*(erfAlphaR - 2*alphaR*exp(-alphaR*alphaR)/SQRT_PI);
                dEdR -= a.lamC[slice]*f;
                eCoul -= qq*inverseR*erfAlphaR;
            }
            else
                eCoul -= a.alpha*1.1283791670955125739*qq;
            if (a.ljpme) {
                // dispersion: only the reciprocal-space (multiplicative C6) part is removed, ReferenceSlicedLJCoulombIxn.cpp:487-504
                const double dar2 = a.dispAlpha*a.dispAlpha*r2, dar4 = dar2*dar2, dar6 = dar4*dar2;
                const double inverseR2 = inverseR*inverseR;
                const double c6ij = a.c6[pr.x]*a.c6[pr.y];
                const double inverseR6 = inverseR2*inverseR2*inverseR2;
                const double expDar2 = exp(-dar2);
                eVdw += c6ij*inverseR6*(1.0 - expDar2*(1.0 + dar2 + 0.5*dar4));
                dEdR += a.lamV[slice]*6.0*c6ij*inverseR6*inverseR2*(1.0 - expDar2*(1.0 + dar2 + 0.5*dar4 + dar6/6.0));
            }
        }
        if (dEdR != 0.0) {
            const int i = a.origToSorted[pr.x], j = a.origToSorted[pr.y];
            atomicAdd(a.force + i, toFixed(dEdR*dx));
            atomicAdd(a.force + a.Npad + i, toFixed(dEdR*dy));
            atomicAdd(a.force + 2*(size_t) a.Npad + i, toFixed(dEdR*dz));
            atomicAdd(a.force + j, toFixed(-dEdR*dx));
            atomicAdd(a.force + a.Npad + j, toFixed(-dEdR*dy));
            atomicAdd(a.force + 2*(size_t) a.Npad + j, toFixed(-dEdR*dz));
        }
    }
    // warp-aggregate the slice energies: one pair of atomics per distinct slice per warp
    unsigned pending = __ballot_sync(FULL_MASK, slice >= 0);
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int s = __shfl_sync(FULL_MASK, slice, leader);
        const bool mine = slice == s;
        const double c = warpSum(mine ? eCoul : 0.0), v = warpSum(mine ? eVdw : 0.0);
        if (lane == leader) {
            if (c != 0.0) atomicAdd(a.energy + 2*s, c);
            if (v != 0.0) atomicAdd(a.energy + 2*s + 1, v);
        }
        pending &= ~__ballot_sync(FULL_MASK, mine);
    }
}

int launchBonded(Context& c, const double* dPos, bool periodicBox) {
    if (c.nExc == 0) return NBS_OK;
    BondedArgs a;
    a.nExc = c.nExc; a.Npad = c.Npad;
    a.rank = c.rank; a.nRanks = c.nRanks;
    a.doExclusionCorrection = c.ewaldDirect() ? 1 : 0;
    a.periodic = (c.excPeriodic && periodicBox) ? 1 : 0;
    a.box = make_double3(c.geom.box[0], c.geom.box[1], c.geom.box[2]);
    a.invBox = make_double3(c.geom.invBox[0], c.geom.invBox[1], c.geom.invBox[2]);
    a.tilt = make_double3(c.geom.tilt[0], c.geom.tilt[1], c.geom.tilt[2]);
    a.alpha = c.alpha;
    a.ljpme = c.ljpme() ? 1 : 0;
    a.dispAlpha = c.dispAlpha;
    a.c6 = c.dC6D.d;
    a.pairs = c.dExcPair.d; a.params = c.dExcParam.d; a.slices = c.dExcSlice.d;
    a.charge = c.dCharge.d; a.pos = dPos; a.origToSorted = c.dOrigToSorted.d;
    a.slotOf = nullptr;
    a.force = c.dForce.d; a.energy = c.dEnergy.d;
    for (int s = 0; s < MAX_SLICES; s++) {
        a.lamC[s] = s < c.nSl ? c.lambdas[2*s] : 1.0;
        a.lamV[s] = s < c.nSl ? c.lambdas[2*s+1] : 1.0;
    }
    const int T = 128;
    const int mine = (c.nExc + c.nRanks - 1)/c.nRanks;
    k_bonded<<<(mine+T-1)/T, T, 0, c.stream>>>(a);
    c.launches++;
    timerMark(c, "bonded");
    return NBS_OK;
}

} // namespace nbs

// ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the reference's own, UNMODIFIED
// Reference-platform arithmetic (compiled from /root/reference by oracle/Makefile into
// oracle/_ref/) the way ReferenceCalcSlicedNonbondedForceKernel::execute does
// (platforms/reference/src/ReferenceNonbondedSlicingKernels.cpp:187-250).  Only the glue that
// needs OpenMM proper -- the Force description, computeParameters and the neighbour list -- comes
// from oracle_common.cpp; every pair, PME and exception formula executed here is the reference's.
#include "oracle_common.h"
#include "internal/ReferenceSlicedLJCoulombIxn.h"
#include "internal/ReferenceSlicedLJCoulomb14.h"
#include <chrono>
#include <string>

namespace nbs_oracle {

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

std::string executeReference(const System& s, const double* pos, const Box& box, const double* lambdaTable,
                             bool includeDirect, bool includeReciprocal, double* forces, double* sliceEnergiesOut,
                             PairList& neighbors, double timings[4]) {
    using namespace NonbondedSlicing;
    try {
        const int method = s.method;
        const bool periodic = method == NBS_METHOD_CUTOFF_PERIODIC;
        const bool ewald = method == NBS_METHOD_EWALD;
        const bool pme = method == NBS_METHOD_PME;
        const bool ljpme = method == NBS_METHOD_LJPME;
        // kmax / the dispersion grid come with the description (calcEwaldParameters / calcPMEParameters are
        // OpenMM's [external]; the host mirror restates them, api.py SlicedNonbondedForceImpl)
        if (ewald && !box.rectangular())
            return "SlicedNonbondedForce: Ewald is not supported with non-rectangular boxes.  Use PME instead.";
        std::vector<OpenMM::Vec3> posData(s.n), forceData(s.n);
        for (int i = 0; i < s.n; i++) {
            posData[i] = OpenMM::Vec3(pos[3*i], pos[3*i+1], pos[3*i+2]);
            forceData[i] = OpenMM::Vec3(forces[3*i], forces[3*i+1], forces[3*i+2]);
        }
        OpenMM::Vec3 boxVectors[3];
        for (int i = 0; i < 3; i++) boxVectors[i] = OpenMM::Vec3(box.v[i][0], box.v[i][1], box.v[i][2]);
        std::vector<std::vector<double>> particleParamArray(s.n, std::vector<double>(4, 0.0));
        for (int i = 0; i < s.n; i++)
            for (int k = 0; k < 3; k++) particleParamArray[i][k] = s.particleParams[i][k];
        std::vector<std::vector<double>> sliceLambdas(s.numSlices, std::vector<double>(2));
        std::vector<std::vector<double>> sliceEnergies(s.numSlices, std::vector<double>(2, 0.0));
        for (int k = 0; k < s.numSlices; k++) {
            sliceLambdas[k][0] = lambdaTable[2*k];
            sliceLambdas[k][1] = lambdaTable[2*k+1];
        }
        int gridSize[3] = {s.grid[0], s.grid[1], s.grid[2]};

        double t0 = now();
        ReferenceSlicedLJCoulombIxn clj;
        OpenMM::NeighborList neighborList;
        if (method != NBS_METHOD_NOCUTOFF) {
            buildNeighborList(s, pos, box, periodic || ewald || pme || ljpme, neighbors);
            neighborList.assign(neighbors.begin(), neighbors.end());
            clj.setUseCutoff(s.cutoff, neighborList, s.rfDielectric);
        }
        double t1 = now();
        if (periodic || ewald || pme || ljpme) {
            double minAllowedSize = 1.999999*s.cutoff;
            if (box.v[0][0] < minAllowedSize || box.v[1][1] < minAllowedSize || box.v[2][2] < minAllowedSize)
                return "The periodic box size has decreased to less than twice the nonbonded cutoff.";
            clj.setPeriodic(boxVectors);
            clj.setPeriodicExceptions(s.exceptionsPeriodic);
        }
        if (ewald)
            clj.setUseEwald(s.alpha, s.kmax[0], s.kmax[1], s.kmax[2]);
        if (pme)
            clj.setUsePME(s.alpha, gridSize);
        if (ljpme) {
            int dispersionGrid[3] = {s.dispersionGrid[0], s.dispersionGrid[1], s.dispersionGrid[2]};
            clj.setUsePME(s.alpha, gridSize);
            clj.setUseLJPME(s.dispersionAlpha, dispersionGrid);
        }
        if (s.useSwitch)
            clj.setUseSwitchingFunction(s.switchingDistance);
        double tRecip = 0;
        if ((pme || ewald || ljpme) && includeReciprocal) {
            double a = now();
            clj.calculatePairIxn(s.n, posData, s.numSubsets, s.subsets, particleParamArray, sliceLambdas, s.exclusions,
                                 forceData, sliceEnergies, false, true);
            tRecip = now() - a;
            clj.calculatePairIxn(s.n, posData, s.numSubsets, s.subsets, particleParamArray, sliceLambdas, s.exclusions,
                                 forceData, sliceEnergies, includeDirect, false);
        }
        else
            clj.calculatePairIxn(s.n, posData, s.numSubsets, s.subsets, particleParamArray, sliceLambdas, s.exclusions,
                                 forceData, sliceEnergies, includeDirect, includeReciprocal);
        if (includeDirect) {
            ReferenceSlicedLJCoulomb14 nonbonded14;
            if (s.exceptionsPeriodic)
                nonbonded14.setPeriodic(boxVectors);
            for (int k = 0; k < s.num14; k++) {
                std::vector<int> indices = {s.index14[k][0], s.index14[k][1]};
                std::vector<double> params = {s.params14[k][0], s.params14[k][1], s.params14[k][2]};
                int slice = s.slice14[k];
                nonbonded14.calculateBondIxn(indices, posData, params, forceData, sliceLambdas[slice], sliceEnergies[slice]);
            }
            if (periodic || ewald || pme) {       // not LJPME: ReferenceNonbondedSlicingKernels.cpp:244
                double volume = box.v[0][0]*box.v[1][1]*box.v[2][2];
                for (int slice = 0; slice < s.numSlices; slice++)
                    sliceEnergies[slice][1] += s.dispersionCoefficients[slice]/volume;
            }
        }
        double t2 = now();
        for (int i = 0; i < s.n; i++)
            for (int k = 0; k < 3; k++) forces[3*i+k] = forceData[i][k];
        for (int k = 0; k < s.numSlices; k++) {
            sliceEnergiesOut[2*k] = sliceEnergies[k][0];
            sliceEnergiesOut[2*k+1] = sliceEnergies[k][1];
        }
        timings[0] = t1 - t0;
        timings[1] = (t2 - t1) - tRecip;
        timings[2] = tRecip;
        timings[3] = t2 - t0;
    }
    catch (std::exception& e) {
        return std::string("reference threw: ") + e.what();
    }
    return "";
}

} // namespace nbs_oracle

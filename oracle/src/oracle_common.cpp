// oracle_common.cpp -- TEST INFRASTRUCTURE ONLY.  See oracle_common.h.
#include "oracle_common.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <tuple>

namespace nbs_oracle {

std::string buildSystem(const nbs_system_desc& d, const double* globalValues, System& s) {
    if (d.struct_size != (int32_t) sizeof(nbs_system_desc))
        return "nbs_system_desc.struct_size mismatch";
    s.n = d.num_particles;
    s.numSubsets = d.num_subsets;
    s.numSlices = d.num_subsets*(d.num_subsets+1)/2;
    s.method = d.method;
    s.cutoff = d.cutoff;
    s.switchingDistance = d.switching_distance;
    s.rfDielectric = d.rf_dielectric;
    s.alpha = d.ewald_alpha;
    for (int k = 0; k < 3; k++) s.grid[k] = d.pme_grid[k];
    for (int k = 0; k < 3; k++) s.kmax[k] = d.ewald_kmax[k];
    s.dispersionAlpha = d.dispersion_alpha;
    for (int k = 0; k < 3; k++) s.dispersionGrid[k] = d.dispersion_grid[k];
    // ReferenceNonbondedSlicingKernels.cpp:146-154 -- NoCutoff disables the switch; :168-171 the
    // non-periodic methods never use periodic exceptions.
    s.useSwitch = d.method != NBS_METHOD_NOCUTOFF && d.use_switching_function;
    s.exceptionsPeriodic = (d.method == NBS_METHOD_NOCUTOFF || d.method == NBS_METHOD_CUTOFF_NONPERIODIC)
                               ? false : d.exceptions_use_periodic != 0;
    s.subsets.assign(d.subsets, d.subsets + s.n);
    for (int i = 0; i < s.n; i++)
        if (s.subsets[i] < 0 || s.subsets[i] >= s.numSubsets)
            return "particle subset out of range";

    // Which exceptions are "1-4" interactions (:88-111): any exception with chargeProd != 0,
    // epsilon != 0 (BASE values) or an attached offset.  Every exception is also an exclusion.
    std::set<int> exceptionsWithOffsets;
    for (int i = 0; i < d.num_exception_offsets; i++)
        exceptionsWithOffsets.insert(d.exception_offset_indices[2*i+1]);
    s.exclusions.assign(s.n, std::set<int>());
    std::vector<int> nb14s;
    std::map<int, int> nb14Index;
    for (int i = 0; i < d.num_exceptions; i++) {
        int p1 = d.exception_particles[2*i], p2 = d.exception_particles[2*i+1];
        if (p1 < 0 || p1 >= s.n || p2 < 0 || p2 >= s.n)
            return "SlicedNonbondedForce: Illegal particle index for an exception";
        s.exclusions[p1].insert(p2);
        s.exclusions[p2].insert(p1);
        double chargeProd = d.exception_params[3*i], epsilon = d.exception_params[3*i+2];
        if (chargeProd != 0.0 || epsilon != 0.0 || exceptionsWithOffsets.count(i)) {
            nb14Index[i] = (int) nb14s.size();
            nb14s.push_back(i);
        }
    }
    s.num14 = (int) nb14s.size();

    // computeParameters (:339-392): offsets are applied to the base values, then transformed.
    std::vector<double> q(d.charges, d.charges + s.n), sig(d.sigmas, d.sigmas + s.n), eps(d.epsilons, d.epsilons + s.n);
    for (int i = 0; i < d.num_particle_offsets; i++) {
        double value = globalValues[d.particle_offset_indices[2*i]];
        int index = d.particle_offset_indices[2*i+1];
        q[index] += value*d.particle_offset_scales[3*i];
        sig[index] += value*d.particle_offset_scales[3*i+1];
        eps[index] += value*d.particle_offset_scales[3*i+2];
    }
    s.particleParams.resize(s.n);
    for (int i = 0; i < s.n; i++)
        s.particleParams[i] = {0.5*sig[i], 2.0*std::sqrt(eps[i]), q[i]};

    s.index14.resize(s.num14);
    s.params14.resize(s.num14);
    s.slice14.resize(s.num14);
    std::vector<double> q14(s.num14), sig14(s.num14), eps14(s.num14);
    for (int k = 0; k < s.num14; k++) {
        int e = nb14s[k];
        s.index14[k] = {d.exception_particles[2*e], d.exception_particles[2*e+1]};
        s.slice14[k] = sliceIndex(s.subsets[s.index14[k][0]], s.subsets[s.index14[k][1]]);
        q14[k] = d.exception_params[3*e];
        sig14[k] = d.exception_params[3*e+1];
        eps14[k] = d.exception_params[3*e+2];
    }
    for (int i = 0; i < d.num_exception_offsets; i++) {
        double value = globalValues[d.exception_offset_indices[2*i]];
        int index = nb14Index[d.exception_offset_indices[2*i+1]];
        q14[index] += value*d.exception_offset_scales[3*i];
        sig14[index] += value*d.exception_offset_scales[3*i+1];
        eps14[index] += value*d.exception_offset_scales[3*i+2];
    }
    for (int k = 0; k < s.num14; k++)
        s.params14[k] = {sig14[k], 4.0*eps14[k], q14[k]};

    s.dispersionCoefficients.assign(s.numSlices, 0.0);
    if (d.dispersion_coefficients != nullptr)
        for (int k = 0; k < s.numSlices; k++) s.dispersionCoefficients[k] = d.dispersion_coefficients[k];
    return "";
}

// ---------------------------------------------------------------------------------------------
// Neighbour list.  Contract of computeNeighborListVoxelHash(list, N, positions, exclusions, box,
// usePeriodic, maxDistance = cutoff, minDistance = 0) as called at
// ReferenceNonbondedSlicingKernels.cpp:197: every unordered pair with minimum-image distance
// r^2 <= cutoff^2 that is not in `exclusions`, reported once.  (OpenMM reports the later atom
// first; the order only affects summation rounding.)
// ---------------------------------------------------------------------------------------------
void buildNeighborList(const System& s, const double* pos, const Box& box, bool periodic, PairList& out) {
    out.clear();
    const int n = s.n;
    const double rc2 = s.cutoff*s.cutoff;
    auto accept = [&](int i, int j) {      // i > j
        double d[3];
        if (periodic) deltaPeriodic(pos + 3*i, pos + 3*j, box, d);
        else for (int k = 0; k < 3; k++) d[k] = pos[3*i+k] - pos[3*j+k];
        double r2 = d[0]*d[0] + d[1]*d[1] + d[2]*d[2];
        if (r2 > rc2) return;
        if (s.exclusions[i].count(j)) return;
        out.push_back(std::make_pair((unsigned) i, (unsigned) j));
    };
    if (periodic && !box.rectangular()) {          // test-sized systems only
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++) accept(i, j);
        return;
    }
    // Cell list: cells at least `cutoff` wide.
    double lo[3], len[3];
    int nc[3];
    if (periodic) {
        for (int k = 0; k < 3; k++) {
            lo[k] = 0;
            len[k] = box.v[k][k];
            nc[k] = std::max(1, (int) std::floor(len[k]/s.cutoff));
        }
    }
    else {
        double hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = 1e300; hi[k] = -1e300; }
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], pos[3*i+k]);
                hi[k] = std::max(hi[k], pos[3*i+k]);
            }
        for (int k = 0; k < 3; k++) {
            len[k] = std::max(hi[k] - lo[k], 1e-9)*(1 + 1e-9);
            nc[k] = std::max(1, std::min(256, (int) std::floor(len[k]/s.cutoff)));
        }
    }
    auto cellOf = [&](int i, int c[3]) {
        for (int k = 0; k < 3; k++) {
            double f = (pos[3*i+k] - lo[k])/len[k];
            if (periodic) f -= std::floor(f);
            int ck = (int) (f*nc[k]);
            c[k] = std::min(std::max(ck, 0), nc[k]-1);
        }
    };
    const size_t numCells = (size_t) nc[0]*nc[1]*nc[2];
    std::vector<int> cellStart(numCells+1, 0), cellAtoms(n), cellIndex(n);
    for (int i = 0; i < n; i++) {
        int c[3];
        cellOf(i, c);
        cellIndex[i] = (c[0]*nc[1] + c[1])*nc[2] + c[2];
        cellStart[cellIndex[i]+1]++;
    }
    for (size_t c = 0; c < numCells; c++) cellStart[c+1] += cellStart[c];
    {
        std::vector<int> cursor(cellStart.begin(), cellStart.end()-1);
        for (int i = 0; i < n; i++) cellAtoms[cursor[cellIndex[i]]++] = i;
    }
    std::vector<int> neighborCells;
    for (int i = 0; i < n; i++) {
        int c[3];
        cellOf(i, c);
        neighborCells.clear();
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    int e[3] = {c[0]+dx, c[1]+dy, c[2]+dz};
                    bool ok = true;
                    for (int k = 0; k < 3; k++) {
                        if (periodic) e[k] = (e[k] + nc[k]) % nc[k];
                        else if (e[k] < 0 || e[k] >= nc[k]) ok = false;
                    }
                    if (ok) neighborCells.push_back((e[0]*nc[1] + e[1])*nc[2] + e[2]);
                }
        std::sort(neighborCells.begin(), neighborCells.end());
        neighborCells.erase(std::unique(neighborCells.begin(), neighborCells.end()), neighborCells.end());
        for (int cell : neighborCells)
            for (int a = cellStart[cell]; a < cellStart[cell+1]; a++) {
                int j = cellAtoms[a];
                if (j < i) accept(i, j);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// SlicedNonbondedForceImpl::calcDispersionCorrections, openmmapi/src/SlicedNonbondedForceImpl.cpp:263-354
// (evalIntegral :150-185).  `int count` products and `numParticles*(numParticles+1)` are kept as
// 32-bit ints on purpose.
// ---------------------------------------------------------------------------------------------
// 32-bit signed multiply with the wrap-around the reference's `int` expressions show on x86-64
// (formally UB there; made explicit here so the oracle is deterministic).
static inline int mulInt32(int a, int b) {
    return (int) (int32_t) ((uint32_t) a*(uint32_t) b);
}

static double evalIntegral(double r, double rs, double rc, double sigma) {
    double A = 1/(rc-rs);
    double A2 = A*A;
    double A3 = A2*A;
    double sig2 = sigma*sigma;
    double sig6 = sig2*sig2*sig2;
    double rs2 = rs*rs;
    double rs3 = rs*rs2;
    double r2 = r*r;
    double r3 = r*r2;
    double r4 = r*r3;
    double r5 = r*r4;
    double r6 = r*r5;
    double r9 = r3*r6;
    return sig6*A3*((
        sig6*(
            + rs3*28*(6*rs2*A2 + 15*rs*A + 10)
            - r*rs2*945*(rs2*A2 + 2*rs*A + 1)
            + r2*rs*1080*(2*rs2*A2 + 3*rs*A + 1)
            - r3*420*(6*rs2*A2 + 6*rs*A + 1)
            + r4*756*(2*rs*A2 + A)
            - r5*378*A2)
        -r6*(
            + rs3*84*(6*rs2*A2 + 15*rs*A + 10)
            - r*rs2*3780*(rs2*A2 + 2*rs*A + 1)
            + r2*rs*7560*(2*rs2*A2 + 3*rs*A + 1))
        )/(252*r9)
     - std::log(r)*10*(6*rs2*A2 + 6*rs*A + 1)
     + r*15*(2*rs*A2 + A)
     - r2*3*A2
    );
}

std::vector<double> dispersionCoefficients(const nbs_system_desc& d, const double* globalDefaults) {
    const int numSlices = d.num_subsets*(d.num_subsets+1)/2;
    std::vector<double> result(numSlices, 0.0);
    if (d.method == NBS_METHOD_NOCUTOFF || d.method == NBS_METHOD_CUTOFF_NONPERIODIC)
        return result;
    const int numParticles = d.num_particles;
    std::vector<double> sigma(d.sigmas, d.sigmas + numParticles), epsilon(d.epsilons, d.epsilons + numParticles);
    for (int i = 0; i < d.num_particle_offsets; i++) {
        double value = globalDefaults[d.particle_offset_indices[2*i]];
        int index = d.particle_offset_indices[2*i+1];
        sigma[index] += value*d.particle_offset_scales[3*i+1];
        epsilon[index] += value*d.particle_offset_scales[3*i+2];
    }
    typedef std::tuple<double, double, int> ParticleClass;
    std::map<ParticleClass, int> classCounts;
    for (int i = 0; i < numParticles; i++)
        classCounts[std::make_tuple(sigma[i], epsilon[i], d.subsets[i])]++;
    std::vector<double> sum1(numSlices, 0), sum2(numSlices, 0), sum3(numSlices, 0);
    const bool useSwitch = d.use_switching_function != 0;
    const double cutoff = d.cutoff, switchDist = d.switching_distance;
    for (auto& entry : classCounts) {
        double sig = std::get<0>(entry.first), eps = std::get<1>(entry.first);
        int subset = std::get<2>(entry.first);
        int count = mulInt32(entry.second, entry.second+1)/2;
        double sig2 = sig*sig, sig6 = sig2*sig2*sig2;
        int slice = subset*(subset+3)/2;
        sum1[slice] += count*eps*sig6*sig6;
        sum2[slice] += count*eps*sig6;
        if (useSwitch)
            sum3[slice] += count*eps*(evalIntegral(cutoff, switchDist, cutoff, sig)-evalIntegral(switchDist, switchDist, cutoff, sig));
    }
    for (auto c1 = classCounts.begin(); c1 != classCounts.end(); ++c1)
        for (auto c2 = classCounts.begin(); c2 != c1; ++c2) {
            double sig = 0.5*(std::get<0>(c1->first) + std::get<0>(c2->first));
            double eps = std::sqrt(std::get<1>(c1->first)*std::get<1>(c2->first));
            int slice = sliceIndex(std::get<2>(c1->first), std::get<2>(c2->first));
            int count = mulInt32(c1->second, c2->second);
            double sig2 = sig*sig, sig6 = sig2*sig2*sig2;
            sum1[slice] += count*eps*sig6*sig6;
            sum2[slice] += count*eps*sig6;
            if (useSwitch)
                sum3[slice] += count*eps*(evalIntegral(cutoff, switchDist, cutoff, sig)-evalIntegral(switchDist, switchDist, cutoff, sig));
        }
    double numInteractions = mulInt32(numParticles, numParticles+1)/2;
    for (int slice = 0; slice < numSlices; slice++) {
        sum1[slice] /= numInteractions;
        sum2[slice] /= numInteractions;
        sum3[slice] /= numInteractions;
        result[slice] = mulInt32(mulInt32(8, numParticles), numParticles)*kPi*(sum1[slice]/(9*std::pow(cutoff, 9)) - sum2[slice]/(3*std::pow(cutoff, 3)) + sum3[slice]);
    }
    return result;
}

} // namespace nbs_oracle

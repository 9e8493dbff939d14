// oracle_api.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points of the CPU oracles (loaded with
// ctypes by oracle/oracle.py).  Built twice:
//   oracle/liboracle_port.so        restatement in oracle_core.cpp                 (kind "port")
//   oracle/_ref/liboracle_ref.so    -DNBS_ORACLE_USE_REFERENCE: the reference's own unmodified
//                                   Reference-platform TUs via ref_bridge.cpp      (kind "reference")
#include "oracle_common.h"
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace nbs_oracle {
std::string executePort(const System& s, const double* pos, const Box& box, const double* lambdaTable,
                        bool includeDirect, bool includeReciprocal, double* forces, double* sliceEnergiesOut,
                        PairList& neighbors, double timings[4],
                        std::vector<double>* spreadDump, std::vector<double>* potentialDump);
std::string executeReference(const System& s, const double* pos, const Box& box, const double* lambdaTable,
                             bool includeDirect, bool includeReciprocal, double* forces, double* sliceEnergiesOut,
                             PairList& neighbors, double timings[4]);
}

static thread_local std::string lastError;

extern "C" {

const char* nbs_oracle_last_error(void) { return lastError.c_str(); }

const char* nbs_oracle_kind(void) {
#ifdef NBS_ORACLE_USE_REFERENCE
    return "reference";
#else
    return "port";
#endif
}

// One evaluation.  forces: double[N][3], ADDED to (callers zero it); slice_energies: double[nSl][2],
// overwritten; pairs (optional): int32[pair_capacity][2] as (min, max) in list order;
// timings: seconds {neighbour list, direct + exceptions, reciprocal, total}.
int nbs_oracle_execute(const nbs_system_desc* desc, const double* global_values, const double* lambdas,
                       const double* positions, const double* box9, int include_direct, int include_reciprocal,
                       double* forces, double* slice_energies, int64_t pair_capacity, int32_t* pairs,
                       int64_t* pair_count, uint64_t* pair_hash, double* timings,
                       double* spread_grid, double* potential_grid) {
    using namespace nbs_oracle;
    System s;
    std::string err = buildSystem(*desc, global_values, s);
    if (!err.empty()) { lastError = err; return NBS_ERR_INVALID; }
    Box box;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) box.v[i][j] = box9[3*i+j];
    PairList neighbors;
    double t[4] = {0, 0, 0, 0};
    std::vector<double> spread, potential;
#ifdef NBS_ORACLE_USE_REFERENCE
    err = executeReference(s, positions, box, lambdas, include_direct != 0, include_reciprocal != 0, forces,
                           slice_energies, neighbors, t);
#else
    err = executePort(s, positions, box, lambdas, include_direct != 0, include_reciprocal != 0, forces,
                      slice_energies, neighbors, t, spread_grid ? &spread : nullptr, potential_grid ? &potential : nullptr);
    if (spread_grid && !spread.empty()) std::memcpy(spread_grid, spread.data(), spread.size()*sizeof(double));
    if (potential_grid && !potential.empty()) std::memcpy(potential_grid, potential.data(), potential.size()*sizeof(double));
#endif
    if (!err.empty()) { lastError = err; return err.find("periodic box size") != std::string::npos ? NBS_ERR_BOX : NBS_ERR_UNSUPPORTED; }
    if (timings) for (int k = 0; k < 4; k++) timings[k] = t[k];
    uint64_t hash = 0;
    for (auto& p : neighbors) {
        uint32_t a = std::min(p.first, p.second), b = std::max(p.first, p.second);
        hash += nbs_pair_hash(a, b);
    }
    if (pair_count) *pair_count = (int64_t) neighbors.size();
    if (pair_hash) *pair_hash = hash;
    if (pairs && (int64_t) neighbors.size() <= pair_capacity)
        for (size_t k = 0; k < neighbors.size(); k++) {
            pairs[2*k] = (int32_t) std::min(neighbors[k].first, neighbors[k].second);
            pairs[2*k+1] = (int32_t) std::max(neighbors[k].first, neighbors[k].second);
        }
    return NBS_OK;
}

// Pairs (excluded or not) whose minimum-image r^2 lies within `band` of cutoff^2 -- used by the
// synthetic-system generator to keep every pair out of the guard band.
int nbs_oracle_band_pairs(int32_t n, const double* positions, const double* box9, double cutoff, double band,
                          int64_t capacity, int32_t* pairs, int64_t* count) {
    using namespace nbs_oracle;
    System s;
    s.n = n;
    s.cutoff = std::sqrt(cutoff*cutoff + band);
    s.exclusions.assign(n, std::set<int>());
    Box box;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) box.v[i][j] = box9[3*i+j];
    PairList all;
    buildNeighborList(s, positions, box, true, all);
    int64_t found = 0;
    for (auto& p : all) {
        double d[3];
        deltaPeriodic(positions + 3*p.first, positions + 3*p.second, box, d);
        double r2 = d[0]*d[0] + d[1]*d[1] + d[2]*d[2];
        if (std::fabs(r2 - cutoff*cutoff) < band) {
            if (found < capacity) { pairs[2*found] = (int32_t) p.first; pairs[2*found+1] = (int32_t) p.second; }
            found++;
        }
    }
    *count = found;
    return NBS_OK;
}

int nbs_oracle_dispersion_coefficients(const nbs_system_desc* desc, const double* global_defaults, double* out) {
    std::vector<double> c = nbs_oracle::dispersionCoefficients(*desc, global_defaults);
    for (size_t k = 0; k < c.size(); k++) out[k] = c[k];
    return NBS_OK;
}

} // extern "C"

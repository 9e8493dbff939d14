// host_tools.cpp -- host-only helpers for the synthetic-system generator (systems.py).
// Not part of the force path: nothing here computes forces or energies.
//
// nbs_tools_band_pairs: all pairs whose minimum-image r^2 (rectangular box, double precision)
// lies within `band` of cutoff^2.  The generator nudges such pairs out of the guard band so that
// "which pairs interact" is well defined for every arithmetic (bit-exact pair sets).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

extern "C" int nbs_tools_band_pairs(int32_t n, const double* pos, const double* boxLengths, double cutoff, double band,
                                    int64_t capacity, int32_t* pairs, int64_t* count) {
    const double reach = std::sqrt(cutoff*cutoff + band);
    int nc[3];
    for (int k = 0; k < 3; k++) nc[k] = std::max(1, (int) std::floor(boxLengths[k]/reach));
    const size_t numCells = (size_t) nc[0]*nc[1]*nc[2];
    std::vector<int> cellOf(n), start(numCells+1, 0), atoms(n);
    std::vector<double> wrapped((size_t) 3*n);
    for (int i = 0; i < n; i++) {
        int c[3];
        for (int k = 0; k < 3; k++) {
            double f = pos[3*i+k]/boxLengths[k];
            f -= std::floor(f);
            wrapped[3*(size_t) i+k] = f*boxLengths[k];
            c[k] = std::min((int) (f*nc[k]), nc[k]-1);
        }
        cellOf[i] = (c[0]*nc[1] + c[1])*nc[2] + c[2];
        start[cellOf[i]+1]++;
    }
    for (size_t c = 0; c < numCells; c++) start[c+1] += start[c];
    {
        std::vector<int> cursor(start.begin(), start.end()-1);
        for (int i = 0; i < n; i++) atoms[cursor[cellOf[i]]++] = i;
    }
    const double lo = cutoff*cutoff - band, hi = cutoff*cutoff + band;
    int64_t found = 0;
    std::vector<int> neighbors;
    for (int cx = 0; cx < nc[0]; cx++)
        for (int cy = 0; cy < nc[1]; cy++)
            for (int cz = 0; cz < nc[2]; cz++) {
                int cell = (cx*nc[1] + cy)*nc[2] + cz;
                neighbors.clear();
                for (int dx = -1; dx <= 1; dx++)
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dz = -1; dz <= 1; dz++) {
                            int ex = (cx+dx+nc[0])%nc[0], ey = (cy+dy+nc[1])%nc[1], ez = (cz+dz+nc[2])%nc[2];
                            neighbors.push_back((ex*nc[1] + ey)*nc[2] + ez);
                        }
                std::sort(neighbors.begin(), neighbors.end());
                neighbors.erase(std::unique(neighbors.begin(), neighbors.end()), neighbors.end());
                for (int a = start[cell]; a < start[cell+1]; a++) {
                    int i = atoms[a];
                    for (int other : neighbors)
                        for (int b = start[other]; b < start[other+1]; b++) {
                            int j = atoms[b];
                            if (j >= i) continue;
                            double r2 = 0;
                            for (int k = 0; k < 3; k++) {
                                double d = wrapped[3*(size_t) i+k] - wrapped[3*(size_t) j+k];
                                d -= boxLengths[k]*std::floor(d/boxLengths[k] + 0.5);
                                r2 += d*d;
                            }
                            if (r2 > lo && r2 < hi) {
                                if (found < capacity) { pairs[2*found] = j; pairs[2*found+1] = i; }
                                found++;
                            }
                        }
                }
            }
    *count = found;
    return 0;
}

// nbs_api.cu -- the C ABI (include/nbslice_b200.h) and the host-side sequencing of one evaluation.
//
// Host logic mirrors what the reference's Reference-platform kernel does around its arithmetic:
//   nbs_create / nbs_update_parameters  <- ReferenceCalcSlicedNonbondedForceKernel::initialize /
//        copyParametersToContext (ReferenceNonbondedSlicingKernels.cpp:59-185, 270-319)
//   applyParameters()                   <- computeParameters (:339-392): offsets, (sigma/2, 2 sqrt(eps), q),
//        exception (sigma, 4 eps, qq), "1-4" selection (:107)
//   nbs_execute                         <- execute (:187-268) minus the lambda-weighted sum, which is the
//        adapter's job (it owns the scaling-parameter names)
// There is no CPU fallback anywhere in this file: every path ends in CUDA kernels or an error code.
#include "nbs_internal.h"
#include <unistd.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <set>

namespace nbs {

unsigned long long gAllocEpoch = 1;
static thread_local std::string gLastError;
void setError(const std::string& message) { gLastError = message; }

void timerMark(Context& c, const char* name) {
    if (!c.profiling) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, c.stream);
    c.timer.names.push_back(name);
    c.timer.events.push_back(ev);
}

static void timerReset(Context& c) {
    for (cudaEvent_t ev : c.timer.events) cudaEventDestroy(ev);
    c.timer.events.clear();
    c.timer.names.clear();
}

static int fail(int status, const std::string& message) {
    setError(message);
    return status;
}

// ---- description -> host arrays ---------------------------------------------------------------
static int readDescription(Context& c, const nbs_system_desc& d, bool creating) {
    if (d.struct_size != (int32_t) sizeof(nbs_system_desc)) return fail(NBS_ERR_INVALID, "nbs_system_desc.struct_size mismatch");
    if (d.num_particles <= 0) return fail(NBS_ERR_INVALID, "num_particles must be positive");
    if (d.num_subsets < 1 || d.num_subsets > MAX_SUBSETS)
        return fail(NBS_ERR_UNSUPPORTED, "num_subsets must be between 1 and 8");
    if (!creating) {
        // ReferenceNonbondedSlicingKernels.cpp:271-272
        if (d.num_particles != c.N) return fail(NBS_ERR_INVALID, "updateParametersInContext: The number of particles has changed");
        if (d.num_subsets != c.nS) return fail(NBS_ERR_INVALID, "updateParametersInContext: The number of subsets has changed");
    }
    if (d.num_particles > J_INDEX_MASK) return fail(NBS_ERR_UNSUPPORTED, "too many particles");
    const int N = d.num_particles;
    std::vector<int> subsets(d.subsets, d.subsets + N);
    for (int s : subsets)
        if (s < 0 || s >= d.num_subsets) return fail(NBS_ERR_INVALID, "particle subset out of range");
    for (int i = 0; i < d.num_exceptions; i++)
        for (int k = 0; k < 2; k++) {
            int p = d.exception_particles[2*i+k];
            if (p < 0 || p >= N) return fail(NBS_ERR_INVALID, "SlicedNonbondedForce: Illegal particle index for an exception: " + std::to_string(p));
        }
    for (int i = 0; i < d.num_particle_offsets; i++) {
        if (d.particle_offset_indices[2*i] < 0 || d.particle_offset_indices[2*i] >= d.num_global_params ||
            d.particle_offset_indices[2*i+1] < 0 || d.particle_offset_indices[2*i+1] >= N)
            return fail(NBS_ERR_INVALID, "SlicedNonbondedForce: Illegal index for a particle parameter offset");
    }
    for (int i = 0; i < d.num_exception_offsets; i++) {
        if (d.exception_offset_indices[2*i] < 0 || d.exception_offset_indices[2*i] >= d.num_global_params ||
            d.exception_offset_indices[2*i+1] < 0 || d.exception_offset_indices[2*i+1] >= d.num_exceptions)
            return fail(NBS_ERR_INVALID, "SlicedNonbondedForce: Illegal index for an exception parameter offset");
    }
    // count the "1-4" exceptions (:88-111) -- base values or an attached offset
    std::set<int> withOffsets;
    for (int i = 0; i < d.num_exception_offsets; i++) withOffsets.insert(d.exception_offset_indices[2*i+1]);
    int num14 = 0;
    for (int i = 0; i < d.num_exceptions; i++)
        if (d.exception_params[3*i] != 0.0 || d.exception_params[3*i+2] != 0.0 || withOffsets.count(i)) num14++;
    if (!creating && num14 != c.num14)
        return fail(NBS_ERR_INVALID, "updateParametersInContext: The number of non-excluded exceptions has changed");

    c.N = N;
    c.nS = d.num_subsets;
    c.nSl = c.nS*(c.nS+1)/2;
    c.num14 = num14;
    c.subsets = subsets;
    c.baseQ.assign(d.charges, d.charges + N);
    c.baseSig.assign(d.sigmas, d.sigmas + N);
    c.baseEps.assign(d.epsilons, d.epsilons + N);
    c.nExc = d.num_exceptions;
    c.excPairs.assign(d.exception_particles, d.exception_particles + 2*(size_t) c.nExc);
    c.excParams.assign(d.exception_params, d.exception_params + 3*(size_t) c.nExc);
    c.pOffIdx.assign(d.particle_offset_indices, d.particle_offset_indices + 2*(size_t) d.num_particle_offsets);
    c.pOffScale.assign(d.particle_offset_scales, d.particle_offset_scales + 3*(size_t) d.num_particle_offsets);
    c.eOffIdx.assign(d.exception_offset_indices, d.exception_offset_indices + 2*(size_t) d.num_exception_offsets);
    c.eOffScale.assign(d.exception_offset_scales, d.exception_offset_scales + 3*(size_t) d.num_exception_offsets);
    if (creating) {
        c.nGlobals = d.num_global_params;
        c.globals.assign(c.nGlobals, 0.0);
        c.lambdas.assign(2*(size_t) c.nSl, 1.0);
        c.method = d.method;
        c.cutoff = d.cutoff;
        c.alpha = d.ewald_alpha;
        c.switchDist = d.switching_distance;
        c.rfDielectric = d.rf_dielectric;
        c.useSwitch = d.method != NBS_METHOD_NOCUTOFF && d.use_switching_function != 0;
        c.periodic = d.method == NBS_METHOD_CUTOFF_PERIODIC || d.method == NBS_METHOD_PME || d.method == NBS_METHOD_EWALD || d.method == NBS_METHOD_LJPME;
        c.dispAlpha = d.dispersion_alpha;
        for (int k = 0; k < 3; k++) c.dispGrid[k] = d.dispersion_grid[k];
        if (d.method == NBS_METHOD_LJPME) c.useSwitch = false;      // ReferenceNonbondedSlicingKernels.cpp:174
        for (int k = 0; k < 3; k++) c.ewaldKmax[k] = d.ewald_kmax[k];
        c.cutoffEff = d.cutoff;
        c.excPeriodic = (d.method == NBS_METHOD_NOCUTOFF || d.method == NBS_METHOD_CUTOFF_NONPERIODIC) ? false : d.exceptions_use_periodic != 0;
        for (int k = 0; k < 3; k++) c.grid[k] = d.pme_grid[k];
        c.flags = d.flags;
        c.device = d.device_index;
    }
    c.dispersion.assign(c.nSl, 0.0);
    if (d.dispersion_coefficients)
        for (int s = 0; s < c.nSl; s++) c.dispersion[s] = d.dispersion_coefficients[s];
    c.paramsDirty = true;
    c.paramVersion++;
    return NBS_OK;
}

// computeParameters (:339-392) + upload
static int applyParameters(Context& c) {
    const int N = c.N;
    std::vector<double> q(c.baseQ), sig(c.baseSig), eps(c.baseEps);
    for (size_t i = 0; i < c.pOffIdx.size()/2; i++) {
        const double value = c.globals[c.pOffIdx[2*i]];
        const int index = c.pOffIdx[2*i+1];
        q[index] += value*c.pOffScale[3*i];
        sig[index] += value*c.pOffScale[3*i+1];
        eps[index] += value*c.pOffScale[3*i+2];
    }
    std::vector<float> chargeF(N);
    std::vector<float2> sigEps(N);
    const double sqrtK = std::sqrt(kOne4PiEps0);
    c.subsetQ.assign(c.nS, 0.0);
    c.subsetQ2.assign(c.nS, 0.0);
    for (int i = 0; i < N; i++) {
        chargeF[i] = (float) (q[i]*sqrtK);
        sigEps[i] = make_float2((float) (0.5*sig[i]), (float) (2.0*std::sqrt(eps[i])));
        c.subsetQ[c.subsets[i]] += q[i];
        c.subsetQ2[c.subsets[i]] += q[i]*q[i];
    }
    if (c.ljpme()) {
        // c6 = 8 (sigma/2)^3 (2 sqrt(eps)), ReferenceSlicedLJCoulombIxn.cpp:248, 395-396; self term :212
        std::vector<float> c6F(N);
        std::vector<double> c6D(N);
        c.subsetC6Self.assign(c.nS, 0.0);
        for (int i = 0; i < N; i++) {
            c6D[i] = 8.0*std::pow(0.5*sig[i], 3.0)*(2.0*std::sqrt(eps[i]));
            c6F[i] = (float) c6D[i];
            c.subsetC6Self[c.subsets[i]] += 64.0*std::pow(0.5*sig[i], 6.0)*std::pow(2.0*std::sqrt(eps[i]), 2.0)/12.0;
        }
        NBS_CUDA_CHECK(c.dC6F.ensure(N));
        NBS_CUDA_CHECK(c.dC6D.ensure(N));
        NBS_CUDA_CHECK(cudaMemcpy(c.dC6F.d, c6F.data(), sizeof(float)*N, cudaMemcpyHostToDevice));
        NBS_CUDA_CHECK(cudaMemcpy(c.dC6D.d, c6D.data(), sizeof(double)*N, cudaMemcpyHostToDevice));
    }
    std::vector<double> eq(c.nExc), es(c.nExc), ee(c.nExc);
    std::vector<char> is14(c.nExc, 0);
    for (int i = 0; i < c.nExc; i++) {
        eq[i] = c.excParams[3*i]; es[i] = c.excParams[3*i+1]; ee[i] = c.excParams[3*i+2];
        is14[i] = (eq[i] != 0.0 || ee[i] != 0.0) ? 1 : 0;
    }
    for (size_t i = 0; i < c.eOffIdx.size()/2; i++) {
        const double value = c.globals[c.eOffIdx[2*i]];
        const int index = c.eOffIdx[2*i+1];
        is14[index] = 1;
        eq[index] += value*c.eOffScale[3*i];
        es[index] += value*c.eOffScale[3*i+1];
        ee[index] += value*c.eOffScale[3*i+2];
    }
    std::vector<double4> excParam(c.nExc);
    std::vector<int2> excPair(c.nExc);
    std::vector<int> excSlice(c.nExc);
    for (int i = 0; i < c.nExc; i++) {
        excParam[i] = make_double4(es[i], 4.0*ee[i], kOne4PiEps0*eq[i], is14[i] ? 1.0 : 0.0);
        excPair[i] = make_int2(c.excPairs[2*i], c.excPairs[2*i+1]);
        const int s1 = c.subsets[excPair[i].x], s2 = c.subsets[excPair[i].y];
        excSlice[i] = s1 > s2 ? s1*(s1+1)/2 + s2 : s2*(s2+1)/2 + s1;
    }
    // exclusion lists (every exception is an exclusion, :101-106), CSR over particles, partners sorted
    std::vector<int> exclStart(N+1, 0);
    for (int i = 0; i < c.nExc; i++) { exclStart[excPair[i].x+1]++; exclStart[excPair[i].y+1]++; }
    for (int i = 0; i < N; i++) exclStart[i+1] += exclStart[i];
    std::vector<int> exclList(std::max<size_t>(1, 2*(size_t) c.nExc)), cursor(exclStart.begin(), exclStart.end()-1);
    for (int i = 0; i < c.nExc; i++) {
        exclList[cursor[excPair[i].x]++] = excPair[i].y;
        exclList[cursor[excPair[i].y]++] = excPair[i].x;
    }
    for (int i = 0; i < N; i++) std::sort(exclList.begin() + exclStart[i], exclList.begin() + exclStart[i+1]);

    NBS_CUDA_CHECK(c.dSubset.ensure(N));
    NBS_CUDA_CHECK(c.dChargeF.ensure(N));
    NBS_CUDA_CHECK(c.dSigEps.ensure(N));
    NBS_CUDA_CHECK(c.dCharge.ensure(N));
    NBS_CUDA_CHECK(c.dExclStart.ensure(N+1));
    NBS_CUDA_CHECK(c.dExclList.ensure(exclList.size()));
    NBS_CUDA_CHECK(c.dExcPair.ensure(std::max(1, c.nExc)));
    NBS_CUDA_CHECK(c.dExcParam.ensure(std::max(1, c.nExc)));
    NBS_CUDA_CHECK(c.dExcSlice.ensure(std::max(1, c.nExc)));
    NBS_CUDA_CHECK(cudaMemcpy(c.dSubset.d, c.subsets.data(), sizeof(int)*N, cudaMemcpyHostToDevice));
    NBS_CUDA_CHECK(cudaMemcpy(c.dChargeF.d, chargeF.data(), sizeof(float)*N, cudaMemcpyHostToDevice));
    NBS_CUDA_CHECK(cudaMemcpy(c.dSigEps.d, sigEps.data(), sizeof(float2)*N, cudaMemcpyHostToDevice));
    NBS_CUDA_CHECK(cudaMemcpy(c.dCharge.d, q.data(), sizeof(double)*N, cudaMemcpyHostToDevice));
    if (c.flags & NBS_FLAG_DOUBLE) {
        std::vector<double2> sigEpsD(N);
        for (int i = 0; i < N; i++) sigEpsD[i] = make_double2(0.5*sig[i], 2.0*std::sqrt(eps[i]));
        NBS_CUDA_CHECK(c.dSigEpsD.ensure(N));
        NBS_CUDA_CHECK(cudaMemcpy(c.dSigEpsD.d, sigEpsD.data(), sizeof(double2)*N, cudaMemcpyHostToDevice));
    }
    NBS_CUDA_CHECK(cudaMemcpy(c.dExclStart.d, exclStart.data(), sizeof(int)*(N+1), cudaMemcpyHostToDevice));
    NBS_CUDA_CHECK(cudaMemcpy(c.dExclList.d, exclList.data(), sizeof(int)*exclList.size(), cudaMemcpyHostToDevice));
    if (c.nExc > 0) {
        NBS_CUDA_CHECK(cudaMemcpy(c.dExcPair.d, excPair.data(), sizeof(int2)*c.nExc, cudaMemcpyHostToDevice));
        NBS_CUDA_CHECK(cudaMemcpy(c.dExcParam.d, excParam.data(), sizeof(double4)*c.nExc, cudaMemcpyHostToDevice));
        NBS_CUDA_CHECK(cudaMemcpy(c.dExcSlice.d, excSlice.data(), sizeof(int)*c.nExc, cudaMemcpyHostToDevice));
    }
    c.paramsDirty = false;
    c.listEpoch++;                 // the sorted per-atom records of a re-used list hold the old parameters
    return NBS_OK;
}

// |DFT of the order-5 cardinal B-spline sampled at the integers|^2 along one axis (what the reference tabulates with
// an O(n^2) cosine/sine sum, ReferencePME.cpp:88-183).  The samples are M5(1..4) = 1/24, 11/24, 11/24, 1/24 --
// symmetric about 2.5 -- so the transform is a phase factor times (cos(3t/2) + 11 cos(t/2))/12 with t = 2 pi k/n and
// the modulus has a closed form.  It vanishes only at t = pi (k = n/2, even n); like the reference (:170-176) such
// an entry is replaced by the mean of its two neighbours.
static void splineModuli(int n, double* out) {
    for (int k = 0; k < n; k++) {
        const double halfT = kPi*k/n;
        const double amplitude = (std::cos(3.0*halfT) + 11.0*std::cos(halfT))/12.0;
        out[k] = amplitude*amplitude;
    }
    for (int k = 0; k < n; k++)
        if (out[k] < 1.0e-7) out[k] = 0.5*(out[(k + n - 1) % n] + out[(k + 1) % n]);
}

// LJPME runs the PME chain twice: charges on the (alpha, grid) set of setUsePME, then C6 coefficients on the set
// of setUseLJPME (ReferenceSlicedLJCoulombIxn.cpp:229-253).  The chain reads its set from the Context's main
// fields; this swaps the two sets (host-side only; work buffers are shared and sized for the larger grid).
void swapPmeTables(Context& c) {
    std::swap(c.alpha, c.dispAlpha);
    for (int k = 0; k < 3; k++) std::swap(c.grid[k], c.dispGrid[k]);
    for (int k = 0; k < 6; k++) std::swap(c.etermBox[k], c.etermBoxDisp[k]);
    std::swap(c.dModuli, c.dModuliDisp);
    std::swap(c.dTwiddle, c.dTwiddleDisp);
    std::swap(c.dTwiddleD, c.dTwiddleDDisp);
    std::swap(c.dEterm, c.dEtermDisp);
    std::swap(c.dEtermD, c.dEtermDDisp);
    std::swap(c.hModuli, c.hModuliDisp);
    c.dispersionPass = !c.dispersionPass;
}

int uploadPmeTables(Context& c) {
    const int nx = c.grid[0], ny = c.grid[1], nz = c.grid[2];
    const int total = nx + ny + nz;
    c.hModuli.assign(total, 0.0);
    splineModuli(nx, c.hModuli.data());
    splineModuli(ny, c.hModuli.data() + nx);
    splineModuli(nz, c.hModuli.data() + nx + ny);
    std::vector<float2> tw(total);
    std::vector<double2> twD(total);
    int off = 0;
    for (int n : {nx, ny, nz}) {
        for (int k = 0; k < n; k++) {
            double ang = -2.0*kPi*k/n;
            tw[off+k] = make_float2((float) cos(ang), (float) sin(ang));
            twD[off+k] = make_double2(cos(ang), sin(ang));
        }
        off += n;
    }
    NBS_CUDA_CHECK(c.dModuli.ensure(total));
    NBS_CUDA_CHECK(c.dTwiddle.ensure(total));
    NBS_CUDA_CHECK(cudaMemcpy(c.dModuli.d, c.hModuli.data(), sizeof(double)*total, cudaMemcpyHostToDevice));
    NBS_CUDA_CHECK(cudaMemcpy(c.dTwiddle.d, tw.data(), sizeof(float2)*total, cudaMemcpyHostToDevice));
    NBS_CUDA_CHECK(c.dTwiddleD.ensure(total));
    NBS_CUDA_CHECK(cudaMemcpy(c.dTwiddleD.d, twD.data(), sizeof(double2)*total, cudaMemcpyHostToDevice));
    const size_t G = (size_t) nx*ny*nz, Gh = (size_t) nx*ny*(nz/2+1);
    NBS_CUDA_CHECK(c.dGrid.ensure(G*c.nS));
    NBS_CUDA_CHECK(c.dGridC.ensure(Gh*c.nS));
    NBS_CUDA_CHECK(c.dPot.ensure(((c.flags & NBS_FLAG_DOUBLE) ? 2 : 1)*G*c.nS));      // (double-precision mode: a double potential grid)
    return NBS_OK;
}

// Piecewise polynomial table of f(s) = erfc(alpha sqrt(s))/sqrt(s) for the double-precision pair energies
// (k_pair.cu pairStep): degree-4 interpolation at Chebyshev nodes on every interval [2^e (1 + m/128),
// 2^e (1 + (m+1)/128)), e = -7 .., in the variable d = (s - centre)/width; c0 in double, c1..c4 in single precision
// (ERFC_TAB_* in nbs_internal.h).  Built from libm's long-double erfc.
// (host only: returns the number of rows; `tab` = c0[rows] doubles, then the rows' float4 coefficient sets)
static int computeErfcTable(double alpha, double cutoff, std::vector<double>& tab) {
    const int M = 1 << ERFC_TAB_PER_OCTAVE_LOG2;
    const int eMax = std::min(ERFC_TAB_MAX_ROWS/M - 8, std::max(-6, (int) std::floor(std::log2(cutoff*cutoff)) + 1));
    const int rows = (eMax + 7 + 1)*M;
    constexpr int D = 4;
    tab.assign((size_t) rows*3, 0.0);                            // c0[rows], then float4[rows] = 2 doubles per row
    float* coef = reinterpret_cast<float*>(tab.data() + rows);
    for (int e = -7; e <= eMax; e++)
        for (int m = 0; m < M; m++) {
            const long double lo = std::ldexp(1.0L + m/(long double) M, e), w = std::ldexp(1.0L/M, e);
            const long double center = lo + w/2;
            long double A[D + 1][D + 2];
            for (int k = 0; k <= D; k++) {
                const long double dn = 0.5L*std::cos(3.14159265358979323846264338327950288L*(2*k + 1)/(2.0L*(D + 1)));
                const long double sv = center + dn*w;
                const long double r = std::sqrt(sv);
                long double pw = 1;
                for (int j = 0; j <= D; j++) { A[k][j] = pw; pw *= dn; }
                A[k][D + 1] = std::erfc((long double) alpha*r)/r;
            }
            for (int col = 0; col <= D; col++) {                 // Gaussian elimination, partial pivoting
                int piv = col;
                for (int r = col + 1; r <= D; r++) if (std::fabs(A[r][col]) > std::fabs(A[piv][col])) piv = r;
                for (int j = 0; j <= D + 1; j++) std::swap(A[col][j], A[piv][j]);
                for (int r = 0; r <= D; r++) {
                    if (r == col) continue;
                    const long double f = A[r][col]/A[col][col];
                    for (int j = col; j <= D + 1; j++) A[r][j] -= f*A[col][j];
                }
            }
            const int row = (e + 7)*M + m;
            tab[row] = (double) (A[0][D + 1]/A[0][0]);
            for (int k = 1; k <= D; k++) coef[4*(size_t) row + k - 1] = (float) (A[k][D + 1]/A[k][k]);
        }
    return rows;
}

static int buildErfcTable(Context& c) {
    std::vector<double> tab;
    c.erfcRows = computeErfcTable(c.alpha, c.cutoff, tab);
    NBS_CUDA_CHECK(c.dErfcTab.ensure(tab.size()));
    NBS_CUDA_CHECK(cudaMemcpy(c.dErfcTab.d, tab.data(), sizeof(double)*tab.size(), cudaMemcpyHostToDevice));
    return NBS_OK;
}

// cell geometry for this box: columns whose 32-atom blocks are roughly cubic, fine z-bins for sorting
static int setupGeometry(Context& c, const double L[3], const double origin[3], const double tilt[3]) {
    CellGeom& g = c.geom;
    for (int k = 0; k < 3; k++) g.tilt[k] = tilt[k];
    g.triclinic = tilt[0] != 0 || tilt[1] != 0 || tilt[2] != 0;
    g.shiftB = std::llround(tilt[0]/L[0]*4294967296.0);
    g.shiftCx = std::llround(tilt[1]/L[0]*4294967296.0);
    g.shiftCy = std::llround(tilt[2]/L[1]*4294967296.0);
    const double volume = L[0]*L[1]*L[2];
    const double density = c.N/volume;
    // columns narrower than a cubic 32-atom block (factor 0.8): the i-blocks get taller, their 4-atom clusters
    // closer to cubes, and the pair kernel's cluster masks prune more (profiles/README.md)
    static const double sideScale = getenv("NBS_COL_SIDE_SCALE") ? atof(getenv("NBS_COL_SIDE_SCALE")) : 0.8;
    double side = sideScale*std::cbrt(32.0/density);
    side = std::max(side, 0.3*c.cutoffEff);
    for (int k = 0; k < 3; k++) {
        g.origin[k] = origin[k];
        g.box[k] = L[k];
        g.invBox[k] = 1.0/L[k];
        g.scale[k] = (float) (L[k]/4294967296.0);
    }
    g.ncx = std::max(1, std::min(512, (int) std::lround(L[0]/side)));
    g.ncy = std::max(1, std::min(512, (int) std::lround(L[1]/side)));
    g.nzb = std::max(1, std::min(8192, (int) std::ceil(L[2]/(side/8))));
    while ((long long) g.ncx*g.ncy*g.nzb > 8000000LL && g.nzb > 1) g.nzb /= 2;
    g.nCols = g.ncx*g.ncy;
    g.nBins = g.nCols*g.nzb;
    g.colW[0] = (float) (L[0]/g.ncx);
    g.colW[1] = (float) (L[1]/g.ncy);
    g.binH = (float) (L[2]/g.nzb);
    const int N = c.N;
    c.Npad = ((N + 31)/32)*32 + 32;
    c.maxBlocks = N/32 + g.nCols + 1;
    c.maxLocalBlocks = std::max(1, ((c.maxBlocks + c.blockPeriod - 1)/c.blockPeriod)*c.blockWidth);
    NBS_CUDA_CHECK(c.dFix.ensure(N));
    NBS_CUDA_CHECK(c.dFixBuild.ensure(c.Npad));
    NBS_CUDA_CHECK(c.dBinCount.ensure(g.nBins + 2));
    NBS_CUDA_CHECK(c.dBinStart.ensure(g.nBins + 2));
    NBS_CUDA_CHECK(c.dBinCursor.ensure(g.nBins + 2));
    NBS_CUDA_CHECK(c.dSortedToOrig.ensure(N));
    NBS_CUDA_CHECK(c.dOrigToSorted.ensure(N));
    NBS_CUDA_CHECK(c.dPosq.ensure(c.Npad));
    NBS_CUDA_CHECK(c.dPar.ensure(c.Npad));
    NBS_CUDA_CHECK(c.dQ64.ensure(c.Npad));
    NBS_CUDA_CHECK(c.dExclRange.ensure(c.Npad));
    NBS_CUDA_CHECK(c.dColBlockStart.ensure(g.nCols + 2));
    NBS_CUDA_CHECK(c.dBlkFirst.ensure(c.maxBlocks));
    NBS_CUDA_CHECK(c.dBlkCount.ensure(c.maxBlocks));
    NBS_CUDA_CHECK(c.dBlkLo.ensure(c.maxBlocks));
    NBS_CUDA_CHECK(c.dBlkHi.ensure(c.maxBlocks));
    NBS_CUDA_CHECK(c.dJCount.ensure(c.maxLocalBlocks));
    NBS_CUDA_CHECK(c.dXCount.ensure(c.maxLocalBlocks));
    NBS_CUDA_CHECK(c.dJList.ensure((size_t) c.maxLocalBlocks*c.capJ));
    NBS_CUDA_CHECK(c.dXList.ensure((size_t) c.maxLocalBlocks*c.capX));
    NBS_CUDA_CHECK(c.dXMask.ensure((size_t) c.maxLocalBlocks*c.capX));
    NBS_CUDA_CHECK(c.dGmJ.ensure((size_t) c.maxLocalBlocks*(c.capJ/32)));
    NBS_CUDA_CHECK(c.dGmX.ensure((size_t) c.maxLocalBlocks*(c.capX/32)));
    // work items: enough warps' worth of items to balance 148 SMs x 16 warps on small systems, larger
    // chunks (fewer i-force flushes) on big ones
    // (sized by the atoms whose i-blocks this rank works on.  Measured, pair kernel with energies: C3 122 / 120 / 129 us at
    // 1 / 2 / 3 tiles per item, C4 381 / 341 / 347 us at 1 / 2 / 4, C5 on 8 ranks -- 133 k atoms' worth of blocks per rank --
    // 0.546 / 0.512 ms at 1 / 8: profiles/r02_chunk_tiles_sweep.log)
    const long long localN = c.nRanks > 1 ? (long long) N*std::max(1, c.blockWidth)/std::max(1, c.blockPeriod) : N;
    c.chunkTiles = localN < 40000 ? 1 : (localN < 110000 ? 2 : (localN < 250000 ? 4 : 8));
    if (const char* env = getenv("NBS_CHUNK_TILES")) c.chunkTiles = std::max(1, atoi(env));     // tuning experiments
    NBS_CUDA_CHECK(c.dItems.ensure((size_t) c.maxLocalBlocks*(((c.capJ + c.capX)/32 + c.chunkTiles - 1)/c.chunkTiles + 1)));
    // PME from particle-order coordinates while the grids are comfortably L2-resident (scattered access is free
    // there); large systems keep the cell-sorted order for locality
    // (the Ewald sum and the LJPME chains always work from the particle-order fixed-point coordinates)
    c.pmeUnsorted = (c.method == NBS_METHOD_PME && N < 300000 && !(c.flags & NBS_FLAG_SORTED_PME)) || c.method == NBS_METHOD_EWALD || c.method == NBS_METHOD_LJPME;
    NBS_CUDA_CHECK(c.dForce.ensure(6*(size_t) c.Npad));
    return NBS_OK;
}

static int checkDevice(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(NBS_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= count) return fail(NBS_ERR_CUDA, "device_index out of range");
    cudaDeviceProp prop;
    NBS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(NBS_ERR_CUDA, std::string("device '") + prop.name + "' is not sm_100: this library is built for B200 only");
    NBS_CUDA_CHECK(cudaSetDevice(device));
    return NBS_OK;
}


static void releaseAll(Context& c) {
    timerReset(c);
    c.dSubset.release(); c.dChargeF.release(); c.dSigEps.release(); c.dCharge.release(); c.dSigEpsD.release(); c.dLamD.release();
    c.dExclStart.release(); c.dExclList.release(); c.dExcPair.release(); c.dExcParam.release(); c.dExcSlice.release();
    c.dPosIn.release(); c.dForceOut.release(); c.dFix.release(); c.dFixBuild.release(); c.dBinCount.release(); c.dBinStart.release();
    c.dBinCursor.release(); c.dScanTmp.release(); c.dSortedToOrig.release(); c.dOrigToSorted.release();
    c.dPosq.release(); c.dPar.release(); c.dQ64.release(); c.dColBlockStart.release(); c.dBlkFirst.release(); c.dBlkCount.release();
    c.dBlkLo.release(); c.dBlkHi.release(); c.dExclRange.release(); c.dJList.release(); c.dJCount.release();
    c.dXList.release(); c.dXCount.release(); c.dXMask.release(); c.dGmJ.release(); c.dGmX.release(); c.dCounters.release(); c.dForce.release(); c.dItems.release();
    c.dEnergy.release(); c.dGrid.release(); c.dGridFixed.release(); c.dGridC.release(); c.dEterm.release(); c.dModuli.release();
    c.dPot.release(); c.dEtermD.release(); c.dTwiddleD.release(); c.dErfcTab.release();
    c.dEwaldK.release(); c.dEwaldSums.release(); c.dEwaldMixed.release();
    c.dC6F.release(); c.dC6D.release(); c.dEtermDisp.release(); c.dEtermDDisp.release(); c.dModuliDisp.release();
    c.dTwiddleDisp.release(); c.dTwiddleDDisp.release();
    c.dTwiddle.release(); c.dPairStats.release(); c.dPairDump.release();
    if (c.hCounters) cudaFreeHost(c.hCounters);
    if (c.hEnergy) cudaFreeHost(c.hEnergy);
    if (c.hForce) cudaFreeHost(c.hForce);
    c.hCounters = nullptr; c.hEnergy = nullptr; c.hForce = nullptr;
    for (int k = 0; k < 2; k++) {
        if (c.graphExecs[k]) cudaGraphExecDestroy(c.graphExecs[k]);
        c.graphExecs[k] = nullptr;
    }
    if (c.ownStream) cudaStreamDestroy(c.ownStream);
    c.ownStream = nullptr;
    if (c.auxStream) cudaStreamDestroy(c.auxStream);
    if (c.evAuxFork) cudaEventDestroy(c.evAuxFork);
    if (c.evAuxDone) cudaEventDestroy(c.evAuxDone);
    if (c.evPlaced) cudaEventDestroy(c.evPlaced);
    if (c.evExclDone) cudaEventDestroy(c.evExclDone);
    c.evPlaced = nullptr; c.evExclDone = nullptr;
    c.auxStream = nullptr; c.evAuxFork = nullptr; c.evAuxDone = nullptr;
    if (c.directStream) cudaStreamDestroy(c.directStream);
    if (c.evSorted) cudaEventDestroy(c.evSorted);
    if (c.evDirectDone) cudaEventDestroy(c.evDirectDone);
    c.directStream = nullptr; c.evSorted = nullptr; c.evDirectDone = nullptr;
    for (int k = 0; k < 3*NBS_MAX_RANKS; k++) {
        if (c.peerOpened[k]) cudaIpcCloseMemHandle(c.peerOpened[k]);
        c.peerOpened[k] = nullptr;
    }
    c.dMailbox.release();
    if (c.hTimedOut) cudaFreeHost(c.hTimedOut);
    c.hTimedOut = nullptr;
}

} // namespace nbs

using namespace nbs;

struct nbs_context { Context c; };

extern "C" {

int nbs_abi_version(void) { return NBS_ABI_VERSION; }

const char* nbs_last_error(void) { return gLastError.c_str(); }

int nbs_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int nbs_create(const nbs_system_desc* desc, nbs_context** out) {
    if (!desc || !out) return fail(NBS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (desc->struct_size != (int32_t) sizeof(nbs_system_desc)) return fail(NBS_ERR_INVALID, "nbs_system_desc.struct_size mismatch");
    switch (desc->method) {
        case NBS_METHOD_PME: case NBS_METHOD_CUTOFF_PERIODIC: case NBS_METHOD_EWALD: case NBS_METHOD_LJPME: break;
        case NBS_METHOD_NOCUTOFF: case NBS_METHOD_CUTOFF_NONPERIODIC: break;
        default: return fail(NBS_ERR_INVALID, "illegal nonbonded method");
    }
    if (desc->method != NBS_METHOD_NOCUTOFF && desc->cutoff <= 0) return fail(NBS_ERR_INVALID, "cutoff must be positive");
    int status = checkDevice(desc->device_index);
    if (status != NBS_OK) return status;
    nbs_context* ctx = new nbs_context();
    Context& c = ctx->c;
    status = readDescription(c, *desc, true);
    if (status != NBS_OK) { delete ctx; return status; }
    if (c.ljpme() && (c.dispAlpha <= 0 || c.dispGrid[0] < PME_ORDER+1 || c.dispGrid[1] < PME_ORDER+1 || c.dispGrid[2] < PME_ORDER+1)) {
        delete ctx;
        return fail(NBS_ERR_INVALID, "LJPME needs dispersion_alpha > 0 and a dispersion grid of at least 6 points per dimension");
    }
    if (c.usesPmeGrid()) {
        if (c.alpha <= 0 || c.grid[0] < PME_ORDER+1 || c.grid[1] < PME_ORDER+1 || c.grid[2] < PME_ORDER+1) {
            delete ctx;
            return fail(NBS_ERR_INVALID, "PME needs ewald_alpha > 0 and a grid of at least 6 points per dimension");
        }
    }
    if (c.method == NBS_METHOD_EWALD) {
        // ReferenceSlicedLJCoulombIxn.cpp:270-271
        if (c.alpha <= 0 || std::max(c.ewaldKmax[0], std::max(c.ewaldKmax[1], c.ewaldKmax[2])) < 1) {
            delete ctx;
            return fail(NBS_ERR_INVALID, "kmax for Ewald summation < 1");
        }
    }
    c.capJ = 2048;
    { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, c.device) == cudaSuccess) c.numSMs = prop.multiProcessorCount; }
    c.ownLo = 0; c.ownHi = c.nS;
    c.capX = 256;
    c.profiling = (c.flags & NBS_FLAG_PROFILE) != 0;
    // neighbour-list re-use: on by default for the periodic cutoff methods (nbs_set_list_skin changes it)
    c.skin = (c.periodic && !(c.flags & NBS_FLAG_NO_LIST_REUSE)) ? 0.07 : 0.0;
    if (const char* env = getenv("NBS_LIST_SKIN")) { if (c.periodic && !(c.flags & NBS_FLAG_NO_LIST_REUSE)) c.skin = std::max(0.0, atof(env)); }
    cudaError_t e = cudaSuccess;
    if ((e = c.dCounters.ensure(16)) != cudaSuccess || (e = c.dEnergy.ensure(ENERGY_WORDS)) != cudaSuccess ||
        (e = c.dPairStats.ensure(4)) != cudaSuccess || (e = c.dPairDump.ensure(1)) != cudaSuccess ||
        (e = cudaMallocHost((void**) &c.hCounters, 16*sizeof(int))) != cudaSuccess ||
        (e = cudaMallocHost((void**) &c.hEnergy, ENERGY_WORDS*sizeof(double))) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c.directStream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c.auxStream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c.evAuxFork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c.evAuxDone, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c.evPlaced, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c.evExclDone, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c.evSorted, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c.evDirectDone, cudaEventDisableTiming)) != cudaSuccess) {
        releaseAll(c);
        delete ctx;
        return fail(NBS_ERR_CUDA, std::string("allocation failed: ") + cudaGetErrorString(e));
    }
    if (c.usesPmeGrid()) {
        status = uploadPmeTables(c);
        if (status == NBS_OK && c.ljpme()) {
            swapPmeTables(c);
            status = uploadPmeTables(c);
            swapPmeTables(c);
        }
        if (status == NBS_OK) status = buildErfcTable(c);
        if (status != NBS_OK) { releaseAll(c); delete ctx; return status; }
    }
    if (c.method == NBS_METHOD_EWALD) {
        status = uploadEwaldVectors(c);
        if (status == NBS_OK) status = buildErfcTable(c);
        if (status != NBS_OK) { releaseAll(c); delete ctx; return status; }
    }
    *out = ctx;
    return NBS_OK;
}

int nbs_destroy(nbs_context* ctx) {
    if (!ctx) return NBS_OK;
    cudaSetDevice(ctx->c.device);
    releaseAll(ctx->c);
    delete ctx;
    return NBS_OK;
}

int nbs_update_parameters(nbs_context* ctx, const nbs_system_desc* desc) {
    if (!ctx || !desc) return fail(NBS_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->c.device);
    ctx->c.listEpoch++;
    return readDescription(ctx->c, *desc, false);
}

int nbs_set_lambdas(nbs_context* ctx, const double* lambdas) {
    if (!ctx || !lambdas) return fail(NBS_ERR_INVALID, "null argument");
    ctx->c.lambdas.assign(lambdas, lambdas + 2*(size_t) ctx->c.nSl);
    ctx->c.paramVersion++;
    return NBS_OK;
}

int nbs_set_global_parameters(nbs_context* ctx, const double* values) {
    if (!ctx) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (c.nGlobals == 0) return NBS_OK;
    if (!values) return fail(NBS_ERR_INVALID, "null argument");
    bool changed = false;
    for (int k = 0; k < c.nGlobals; k++)
        if (c.globals[k] != values[k]) { c.globals[k] = values[k]; changed = true; }
    // only parameters that feed an offset change the per-particle data (CommonNonbondedSlicingKernels.cpp:1152-1163)
    if (changed && (!c.pOffIdx.empty() || !c.eOffIdx.empty())) { c.paramsDirty = true; c.paramVersion++; }
    return NBS_OK;
}

// ---- one evaluation, in three phases --------------------------------------------------------------
// A single-rank evaluation (nbs_execute) runs them back to back.  A multi-rank evaluation (one process
// per GPU) interleaves two exchanges that the caller performs on the same stream:
//   begin    : sort, neighbour list + pair kernel + exceptions for this rank's i-blocks (on an internal
//              stream, concurrent with PME), spreading + forward z/y FFT of this rank's subset grids
//   [exchange: every rank that owns grids receives the other subsets' half spectra]
//   convolve : fused x pass (cross-subset products per slice, lambda mixing), inverse FFT, force gather
//   [exchange: all-reduce of the fixed-point force accumulators and of the slice-energy table]
//   finish   : forces to the caller's layout, energies to the host, host-side constants
static int validateExec(Context& c, const nbs_exec_args* args) {
    if (args->struct_size != (int32_t) sizeof(nbs_exec_args)) return fail(NBS_ERR_INVALID, "nbs_exec_args.struct_size mismatch");
    const double* box = args->box;
    if (c.periodic) {
        // OpenMM's reduced form (Context::setPeriodicBoxVectors [external] enforces it): a = (ax, 0, 0),
        // b = (bx, by, 0), c = (cx, cy, cz), |bx| <= ax/2, |cx| <= ax/2, |cy| <= by/2
        if (box[1] != 0 || box[2] != 0 || box[5] != 0)
            return fail(NBS_ERR_INVALID, "First periodic box vector must be parallel to x, second must be in the x-y plane.");
        if (box[0] <= 0 || box[4] <= 0 || box[8] <= 0 ||
            std::fabs(box[3]) > 0.5*box[0]*(1 + 1e-12) || std::fabs(box[6]) > 0.5*box[0]*(1 + 1e-12) || std::fabs(box[7]) > 0.5*box[4]*(1 + 1e-12))
            return fail(NBS_ERR_INVALID, "Periodic box vectors must be in reduced form.");
        if (c.method == NBS_METHOD_EWALD && (box[3] != 0 || box[6] != 0 || box[7] != 0))     // SlicedNonbondedForceImpl.cpp:112-113
            return fail(NBS_ERR_UNSUPPORTED, "SlicedNonbondedForce: Ewald is not supported with non-rectangular boxes.  Use PME instead.");
        // ReferenceNonbondedSlicingKernels.cpp:200-204
        const double minAllowedSize = 1.999999*c.cutoff;
        if (box[0] < minAllowedSize || box[4] < minAllowedSize || box[8] < minAllowedSize)
            return fail(NBS_ERR_BOX, "The periodic box size has decreased to less than twice the nonbonded cutoff.");
    }
    if (!args->positions) return fail(NBS_ERR_INVALID, "positions is null");
    if (args->positions_format != NBS_POS_F64_XYZ && args->positions_space != NBS_MEM_DEVICE)
        return fail(NBS_ERR_INVALID, "xyzw positions must be device memory");
    if (args->positions_format < 0 || args->positions_format > NBS_POS_F64_XYZW)
        return fail(NBS_ERR_INVALID, "illegal positions_format");
    if (args->forces_format == NBS_FORCE_I64_FIXED && (args->forces_space != NBS_MEM_DEVICE || args->padded_num_atoms < c.N))
        return fail(NBS_ERR_INVALID, "fixed-point forces need device memory and padded_num_atoms >= num_particles");
    if (args->atom_index && args->positions_space == NBS_MEM_HOST)
        return fail(NBS_ERR_INVALID, "atom_index requires device-resident positions");
    return NBS_OK;
}

// The stream an evaluation is enqueued on: the caller's, or -- when the caller passes the legacy default
// stream, which cannot be captured into a graph -- a blocking stream of our own (blocking streams order
// themselves after work already queued on the legacy default stream, and every evaluation ends with a
// synchronisation, so the caller sees the same ordering).
static cudaStream_t workStream(Context& c, const nbs_exec_args* args) {
    if (args->stream != nullptr || c.nRanks != 1) return (cudaStream_t) args->stream;
    if (!c.ownStream && cudaStreamCreate(&c.ownStream) != cudaSuccess) { cudaGetLastError(); c.ownStream = nullptr; }
    return c.ownStream;
}

// Decides whether this evaluation re-uses the sort order and the neighbour list of an earlier one.  The list was
// built with cutoff + skin; it stays complete while no atom has moved more than skin/2 since, which k_reprep
// measures on the device DURING the evaluation -- phaseComplete then either accepts the result or has the
// evaluation redone with a fresh list (and asks for a fresh list ahead of time when the next step would
// probably exceed the limit).
static void planListReuse(Context& c, const nbs_exec_args* args) {
    const bool direct = args->include_direct != 0;
    c.reuseNow = c.skin > 0 && c.periodic && c.listValid && !c.forceRebuild && !c.paramsDirty &&
                 std::memcmp(c.listBox, args->box, sizeof(double)*9) == 0 && c.listEpochAtBuild == c.listEpoch &&
                 c.listAllocEpoch == gAllocEpoch && c.listCapJ == c.capJ && (!direct || c.listHasDirect);
}

static int phaseBegin(Context& c, const nbs_exec_args* args, bool planned = false) {
    int status = validateExec(c, args);
    if (status != NBS_OK) return status;
    if (c.paramsDirty) { NBS_CUDA_CHECK(cudaSetDevice(c.device)); if ((status = applyParameters(c)) != NBS_OK) return status; }
    if (!planned) planListReuse(c, args);
    c.evalCount++;
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    const double* box = args->box;
    c.stream = workStream(c, args);
    cudaStream_t st = c.stream;
    const bool pme = c.ewaldDirect();
    c.phaseEnergy = args->slice_energies != nullptr;
    c.phaseDirect = args->include_direct != 0;
    c.phaseRecip = args->include_reciprocal != 0 && pme;
    c.phase = 0;
    const int N = c.N;
    timerReset(c);
    timerMark(c, "begin");
    // positions
    PosInput in;
    in.format = args->positions_format;
    in.atomIndex = args->atom_index;
    in.pos64out = nullptr;
    const double* dPos64 = nullptr;
    if (args->positions_space == NBS_MEM_HOST) {
        NBS_CUDA_CHECK(c.dPosIn.ensure(3*(size_t) N));
        NBS_CUDA_CHECK(cudaMemcpyAsync(c.dPosIn.d, args->positions, sizeof(double)*3*N, cudaMemcpyHostToDevice, st));
        in.ptr = c.dPosIn.d;
        dPos64 = c.dPosIn.d;
    }
    else {
        in.ptr = args->positions;
        if (args->positions_format == NBS_POS_F64_XYZ && !args->atom_index) dPos64 = (const double*) args->positions;
        else if (c.nExc > 0 && c.phaseDirect) {     // particle-ordered double copy for the exception kernel
            NBS_CUDA_CHECK(c.dPosIn.ensure(3*(size_t) N));
            in.pos64out = c.dPosIn.d;
            dPos64 = c.dPosIn.d;
        }
    }
    {
        // the box the cell grid and the fixed-point coordinates live in: the periodic box, or -- NoCutoff /
        // CutoffNonPeriodic -- a virtual box twice the size of the system around its bounding box, in which no
        // coordinate difference ever wraps (the wrapped integer difference is then the plain difference)
        double L[3] = {box[0], box[4], box[8]}, origin[3] = {0, 0, 0};
        if (!c.periodic) {
            float bb[6];
            if (args->positions_space == NBS_MEM_HOST) {
                const double* p = (const double*) args->positions;
                for (int d = 0; d < 3; d++) { bb[d] = 3.0e38f; bb[3+d] = -3.0e38f; }
                for (int i = 0; i < N; i++)
                    for (int d = 0; d < 3; d++) { bb[d] = std::min(bb[d], (float) p[3*i+d]); bb[3+d] = std::max(bb[3+d], (float) p[3*i+d]); }
            }
            else if ((status = launchBBox(c, in, bb)) != NBS_OK) return status;
            double diag2 = 0;
            for (int d = 0; d < 3; d++) {
                const double extent = std::max(1.0e-3, (double) bb[3+d] - (double) bb[d]);
                L[d] = 2.0*extent + 4.0;                          // >= 2 x (extent + margin): differences stay below L/2
                origin[d] = 0.5*((double) bb[d] + (double) bb[3+d]) - 0.5*L[d];
                diag2 += extent*extent;
            }
            c.cutoffEff = c.method == NBS_METHOD_NOCUTOFF ? std::sqrt(diag2) + 1.0 : c.cutoff;
        }
        // triclinic box a = (ax, 0, 0), b = (bx, by, 0), c = (cx, cy, cz): the cell grid lives in the brick ax x by x cz
        const double tilt[3] = {c.periodic ? box[3] : 0.0, c.periodic ? box[6] : 0.0, c.periodic ? box[7] : 0.0};
        if ((status = setupGeometry(c, L, origin, tilt)) != NBS_OK) return status;
    }
    NBS_CUDA_CHECK(cudaMemsetAsync(c.dForce.d, 0, sizeof(unsigned long long)*(c.pmeUnsorted ? 6 : 3)*c.Npad, st));
    // (the slice-energy table and the counters are zeroed by k_prep)
    timerMark(c, "h2d_zero");
    if ((status = (c.reuseNow ? launchReprep(c, in) : launchPrep(c, in))) != NBS_OK) return status;
    // The cell sort, the neighbour list and direct space run on their own stream, concurrently with the PME
    // chain on `st` (serial when profiling).  With particle-order PME the fork is right here, after k_prep;
    // otherwise PME needs the sorted records and the fork comes after the sort.
    const bool overlap = !c.profiling && c.phaseRecip && c.directStream != nullptr;
    const bool forkEarly = overlap && c.pmeUnsorted;
    c.directOverlapped = overlap;
    if (forkEarly) {
        NBS_CUDA_CHECK(cudaEventRecord(c.evSorted, st));
        NBS_CUDA_CHECK(cudaStreamWaitEvent(c.directStream, c.evSorted, 0));
        c.stream = c.directStream;
    }
    // (only when this rank builds lists: launchBuildLists is where the side stream joins again)
    c.sideFork = overlap && c.phaseDirect && c.auxStream != nullptr && c.blockWidth != 0;
    status = c.reuseNow ? NBS_OK : launchSortRest(c);
    if (status == NBS_OK && c.phaseDirect) {
        if (overlap && !forkEarly) {
            NBS_CUDA_CHECK(cudaEventRecord(c.evSorted, st));
            NBS_CUDA_CHECK(cudaStreamWaitEvent(c.directStream, c.evSorted, 0));
            c.stream = c.directStream;
        }
        // exceptions / exclusion corrections only need the sorted order: a third stream runs them beside the list build
        const bool forkBonded = overlap && c.nExc > 0 && dPos64 && c.auxStream != nullptr && c.stream == c.directStream;
        if (forkBonded) {
            cudaEventRecord(c.evAuxFork, c.directStream);
            cudaStreamWaitEvent(c.auxStream, c.evAuxFork, 0);
            c.stream = c.auxStream;
            status = launchBonded(c, dPos64, c.periodic);
            cudaEventRecord(c.evAuxDone, c.auxStream);
            c.stream = c.directStream;
        }
        if (status == NBS_OK && !c.reuseNow) status = launchBuildLists(c);
        if (status == NBS_OK) status = launchPairs(c, c.phaseEnergy, 0);
        if (status == NBS_OK && c.nExc > 0 && !forkBonded) {
            if (!dPos64) status = fail(NBS_ERR_UNSUPPORTED, "exceptions need double-precision positions in this version");
            else status = launchBonded(c, dPos64, c.periodic);
        }
        if (forkBonded) cudaStreamWaitEvent(c.directStream, c.evAuxDone, 0);
    }
    if (overlap) {
        if (c.stream == c.directStream) cudaEventRecord(c.evDirectDone, c.directStream);
        else c.directOverlapped = false;                        // nothing was forked (no direct space, late fork)
        c.stream = st;
        if (status != NBS_OK && c.directOverlapped) cudaStreamWaitEvent(st, c.evDirectDone, 0);
    }
    if (status != NBS_OK) return status;
    if (c.phaseRecip && (status = (c.method == NBS_METHOD_EWALD ? launchEwald(c, c.phaseEnergy) : launchPme(c, c.phaseEnergy, 0))) != NBS_OK) {
        if (c.directOverlapped) cudaStreamWaitEvent(st, c.evDirectDone, 0);
        return status;
    }
    c.phase = 1;
    return NBS_OK;
}

static int phaseConvolve(Context& c, const nbs_exec_args* args) {
    if (c.phase != 1) return fail(NBS_ERR_INVALID, "nbs_execute_convolve called without nbs_execute_begin");
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    cudaStream_t st = c.stream;
    int status = NBS_OK;
    if (c.phaseRecip && c.usesPmeGrid()) status = launchPme(c, c.phaseEnergy, 1);
    if (status == NBS_OK && c.phaseRecip && c.ljpme()) {
        // dispersion chain (:241-253): same kernels, the C6 coefficients as "charges", the dispersion set of tables;
        // it reuses the work grids, so it follows the Coulomb chain on the same stream
        swapPmeTables(c);
        status = launchPme(c, c.phaseEnergy, 0);
        if (status == NBS_OK) status = launchPme(c, c.phaseEnergy, 1);
        swapPmeTables(c);
    }
    if (c.directOverlapped) NBS_CUDA_CHECK(cudaStreamWaitEvent(st, c.evDirectDone, 0));   // join the direct-space stream
    if (status != NBS_OK) { c.phase = 0; return status; }
    c.phase = 2;
    return NBS_OK;
}

// Enqueues the output stage: forces to the caller's layout, energies and counters to pinned host memory.
static int phaseFinishEnqueue(Context& c, const nbs_exec_args* args) {
    if (c.phase != 2) return fail(NBS_ERR_INVALID, "nbs_execute_finish called without nbs_execute_convolve");
    c.phase = 0;
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    cudaStream_t st = c.stream;
    const int N = c.N;
    const double* box = args->box;
    int status;
    if (args->forces) {
        if (args->forces_space == NBS_MEM_DEVICE) {
            if ((status = launchFinalize(c, args->forces, args->forces_format, args->padded_num_atoms,
                                         args->forces_accumulate, args->atom_index)) != NBS_OK) return status;
        }
        else {
            NBS_CUDA_CHECK(c.dForceOut.ensure(3*(size_t) N));
            if ((status = launchFinalize(c, c.dForceOut.d, NBS_FORCE_F64_XYZ, 0, 0, nullptr)) != NBS_OK) return status;
            if (args->forces_accumulate) {
                if (c.hForceCap < 3*(size_t) N) {
                    if (c.hForce) cudaFreeHost(c.hForce);
                    NBS_CUDA_CHECK(cudaMallocHost((void**) &c.hForce, sizeof(double)*3*N));
                    c.hForceCap = 3*(size_t) N;
                }
                NBS_CUDA_CHECK(cudaMemcpyAsync(c.hForce, c.dForceOut.d, sizeof(double)*3*N, cudaMemcpyDeviceToHost, st));
            }
            else
                NBS_CUDA_CHECK(cudaMemcpyAsync(args->forces, c.dForceOut.d, sizeof(double)*3*N, cudaMemcpyDeviceToHost, st));
        }
    }
    NBS_CUDA_CHECK(cudaMemcpyAsync(c.hEnergy, c.dEnergy.d, sizeof(double)*ENERGY_WORDS, cudaMemcpyDeviceToHost, st));
    NBS_CUDA_CHECK(cudaMemcpyAsync(c.hCounters, c.dCounters.d, sizeof(int)*16, cudaMemcpyDeviceToHost, st));
    timerMark(c, "d2h");
    return NBS_OK;
}

// Waits for the evaluation and does the host-side part: overflow check, constants, output tables.
static int phaseComplete(Context& c, const nbs_exec_args* args) {
    cudaStream_t st = c.stream;
    const int N = c.N;
    const double* box = args->box;
    NBS_CUDA_CHECK(cudaStreamSynchronize(st));
    NBS_CUDA_CHECK(cudaGetLastError());
    if (c.hCounters[1] != 0 || c.hEnergy[2*MAX_SLICES] != 0.0) {
        // neighbour-list capacity overflow (here or, after the all-reduce, on any rank): grow and redo
        if (c.capJ >= 65536) return fail(NBS_ERR_CAPACITY, "neighbour list capacity exceeded (system too dense for the tile list)");
        c.capJ *= 2;
        c.capX *= 2;
        return NBS_RETRY;
    }
    {
        float d2;
        std::memcpy(&d2, &c.hCounters[4], sizeof(float));
        const double disp = std::sqrt((double) d2), limit = 0.5*c.skin;
        if (c.reuseNow) {
            if (disp > limit) {
                c.listValid = false;
                if (c.phaseDirect) {              // the list may have missed pairs: redo with a fresh one
                    c.redoCount++;
                    c.forceRebuild = true;
                    return NBS_RETRY;
                }
            }
            c.dispStepMax = std::max(c.dispStepMax, disp - c.dispLast);
            c.dispLast = disp;
            if (disp + 1.5*c.dispStepMax > limit) c.forceRebuild = true;     // the next evaluation would probably exceed it
        }
        else {
            // a fresh sort (+ list): re-usable unless a block is so long that the pair kernel's window arithmetic
            // (corner - guard .. corner + box - guard) could not place a moved atom unambiguously
            c.buildCount++;
            // (slab sharding: the sorted-atom ranges of the cell columns that reach the own planes, k_slab_ranges)
            for (int k = 0; k < 4; k++) c.slabRange[k] = c.hCounters[8 + k];
            c.slabRangeValid = c.slabMode && (c.slabRange[1] > c.slabRange[0] || c.slabRange[3] > c.slabRange[2]);
            bool fits = true;
            for (int d = 0; d < 3; d++) {
                const double guard = (0.5*c.skin + 1.0e-3)/c.geom.box[d]*4294967296.0;
                if ((double) (unsigned) c.hCounters[5+d] + 2.0*guard + 65536.0 >= 4294967296.0) fits = false;
            }
            c.listValid = c.skin > 0 && c.periodic && fits;
            c.listHasDirect = c.phaseDirect;
            std::memcpy(c.listBox, box, sizeof(double)*9);
            c.listEpochAtBuild = c.listEpoch;
            c.listAllocEpoch = gAllocEpoch;
            c.listCapJ = c.capJ;
            c.forceRebuild = false;
            c.dispLast = 0;
            c.dispStepMax = 0;
        }
    }
    c.nBlocksLast = c.hCounters[0];
    c.haveLast = true;
    c.lastDirect = c.phaseDirect;
    std::memcpy(c.lastBox, box, sizeof(double)*9);
    if (args->forces && args->forces_space == NBS_MEM_HOST && args->forces_accumulate) {
        double* out = (double*) args->forces;
        for (size_t k = 0; k < 3*(size_t) N; k++) out[k] += c.hForce[k];
    }
    if (c.phaseEnergy) {
        double* E = args->slice_energies;
        for (int k = 0; k < 2*c.nSl; k++) E[k] = c.hEnergy[k];
        const double volume = box[0]*box[4]*box[8];
        if (c.phaseRecip) {
            // self energy and neutralising background, ReferenceSlicedLJCoulombIxn.cpp:203-222
            const double selfFactor = kOne4PiEps0*c.alpha/std::sqrt(kPi);
            const double factor = (-1/(4*c.alpha*c.alpha))/(2*kEpsilon0*volume);
            for (int i = 0; i < c.nS; i++) {
                E[2*(i*(i+3)/2)] -= selfFactor*c.subsetQ2[i];
                for (int j = i; j < c.nS; j++)
                    E[2*(j*(j+1)/2+i)] += (i == j ? 1 : 2)*c.subsetQ[i]*c.subsetQ[j]*factor;
            }
        }
        if (c.phaseRecip && c.ljpme()) {     // dispersion self term, ReferenceSlicedLJCoulombIxn.cpp:211-212
            const double a6 = std::pow(c.dispAlpha, 6.0);
            for (int i = 0; i < c.nS; i++) E[2*(i*(i+3)/2) + 1] += a6*c.subsetC6Self[i];
        }
        // dispersion correction: `periodic || ewald || pme` in the reference, i.e. NOT with LJPME, whose reciprocal
        // sum already carries the long-range dispersion (ReferenceNonbondedSlicingKernels.cpp:244-249, :317)
        if (c.phaseDirect && c.periodic && !c.ljpme())
            for (int s = 0; s < c.nSl; s++) E[2*s+1] += c.dispersion[s]/volume;
    }
    return NBS_OK;
}

// Returns NBS_OK, an error, or NBS_RETRY: the neighbour-list capacity was exceeded (on this or any
// other rank), it has been doubled, and the whole evaluation must be repeated.
static int phaseFinish(Context& c, const nbs_exec_args* args) {
    int status = phaseFinishEnqueue(c, args);
    if (status != NBS_OK) return status;
    return phaseComplete(c, args);
}

static unsigned long long mix64(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return h;
}

// Everything a captured evaluation depends on besides the contents of device memory.  Device memory itself is read at
// replay: positions, the caller's atom_index array (OpenMM re-orders atoms every few hundred steps -- the array's
// CONTENTS change, its address does not, and k_prep / k_reprep read it afresh in every replay), lambdas' effect
// travels through paramVersion (they are kernel arguments, so a change re-captures).
static unsigned long long graphSignature(const Context& c, const nbs_exec_args* a) {
    unsigned long long h = 0x243F6A8885A308D3ull;
    h = mix64(h, (unsigned long long) a->positions); h = mix64(h, (unsigned long long) a->forces);
    h = mix64(h, (unsigned long long) a->atom_index); h = mix64(h, (unsigned long long) a->stream);
    h = mix64(h, (unsigned long long) a->slice_energies != 0);
    h = mix64(h, ((unsigned long long) a->positions_format << 40) | ((unsigned long long) a->positions_space << 32) |
                 ((unsigned long long) a->forces_format << 24) | ((unsigned long long) a->forces_space << 16) |
                 ((unsigned long long) a->forces_accumulate << 8) | ((unsigned long long) (a->include_direct != 0) << 1) |
                 (unsigned long long) (a->include_reciprocal != 0));
    h = mix64(h, (unsigned long long) a->padded_num_atoms);
    for (int k = 0; k < 9; k++) { unsigned long long b; std::memcpy(&b, &a->box[k], 8); h = mix64(h, b); }
    h = mix64(h, c.paramVersion);
    h = mix64(h, gAllocEpoch);
    h = mix64(h, ((unsigned long long) c.capJ << 32) | (unsigned long long) c.capX);
    h = mix64(h, c.reuseNow ? 0x5bd1e995ull : 0ull);
    return h | 1ull;
}

// A captured evaluation never contains k_eterm (it is captured on the second identical evaluation, when the
// influence function is already up to date), so a replay is only valid while the tables on the device are still
// the ones of this box: an evaluation at another box in between (a rejected barostat trial) rewrites them.
static bool etermMatchesBox(const Context& c, const double* box) {
    if (!c.usesPmeGrid()) return true;
    const double want[6] = {box[0], box[4], box[8], box[3], box[6], box[7]};
    for (int k = 0; k < 6; k++) {
        if (c.etermBox[k] != want[k]) return false;
        if (c.ljpme() && c.etermBoxDisp[k] != want[k]) return false;
    }
    return true;
}

static bool hostPointerIsPinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

static int slabStepRun(Context& c, const nbs_exec_args* args, int step);

int nbs_execute(nbs_context* ctx, const nbs_exec_args* args) {
    if (!ctx || !args) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (c.slabMode && c.peersImported && c.peerBarrier) {
        // peer-memory sharding with in-kernel barriers: the five steps back to back, every rank the same
        for (int attempt = 0; attempt < 9; attempt++) {
            int status = NBS_OK;
            for (int step = 0; step < NBS_NUM_STEPS && status == NBS_OK; step++) status = slabStepRun(c, args, step);
            if (status != NBS_RETRY) return status;
        }
        return fail(NBS_ERR_CAPACITY, "neighbour list capacity exceeded");
    }
    if (c.nRanks != 1) return fail(NBS_ERR_INVALID, "a sharded context is driven through nbs_execute_begin/convolve/finish (or nbs_execute_step)");
    // A whole evaluation is ~20 short kernels: once the same evaluation (same buffers, box, parameters) has
    // run once, it is captured into a CUDA graph and replayed, which removes the per-launch gaps.
    bool graphable = !(c.flags & NBS_FLAG_NO_GRAPH) && !c.profiling && !c.paramsDirty && c.periodic && workStream(c, args) != nullptr;
    if (graphable && args->positions_space == NBS_MEM_HOST) graphable = hostPointerIsPinned(args->positions);
    if (graphable && args->forces && args->forces_space == NBS_MEM_HOST && !args->forces_accumulate)
        graphable = hostPointerIsPinned(args->forces);       // (accumulation goes through our own pinned staging buffer)
    for (int attempt = 0; attempt < 9; attempt++) {
        int status;
        planListReuse(c, args);
        const int slot = c.reuseNow ? 1 : 0;
        const unsigned long long key = graphable ? graphSignature(c, args) : 0;
        if (graphable && c.graphExecs[slot] && c.graphKeys[slot] == key && args->include_reciprocal && !etermMatchesBox(c, args->box)) {
            cudaGraphExecDestroy(c.graphExecs[slot]);     // stale influence function: plain path now, capture again later
            c.graphExecs[slot] = nullptr;
            c.graphKeys[slot] = 0;
        }
        if (graphable && c.graphExecs[slot] && c.graphKeys[slot] == key) {
            // replay
            status = validateExec(c, args);
            if (status != NBS_OK) return status;
            NBS_CUDA_CHECK(cudaSetDevice(c.device));
            c.stream = workStream(c, args);
            c.phaseEnergy = args->slice_energies != nullptr;
            c.phaseDirect = args->include_direct != 0;
            c.phaseRecip = args->include_reciprocal != 0 && c.ewaldDirect();
            c.evalCount++;
            NBS_CUDA_CHECK(cudaGraphLaunch(c.graphExecs[slot], c.stream));
            c.launches += c.graphLaunchCounts[slot];
            status = phaseComplete(c, args);
        }
        else if (graphable && c.warmKeys[slot] == key) {
            // second identical evaluation: capture it
            status = validateExec(c, args);
            if (status != NBS_OK) return status;
            NBS_CUDA_CHECK(cudaSetDevice(c.device));
            cudaStream_t st = workStream(c, args);
            if (c.graphExecs[slot]) { cudaGraphExecDestroy(c.graphExecs[slot]); c.graphExecs[slot] = nullptr; }
            const long long before = c.launches;
            NBS_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            status = phaseBegin(c, args, true);
            if (status == NBS_OK) status = phaseConvolve(c, args);
            if (status == NBS_OK) status = phaseFinishEnqueue(c, args);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (status != NBS_OK || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                c.warmKeys[slot] = 0;
                graphable = false;                 // fall back to plain launches for this call
                if (status != NBS_OK) return status;
                attempt--;
                continue;
            }
            e = cudaGraphInstantiate(&c.graphExecs[slot], graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { c.graphExecs[slot] = nullptr; cudaGetLastError(); c.warmKeys[slot] = 0; graphable = false; attempt--; continue; }
            c.graphKeys[slot] = key;
            c.graphLaunchCounts[slot] = c.launches - before;
            c.launches = before;
            NBS_CUDA_CHECK(cudaGraphLaunch(c.graphExecs[slot], st));
            c.launches += c.graphLaunchCounts[slot];
            status = phaseComplete(c, args);
        }
        else {
            status = phaseBegin(c, args, true);
            if (status == NBS_OK) status = phaseConvolve(c, args);
            if (status == NBS_OK) status = phaseFinish(c, args);
            if (status == NBS_OK && graphable) c.warmKeys[slot] = key;
        }
        if (status != NBS_RETRY) return status;
    }
    return fail(NBS_ERR_CAPACITY, "neighbour list capacity exceeded");
}

// ---- peer-memory sharding: the evaluation in five steps (include/nbslice_b200.h, nbs_execute_step) ------------------
static int slabStepRun(Context& c, const nbs_exec_args* args, int step) {
    if (!c.slabMode || !c.peersImported) return fail(NBS_ERR_INVALID, "nbs_execute_step needs nbs_set_slab_shard and nbs_import_peers first");
    if (step != c.slabStep) { c.slabStep = 0; c.phase = 0; return fail(NBS_ERR_INVALID, "nbs_execute_step: steps must run in order 0 .. NBS_NUM_STEPS-1"); }
    int status = NBS_OK;
    switch (step) {
        case 0:
            // sort, lists and direct space (own stream), spreading and z/y transforms of the own planes
            if ((status = phaseBegin(c, args)) != NBS_OK) break;
            if (c.phaseRecip && c.peerBarrier) status = launchPeerBarrier(c, false);      // every rank's planes are transformed
            break;
        case 1:
            NBS_CUDA_CHECK(cudaSetDevice(c.device));
            if (c.phaseRecip) {
                if ((status = launchPme(c, c.phaseEnergy, 2)) != NBS_OK) break;
                if (c.peerBarrier) status = launchPeerBarrier(c, false);                  // every rank's x pass has written its rows
            }
            break;
        case 2:
            NBS_CUDA_CHECK(cudaSetDevice(c.device));
            if (c.phaseRecip && (status = launchPme(c, c.phaseEnergy, 3)) != NBS_OK) break;
            if (c.directOverlapped) NBS_CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.evDirectDone, 0));
            status = launchPeerBarrier(c, true);                                          // forces are complete; energies published
            break;
        case 3:
            NBS_CUDA_CHECK(cudaSetDevice(c.device));
            if ((status = launchPeerReduce(c)) != NBS_OK) break;
            if (c.peerBarrier) status = launchPeerBarrier(c, false);                      // every rank's slice has been written back
            c.phase = 2;
            break;
        case 4: {
            c.slabStep = 0;
            NBS_CUDA_CHECK(cudaMemcpyAsync(c.hTimedOut, &c.dMailbox.d->timedOut, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
            status = phaseFinish(c, args);
            if (status >= 0 && *c.hTimedOut != 0)
                return fail(NBS_ERR_CUDA, "a peer-memory barrier timed out: some rank never reached the same step of the evaluation");
            return status;
        }
        default:
            return fail(NBS_ERR_INVALID, "nbs_execute_step: illegal step");
    }
    if (status != NBS_OK) {
        if (c.directOverlapped && step < 2) cudaStreamWaitEvent(c.stream, c.evDirectDone, 0);
        c.slabStep = 0; c.phase = 0;
        return status;
    }
    c.slabStep = step + 1;
    return NBS_OK;
}

int nbs_execute_step(nbs_context* ctx, const nbs_exec_args* args, int32_t step) {
    if (!ctx || !args) return fail(NBS_ERR_INVALID, "null argument");
    return slabStepRun(ctx->c, args, step);
}

int nbs_set_slab_shard(nbs_context* ctx, int32_t rank, int32_t num_ranks, int32_t block_period, int32_t block_offset,
                       int32_t block_width) {
    if (!ctx) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (num_ranks < 1 || num_ranks > NBS_MAX_RANKS || rank < 0 || rank >= num_ranks) return fail(NBS_ERR_INVALID, "illegal rank / num_ranks");
    if (block_period < 1 || block_width < 0 || block_offset < 0 || block_offset + block_width > block_period)
        return fail(NBS_ERR_INVALID, "illegal i-block share: need 0 <= offset, offset + width <= period");
    if (c.method != NBS_METHOD_PME) return fail(NBS_ERR_UNSUPPORTED, "peer-memory sharding is implemented for PME only");
    if (c.grid[0] < num_ranks || c.grid[1] < num_ranks) return fail(NBS_ERR_UNSUPPORTED, "more ranks than grid planes");
    if (c.flags & NBS_FLAG_LINE_FFT) return fail(NBS_ERR_UNSUPPORTED, "peer-memory sharding needs the plane-FFT path");
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    c.rank = rank; c.nRanks = num_ranks;
    c.paramVersion++;
    c.listEpoch++;
    c.blockPeriod = block_period; c.blockOffset = block_offset; c.blockWidth = block_width;
    c.ownLo = 0; c.ownHi = c.nS;
    c.slabMode = true; c.peersImported = false; c.slabStep = 0;
    c.xLo = (int) ((long long) rank*c.grid[0]/num_ranks); c.xHi = (int) ((long long) (rank + 1)*c.grid[0]/num_ranks);
    c.yLo = (int) ((long long) rank*c.grid[1]/num_ranks); c.yHi = (int) ((long long) (rank + 1)*c.grid[1]/num_ranks);
    c.haveLast = false;
    // the buffers the peers map must exist (and never move) before they are exported
    c.Npad = ((c.N + 31)/32)*32 + 32;
    NBS_CUDA_CHECK(c.dForce.ensure(6*(size_t) c.Npad));
    NBS_CUDA_CHECK(c.dMailbox.ensure(1));
    NBS_CUDA_CHECK(cudaMemset(c.dMailbox.d, 0, sizeof(PeerMailbox)));
    if (!c.hTimedOut) NBS_CUDA_CHECK(cudaMallocHost((void**) &c.hTimedOut, sizeof(unsigned long long)));
    *c.hTimedOut = 0;
    return NBS_OK;
}

int nbs_export_peer(nbs_context* ctx, nbs_peer_export* out) {
    if (!ctx || !out) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (out->struct_size != (int32_t) sizeof(nbs_peer_export)) return fail(NBS_ERR_INVALID, "nbs_peer_export.struct_size mismatch");
    if (!c.slabMode) return fail(NBS_ERR_INVALID, "nbs_export_peer needs nbs_set_slab_shard first");
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    out->rank = c.rank;
    out->process_id = (int64_t) getpid();
    out->device = c.device;
    out->reserved = 1;                       // 1 = the IPC handles below are valid
    out->spectra = c.dGridC.d; out->forces = c.dForce.d; out->mailbox = c.dMailbox.d;
    cudaIpcMemHandle_t h[3];
    void* ptrs[3] = {c.dGridC.d, c.dForce.d, c.dMailbox.d};
    for (int k = 0; k < 3; k++)
        if (cudaIpcGetMemHandle(&h[k], ptrs[k]) != cudaSuccess) { cudaGetLastError(); std::memset(&h[k], 0, sizeof(h[k])); out->reserved = 0; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(out->spectra_ipc, &h[0], 64); std::memcpy(out->forces_ipc, &h[1], 64); std::memcpy(out->mailbox_ipc, &h[2], 64);
    return NBS_OK;
}

int nbs_import_peers(nbs_context* ctx, int32_t count, const nbs_peer_export* all, int32_t in_kernel_barrier) {
    if (!ctx || !all) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (!c.slabMode || count != c.nRanks) return fail(NBS_ERR_INVALID, "nbs_import_peers: need one export per rank of nbs_set_slab_shard");
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    for (int r = 0; r < count; r++) {
        const nbs_peer_export& e = all[r];
        if (e.struct_size != (int32_t) sizeof(nbs_peer_export) || e.rank != r) return fail(NBS_ERR_INVALID, "nbs_import_peers: exports must be ordered by rank");
        if (r == c.rank) {
            c.peerSpectra[r] = c.dGridC.d; c.peerForce[r] = c.dForce.d; c.peerMailbox[r] = c.dMailbox.d;
        }
        else if (e.process_id == (int64_t) getpid()) {
            if (e.device != c.device) {          // same process, another device: plain peer access
                cudaError_t pe = cudaDeviceEnablePeerAccess(e.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(NBS_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
                cudaGetLastError();
            }
            c.peerSpectra[r] = e.spectra; c.peerForce[r] = (unsigned long long*) e.forces; c.peerMailbox[r] = (PeerMailbox*) e.mailbox;
        }
        else {
            if (!e.reserved) return fail(NBS_ERR_CUDA, "a peer could not export CUDA IPC handles for its buffers");
            const unsigned char* handles[3] = {e.spectra_ipc, e.forces_ipc, e.mailbox_ipc};
            void* mapped[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < 3; k++) {
                cudaIpcMemHandle_t h;
                std::memcpy(&h, handles[k], 64);
                cudaError_t ie = cudaIpcOpenMemHandle(&mapped[k], h, cudaIpcMemLazyEnablePeerAccess);
                if (ie != cudaSuccess) return fail(NBS_ERR_CUDA, std::string("cudaIpcOpenMemHandle (is peer-to-peer access available between the GPUs?): ") + cudaGetErrorString(ie));
                c.peerOpened[3*r + k] = mapped[k];
            }
            c.peerSpectra[r] = mapped[0]; c.peerForce[r] = (unsigned long long*) mapped[1]; c.peerMailbox[r] = (PeerMailbox*) mapped[2];
        }
    }
    c.peerBarrier = in_kernel_barrier != 0;
    c.peersImported = true;
    c.slabStep = 0;
    return NBS_OK;
}

int nbs_execute_begin(nbs_context* ctx, const nbs_exec_args* args) {
    if (!ctx || !args) return fail(NBS_ERR_INVALID, "null argument");
    return phaseBegin(ctx->c, args);
}

int nbs_execute_convolve(nbs_context* ctx, const nbs_exec_args* args) {
    if (!ctx || !args) return fail(NBS_ERR_INVALID, "null argument");
    return phaseConvolve(ctx->c, args);
}

int nbs_execute_finish(nbs_context* ctx, const nbs_exec_args* args) {
    if (!ctx || !args) return fail(NBS_ERR_INVALID, "null argument");
    return phaseFinish(ctx->c, args);
}

int nbs_set_shard(nbs_context* ctx, int32_t rank, int32_t num_ranks, int32_t block_period, int32_t block_offset,
                  int32_t block_width, int32_t subset_begin, int32_t subset_end) {
    if (!ctx) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (num_ranks < 1 || rank < 0 || rank >= num_ranks) return fail(NBS_ERR_INVALID, "illegal rank / num_ranks");
    if (block_period < 1 || block_width < 0 || block_offset < 0 || block_offset + block_width > block_period)
        return fail(NBS_ERR_INVALID, "illegal i-block share: need 0 <= offset, offset + width <= period");
    if (subset_begin < 0 || subset_end > c.nS || subset_begin > subset_end)
        return fail(NBS_ERR_INVALID, "illegal subset range");
    if (num_ranks > 1 && (c.method == NBS_METHOD_EWALD || c.method == NBS_METHOD_LJPME))
        return fail(NBS_ERR_UNSUPPORTED, "the plain Ewald sum and LJPME are not sharded across ranks (use PME)");
    c.rank = rank; c.nRanks = num_ranks;
    c.slabMode = false; c.peersImported = false;
    c.paramVersion++;
    c.listEpoch++;
    c.blockPeriod = block_period; c.blockOffset = block_offset; c.blockWidth = block_width;
    c.ownLo = subset_begin; c.ownHi = subset_end;
    c.haveLast = false;
    return NBS_OK;
}

int nbs_debug_set_list_capacity(nbs_context* ctx, int32_t j_capacity, int32_t x_capacity) {
    if (!ctx || j_capacity < 64 || x_capacity < 64 || j_capacity % 32 || x_capacity % 32)
        return fail(NBS_ERR_INVALID, "capacities must be multiples of 32, at least 64");
    ctx->c.paramVersion++;
    ctx->c.listEpoch++;
    ctx->c.capJ = j_capacity;
    ctx->c.capX = x_capacity;
    return NBS_OK;
}

int nbs_set_list_skin(nbs_context* ctx, double skin) {
    if (!ctx || !(skin >= 0) || skin > 1.0) return fail(NBS_ERR_INVALID, "the neighbour-list skin must lie in [0, 1] nm");
    Context& c = ctx->c;
    c.skin = c.periodic ? skin : 0.0;
    c.listEpoch++;
    c.paramVersion++;
    return NBS_OK;
}

int nbs_get_list_stats(const nbs_context* ctx, double out[8]) {
    if (!ctx || !out) return fail(NBS_ERR_INVALID, "null argument");
    const Context& c = ctx->c;
    out[0] = (double) c.evalCount; out[1] = (double) c.buildCount; out[2] = (double) c.redoCount;
    out[3] = c.dispLast; out[4] = c.skin; out[5] = c.listValid ? 1 : 0; out[6] = c.reuseNow ? 1 : 0; out[7] = c.dispStepMax;
    return NBS_OK;
}

int nbs_get_exchange_buffers(nbs_context* ctx, nbs_exchange_buffers* out) {
    if (!ctx || !out) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (out->struct_size != (int32_t) sizeof(nbs_exchange_buffers)) return fail(NBS_ERR_INVALID, "nbs_exchange_buffers.struct_size mismatch");
    const bool fp64 = (c.phaseEnergy && !(c.flags & NBS_FLAG_FP32_ENERGY)) || (c.flags & NBS_FLAG_DOUBLE);
    const size_t Gh = (size_t) c.grid[0]*c.grid[1]*(c.grid[2]/2 + 1);
    out->spectra = c.dGridC.d;
    out->spectrum_bytes_per_subset = (int64_t) (Gh*(fp64 ? sizeof(double2) : sizeof(float2)));
    out->spectrum_is_double = fp64 ? 1 : 0;
    out->forces = c.dForce.d;
    out->force_words = (c.pmeUnsorted ? 6 : 3)*(int64_t) c.Npad;
    out->energies = c.dEnergy.d;
    out->energy_words = ENERGY_WORDS;
    return NBS_OK;
}

// Host-only diagnostic: f(s) = erfc(alpha sqrt(s))/sqrt(s) exactly as the pair kernel's energy path evaluates it (k_pair.cu
// pairStep: row and position from the bits of the double, c0 in double + the single-precision remainder), from the table
// buildErfcTable uploads.  NaN where the kernel would take its analytic branch.  No device is touched.
int nbs_debug_erfc_table(double alpha, double cutoff, int32_t n, const double* s, double* f) {
    if (!s || !f || n < 0 || !(alpha > 0) || !(cutoff > 0)) return fail(NBS_ERR_INVALID, "illegal argument");
    std::vector<double> tab;
    const int rows = computeErfcTable(alpha, cutoff, tab);
    const float* coef = reinterpret_cast<const float*>(tab.data() + rows);
    constexpr int L = ERFC_TAB_PER_OCTAVE_LOG2;
    for (int i = 0; i < n; i++) {
        unsigned long long bits;
        std::memcpy(&bits, &s[i], 8);
        const int hi = (int) (bits >> 32);
        const unsigned lo = (unsigned) bits;
        const int idx = (hi >> (20 - L)) - ((1023 - 7) << L);
        if (idx < 0 || idx >= rows) { f[i] = std::nan(""); continue; }
        const unsigned frac = (((unsigned) hi & ((1u << (20 - L)) - 1u)) << (3 + L)) | (lo >> (29 - L));
        const unsigned fbits = 0x3f800000u | frac;
        float one;
        std::memcpy(&one, &fbits, 4);
        const float d = (one - 1.5f) + 5.9604645e-8f;
        const float* cf = coef + 4*(size_t) idx;
        const float rem = d*std::fmaf(d, std::fmaf(d, std::fmaf(d, cf[3], cf[2]), cf[1]), cf[0]);
        f[i] = tab[idx] + (double) rem;
    }
    return NBS_OK;
}

int nbs_get_pme_parameters(const nbs_context* ctx, double* alpha, int32_t* nx, int32_t* ny, int32_t* nz) {
    if (!ctx || !alpha || !nx || !ny || !nz) return fail(NBS_ERR_INVALID, "null argument");
    if (ctx->c.method != NBS_METHOD_PME && ctx->c.method != NBS_METHOD_LJPME)
        return fail(NBS_ERR_INVALID, "getPMEParametersInContext: This Context is not using PME or LJPME");
    *alpha = ctx->c.alpha; *nx = ctx->c.grid[0]; *ny = ctx->c.grid[1]; *nz = ctx->c.grid[2];
    return NBS_OK;
}

int nbs_get_ljpme_parameters(const nbs_context* ctx, double* alpha, int32_t* nx, int32_t* ny, int32_t* nz) {
    if (!ctx || !alpha || !nx || !ny || !nz) return fail(NBS_ERR_INVALID, "null argument");
    if (ctx->c.method != NBS_METHOD_LJPME)                       // ReferenceNonbondedSlicingKernels.cpp:330-337
        return fail(NBS_ERR_INVALID, "getPMEParametersInContext: This Context is not using LJPME");
    *alpha = ctx->c.dispAlpha; *nx = ctx->c.dispGrid[0]; *ny = ctx->c.dispGrid[1]; *nz = ctx->c.dispGrid[2];
    return NBS_OK;
}

int nbs_get_num_slices(const nbs_context* ctx, int32_t* num_slices) {
    if (!ctx || !num_slices) return fail(NBS_ERR_INVALID, "null argument");
    *num_slices = ctx->c.nSl;
    return NBS_OK;
}

int nbs_get_pair_set(nbs_context* ctx, int64_t capacity, int32_t* pairs, int64_t* count, uint64_t* hash) {
    if (!ctx || !count || !hash) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    if (!c.haveLast || !c.lastDirect) return fail(NBS_ERR_INVALID, "no direct-space evaluation to inspect");
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    c.stream = nullptr;
    const bool dump = pairs != nullptr && capacity > 0;
    if (dump) NBS_CUDA_CHECK(c.dPairDump.ensure((size_t) capacity));
    NBS_CUDA_CHECK(cudaMemset(c.dPairStats.d, 0, sizeof(unsigned long long)*4));
    const bool saved = c.profiling;
    c.profiling = false;
    int status = launchPairs(c, false, dump ? 2 : 1);
    c.profiling = saved;
    if (status != NBS_OK) return status;
    unsigned long long stats[4];
    NBS_CUDA_CHECK(cudaMemcpy(stats, c.dPairStats.d, sizeof(stats), cudaMemcpyDeviceToHost));
    *count = (int64_t) stats[0];
    *hash = stats[1];
    if (dump) {
        if ((int64_t) stats[0] > capacity) return fail(NBS_ERR_CAPACITY, "pair buffer too small");
        NBS_CUDA_CHECK(cudaMemcpy(pairs, c.dPairDump.d, sizeof(int2)*stats[0], cudaMemcpyDeviceToHost));
    }
    return NBS_OK;
}

int nbs_get_exclusion_set(nbs_context* ctx, int64_t capacity, int32_t* pairs, int64_t* count) {
    if (!ctx || !count) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    int status;
    if (c.paramsDirty && (status = applyParameters(c)) != NBS_OK) return status;
    // read back what the device actually uses
    std::vector<int> start(c.N+1), list(std::max<size_t>(1, 2*(size_t) c.nExc));
    NBS_CUDA_CHECK(cudaMemcpy(start.data(), c.dExclStart.d, sizeof(int)*(c.N+1), cudaMemcpyDeviceToHost));
    NBS_CUDA_CHECK(cudaMemcpy(list.data(), c.dExclList.d, sizeof(int)*list.size(), cudaMemcpyDeviceToHost));
    int64_t n = 0;
    for (int i = 0; i < c.N; i++)
        for (int k = start[i]; k < start[i+1]; k++)
            if (list[k] > i && (k == start[i] || list[k] != list[k-1])) {
                if (pairs && n < capacity) { pairs[2*n] = i; pairs[2*n+1] = list[k]; }
                n++;
            }
    *count = n;
    return NBS_OK;
}

int nbs_get_kernel_times(nbs_context* ctx, int32_t capacity, const char** names, float* milliseconds, int32_t* count) {
    if (!ctx || !count) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    int n = 0;
    for (size_t k = 1; k < c.timer.events.size(); k++) {
        if (n >= capacity) break;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c.timer.events[k-1], c.timer.events[k]) != cudaSuccess) { cudaGetLastError(); ms = -1; }
        names[n] = c.timer.names[k];
        milliseconds[n] = ms;
        n++;
    }
    *count = n;
    return NBS_OK;
}

int nbs_get_launch_count(const nbs_context* ctx, int64_t* launches) {
    if (!ctx || !launches) return fail(NBS_ERR_INVALID, "null argument");
    *launches = ctx->c.launches;
    return NBS_OK;
}

int nbs_get_nlist_stats(nbs_context* ctx, int64_t stats[8]) {
    if (!ctx || !stats) return fail(NBS_ERR_INVALID, "null argument");
    Context& c = ctx->c;
    for (int k = 0; k < 8; k++) stats[k] = 0;
    if (!c.haveLast || !c.lastDirect || c.blockWidth == 0) return NBS_OK;
    NBS_CUDA_CHECK(cudaSetDevice(c.device));
    int nb = 0;                                   // rank-local blocks that exist
    while (nb < c.maxLocalBlocks && localToGlobalBlock(nb, c.blockPeriod, c.blockOffset, c.blockWidth) < c.nBlocksLast) nb++;
    if (nb == 0) return NBS_OK;
    std::vector<int> jc(nb), xc(nb);
    NBS_CUDA_CHECK(cudaMemcpy(jc.data(), c.dJCount.d, sizeof(int)*nb, cudaMemcpyDeviceToHost));
    NBS_CUDA_CHECK(cudaMemcpy(xc.data(), c.dXCount.d, sizeof(int)*nb, cudaMemcpyDeviceToHost));
    // pair evaluations = 32 per step, one step per set bit of a group's cluster mask (k_pair.cu tileLoop)
    std::vector<unsigned> gj((size_t) nb*(c.capJ/32)), gx((size_t) nb*(c.capX/32));
    NBS_CUDA_CHECK(cudaMemcpy(gj.data(), c.dGmJ.d, sizeof(unsigned)*gj.size(), cudaMemcpyDeviceToHost));
    NBS_CUDA_CHECK(cudaMemcpy(gx.data(), c.dGmX.d, sizeof(unsigned)*gx.size(), cudaMemcpyDeviceToHost));
    long long entries = 0, tiles = 0, xentries = 0, steps = 0;
    for (int b = 0; b < nb; b++) {
        entries += jc[b] + xc[b];
        xentries += xc[b];
        tiles += (jc[b]+31)/32 + (xc[b]+31)/32;
        for (int t = 0; t < (jc[b]+31)/32; t++) steps += __builtin_popcount(gj[(size_t) b*(c.capJ/32) + t]);
        for (int t = 0; t < (xc[b]+31)/32; t++) steps += __builtin_popcount(gx[(size_t) b*(c.capX/32) + t]);
    }
    stats[0] = nb; stats[1] = entries; stats[2] = tiles; stats[3] = steps*32; stats[4] = xentries;
    stats[5] = c.capJ; stats[6] = c.geom.nCols; stats[7] = c.geom.nBins;
    return NBS_OK;
}

} // extern "C"

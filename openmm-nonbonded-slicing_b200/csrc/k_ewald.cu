// k_ewald.cu -- plain Ewald reciprocal sum with per-slice energies and lambda-scaled forces.
//
// What is computed is the reference's Ewald branch (platforms/reference/src/
// ReferenceSlicedLJCoulombIxn.cpp:256-358): for every reciprocal vector k = 2 pi (rx/Lx, ry/Ly, rz/Lz) of the
// half space {(0, 0, rz >= 1), (0, ry >= 1, any rz), (rx >= 1, any ry, any rz)}, |r_m| < numR_m, and every
// subset J the structure factor S_J(k) = sum_{n in J} q_n exp(i k.x_n); then
//   E_coul[slice(J,J)] += c_k |S_J|^2,  E_coul[slice(I,J)] += 2 c_k Re(S_I conj(S_J))          (:345-349)
//   F_n += 2 c_k k sum_J lambda_IJ (Re S_J Im(q_n e^{ik.x_n}) - Im S_J Re(q_n e^{ik.x_n}))        (:333-342)
// with c_k = ONE_4PI_EPS0 (4 pi / V) exp(-k^2 / (4 alpha^2)) / k^2.
//
// How it is computed is not the reference's (per-atom tables of exp(i m k_1 x) built by a power recurrence
// and a triple loop on one thread; the plugin's CUDA platform: one THREAD per reciprocal vector looping
// over all atoms, platforms/common/src/kernels/ewald.cc):
//   * the positions are the 32-bit fixed-point FRACTIONAL coordinates the cell sort already produced, so
//     the phase k.x / (2 pi) = rx fx + ry fy + rz fz is formed in 32-bit INTEGER arithmetic, where wrap-around
//     IS the reduction modulo one period: the argument of sincospi is exact to 2^-32 of a turn for any
//     |r|, and there is no table and no recurrence whose rounding grows with kmax;
//   * k_ewald_sums: one WARP per reciprocal vector (lanes stride over atoms, per-subset register
//     accumulators selected by predicates, shuffle reduction) -- thousands of warps even for a 648-atom box;
//   * k_ewald_mix: per vector, the slice energies (fixed-order block reduction: reproducible) and the
//     lambda-mixed factors T_I(k) = 2 c_k sum_J lambda_IJ S_J(k), so that
//   * k_ewald_forces (one warp per atom, lanes stride over vectors) needs one complex multiply per vector:
//     F_n = sum_k k (Re T_I Im(q_n e) - Im T_I Re(q_n e)); forces go to the 64-bit fixed-point accumulators.
// Everything is double precision: the sum is used for small systems, where it is cheap, and its slice energies
// are differences of large numbers.  Bound: FP64 / special-function issue rate (N x K sincospi).
#include "nbs_internal.h"
#include "nbs_device.cuh"

namespace nbs {

struct EwaldArgs {
    int N, Npad, nS, nK;
    double recipBox[3];                  // 2 pi / L
    double factorEwald;                  // -1 / (4 alpha^2)
    double recipCoeff;                   // ONE_4PI_EPS0 4 pi / V
    const uint4* fix;                    // particle order: fixed-point fractional xyz
    const double* charge;                // particle order
    const int* subsetOf;
    const int4* kvec;                    // [nK] (rx, ry, rz, 0)
    double2* sums;                       // [nK][MAX_SUBSETS] structure factors (cos sum, sin sum)
    double2* mixed;                      // [nK][MAX_SUBSETS] T_I(k)
    unsigned long long* force;           // [3][Npad] particle order
    double* energy;                      // [nSl][2]
    int wantEnergy;
    double lamC[MAX_SLICES];
};

__device__ __forceinline__ void phaseFactor(const uint4 f, const int4 k, double& c, double& s) {
    // turns = rx fx + ry fy + rz fz (mod 1), exact in 32-bit wrap-around arithmetic
    const unsigned turns = (unsigned) k.x*f.x + (unsigned) k.y*f.y + (unsigned) k.z*f.z;
    sincospi((double) (int) turns*(1.0/2147483648.0), &s, &c);
}

__global__ void __launch_bounds__(256) k_ewald_sums(const EwaldArgs a) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= a.nK) return;
    const int4 kv = a.kvec[k];
    double cs[MAX_SUBSETS], ss[MAX_SUBSETS];
#pragma unroll
    for (int j = 0; j < MAX_SUBSETS; j++) { cs[j] = 0.0; ss[j] = 0.0; }
    for (int n = lane; n < a.N; n += 32) {
        double c, s;
        phaseFactor(a.fix[n], kv, c, s);
        const double q = a.charge[n];
        const int sub = a.subsetOf[n];
#pragma unroll
        for (int j = 0; j < MAX_SUBSETS; j++)
            if (j == sub) { cs[j] += q*c; ss[j] += q*s; }
    }
#pragma unroll
    for (int j = 0; j < MAX_SUBSETS; j++) {
        if (j < a.nS) {
            const double c = warpSum(cs[j]), s = warpSum(ss[j]);
            if (lane == 0) a.sums[(size_t) k*MAX_SUBSETS + j] = make_double2(c, s);
        }
    }
}

// One CTA.  Thread t owns vectors t, t + 256, ...; slice energies are reduced in a fixed order.
__global__ void __launch_bounds__(256) k_ewald_mix(const EwaldArgs a) {
    __shared__ double red[8];
    double e[MAX_SLICES];
#pragma unroll
    for (int s = 0; s < MAX_SLICES; s++) e[s] = 0.0;
    for (int k = threadIdx.x; k < a.nK; k += blockDim.x) {
        const int4 kv = a.kvec[k];
        const double kx = kv.x*a.recipBox[0], ky = kv.y*a.recipBox[1], kz = kv.z*a.recipBox[2];
        const double k2 = kx*kx + ky*ky + kz*kz;
        const double ck = a.recipCoeff*exp(k2*a.factorEwald)/k2;
        double2 S[MAX_SUBSETS];
#pragma unroll
        for (int j = 0; j < MAX_SUBSETS; j++) S[j] = j < a.nS ? a.sums[(size_t) k*MAX_SUBSETS + j] : make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < MAX_SUBSETS; i++) {
            if (i >= a.nS) break;
            double tr = 0.0, ti = 0.0;
#pragma unroll
            for (int j = 0; j < MAX_SUBSETS; j++) {
                if (j >= a.nS) break;
                const double lam = a.lamC[triSlice(i, j)];
                tr += lam*S[j].x; ti += lam*S[j].y;
                if (j < i) e[i*(i+1)/2 + j] += 2*ck*(S[i].x*S[j].x + S[i].y*S[j].y);
            }
            e[i*(i+3)/2] += ck*(S[i].x*S[i].x + S[i].y*S[i].y);
            a.mixed[(size_t) k*MAX_SUBSETS + i] = make_double2(2*ck*tr, 2*ck*ti);
        }
    }
    if (!a.wantEnergy) return;
    const int nSl = a.nS*(a.nS+1)/2;
    for (int s = 0; s < nSl; s++) {
        const double w = warpSum(e[s]);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int q = 0; q < (int) (blockDim.x >> 5); q++) t += red[q];
            atomicAdd(a.energy + 2*s, t);
        }
    }
}

__global__ void __launch_bounds__(256) k_ewald_forces(const EwaldArgs a) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= a.N) return;
    const uint4 f = a.fix[n];
    const double q = a.charge[n];
    const int sub = a.subsetOf[n];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int k = lane; k < a.nK; k += 32) {
        const int4 kv = a.kvec[k];
        double c, s;
        phaseFactor(f, kv, c, s);
        const double2 T = a.mixed[(size_t) k*MAX_SUBSETS + sub];
        const double w = q*(T.x*s - T.y*c);
        fx += w*kv.x; fy += w*kv.y; fz += w*kv.z;
    }
    fx = warpSum(fx); fy = warpSum(fy); fz = warpSum(fz);
    if (lane == 0) {
        atomicAdd(a.force + n, toFixed(fx*a.recipBox[0]));
        atomicAdd(a.force + a.Npad + n, toFixed(fy*a.recipBox[1]));
        atomicAdd(a.force + 2*(size_t) a.Npad + n, toFixed(fz*a.recipBox[2]));
    }
}

// The half space of reciprocal vectors, in the reference's loop order (:288-353).
int uploadEwaldVectors(Context& c) {
    std::vector<int4> kv;
    const int* numR = c.ewaldKmax;
    for (int rx = 0; rx < numR[0]; rx++)
        for (int ry = (rx == 0 ? 0 : 1 - numR[1]); ry < numR[1]; ry++)
            for (int rz = (rx == 0 && ry == 0 ? 1 : 1 - numR[2]); rz < numR[2]; rz++)
                kv.push_back(make_int4(rx, ry, rz, 0));
    c.ewaldNK = (int) kv.size();
    if (c.ewaldNK == 0) return NBS_OK;
    NBS_CUDA_CHECK(c.dEwaldK.ensure(kv.size()));
    NBS_CUDA_CHECK(c.dEwaldSums.ensure(kv.size()*MAX_SUBSETS));
    NBS_CUDA_CHECK(c.dEwaldMixed.ensure(kv.size()*MAX_SUBSETS));
    NBS_CUDA_CHECK(cudaMemcpy(c.dEwaldK.d, kv.data(), sizeof(int4)*kv.size(), cudaMemcpyHostToDevice));
    return NBS_OK;
}

// Reciprocal part of an Ewald evaluation, on c.stream; needs k_prep's fixed-point coordinates only.
int launchEwald(Context& c, bool wantEnergy) {
    if (c.ewaldNK == 0) return NBS_OK;
    const CellGeom& g = c.geom;
    EwaldArgs a;
    a.N = c.N; a.Npad = c.Npad; a.nS = c.nS; a.nK = c.ewaldNK;
    for (int k = 0; k < 3; k++) a.recipBox[k] = 2*kPi/g.box[k];
    a.factorEwald = -1/(4*c.alpha*c.alpha);
    a.recipCoeff = kOne4PiEps0*4*kPi/(g.box[0]*g.box[1]*g.box[2]);
    a.fix = c.dFix.d; a.charge = c.dCharge.d; a.subsetOf = c.dSubset.d;
    a.kvec = c.dEwaldK.d; a.sums = c.dEwaldSums.d; a.mixed = c.dEwaldMixed.d;
    a.force = c.dForce.d + 3*(size_t) c.Npad;
    a.energy = c.dEnergy.d;
    a.wantEnergy = wantEnergy ? 1 : 0;
    for (int s = 0; s < MAX_SLICES; s++) a.lamC[s] = s < c.nSl ? c.lambdas[2*s] : 1.0;
    cudaStream_t st = c.stream;
    k_ewald_sums<<<(a.nK + 7)/8, 256, 0, st>>>(a);
    k_ewald_mix<<<1, 256, 0, st>>>(a);
    k_ewald_forces<<<(a.N + 7)/8, 256, 0, st>>>(a);
    c.launches += 3;
    timerMark(c, "ewald");
    return NBS_OK;
}

} // namespace nbs

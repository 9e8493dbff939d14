// k_pme.cu -- sliced smooth-PME reciprocal space: per-subset B-spline spreading, a hand-written
// 3D real FFT, the sliced reciprocal convolution (cross-subset structure-factor products per slice),
// and the force gather.
//
// Arithmetic follows platforms/reference/src/ReferencePME.cpp (order 5, :754-811): index/fraction
// :196-256, splines :264-317, spread :320-396, convolution + slice energies :400-496, gather :598-702.
// Differences in STRUCTURE (not in results), all B200-motivated:
//   * charges carry sqrt(ONE_4PI_EPS0) so the influence function needs no unit factor;
//   * real-to-complex transforms on half spectra instead of the reference's complex-to-complex;
//   * the x-direction forward FFT, the convolution, the slice energies, the lambda mixing of the
//     subset potentials (G_I = eterm * sum_J lambda_IJ S_J, legal because everything is linear) and the
//     x-direction inverse FFT are ONE kernel that keeps the lines of all subsets in shared memory; the
//     gather then reads 125 points from one grid per atom instead of 125 * nSubsets (pme.cc:360-371);
//   * FFT lines live in shared memory, one warp per line, Stockham passes staged through registers
//     (radices 4, 2, 3, 5, 7, 11, 13 -- the factor set the reference's VkFFT path accepts,
//     platforms/common/include/FFT3DFactory.h:45-47).
// Bound: HBM/L2 bandwidth and launch latency (grids of every BASELINE config fit in the 126 MB L2).
#include "nbs_internal.h"
#include "nbs_device.cuh"

namespace nbs {

constexpr int FFT_MAX_N = 512;

struct FftPlan {           // factorisation of one dimension
    int n, nf;
    int f[12];
};

static bool makePlan(int n, FftPlan& p) {
    p.n = n; p.nf = 0;
    int m = n;
    while (m % 4 == 0) { p.f[p.nf++] = 4; m /= 4; }
    const int radices[] = {2, 3, 5, 7, 11, 13};
    for (int r : radices)
        while (m % r == 0) { p.f[p.nf++] = r; m /= r; }
    return m == 1 && n <= FFT_MAX_N;
}

// ---------------------------------------------------------------------------------------------
// One Stockham pass of radix R over a line of length n in shared memory, executed by one warp.
// Inputs are staged through registers, so the pass is in place (read all, sync, write all).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x);
}

template <int R>
__device__ __forceinline__ void fftPass(float2* line, int n, int Ns, const float2* tw, int lane) {
    constexpr int MAXB = (FFT_MAX_N/R + 31)/32;
    const int nb = n/R;
    const int tstep = n/(Ns*R);
    const int rstep = n/R;
    float2 v[MAXB][R];
#pragma unroll
    for (int q = 0; q < MAXB; q++) {
        const int j = lane + 32*q;
        if (j < nb) {
#pragma unroll
            for (int t = 0; t < R; t++) v[q][t] = line[j + t*nb];
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < MAXB; q++) {
        const int j = lane + 32*q;
        if (j < nb) {
            const int k = j % Ns;
#pragma unroll
            for (int t = 1; t < R; t++) v[q][t] = cmul(v[q][t], tw[t*k*tstep]);
            const int j0 = (j/Ns)*Ns*R + k;
            if (R == 2) {
                line[j0] = make_float2(v[q][0].x + v[q][1].x, v[q][0].y + v[q][1].y);
                line[j0 + Ns] = make_float2(v[q][0].x - v[q][1].x, v[q][0].y - v[q][1].y);
            }
            else if (R == 4) {
                const float2 a0 = make_float2(v[q][0].x + v[q][2].x, v[q][0].y + v[q][2].y);
                const float2 a1 = make_float2(v[q][0].x - v[q][2].x, v[q][0].y - v[q][2].y);
                const float2 a2 = make_float2(v[q][1].x + v[q][3].x, v[q][1].y + v[q][3].y);
                const float2 a3 = make_float2(v[q][1].x - v[q][3].x, v[q][1].y - v[q][3].y);
                // forward transform: multiply a3 by -i
                line[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
                line[j0 + Ns] = make_float2(a1.x + a3.y, a1.y - a3.x);
                line[j0 + 2*Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
                line[j0 + 3*Ns] = make_float2(a1.x - a3.y, a1.y + a3.x);
            }
            else {
#pragma unroll
                for (int o = 0; o < R; o++) {
                    float2 acc = v[q][0];
#pragma unroll
                    for (int t = 1; t < R; t++) {
                        const float2 w = tw[((o*t) % R)*rstep];
                        acc.x += v[q][t].x*w.x - v[q][t].y*w.y;
                        acc.y += v[q][t].x*w.y + v[q][t].y*w.x;
                    }
                    line[j0 + o*Ns] = acc;
                }
            }
        }
    }
    __syncwarp();
}

// Forward (e^{-i...}) unnormalised FFT of one shared-memory line by one warp.
__device__ __forceinline__ void warpFft(float2* line, const FftPlan& plan, const float2* tw, int lane) {
    int Ns = 1;
    for (int f = 0; f < plan.nf; f++) {
        const int R = plan.f[f];
        switch (R) {
            case 2: fftPass<2>(line, plan.n, Ns, tw, lane); break;
            case 3: fftPass<3>(line, plan.n, Ns, tw, lane); break;
            case 4: fftPass<4>(line, plan.n, Ns, tw, lane); break;
            case 5: fftPass<5>(line, plan.n, Ns, tw, lane); break;
            case 7: fftPass<7>(line, plan.n, Ns, tw, lane); break;
            case 11: fftPass<11>(line, plan.n, Ns, tw, lane); break;
            default: fftPass<13>(line, plan.n, Ns, tw, lane); break;
        }
        Ns *= R;
    }
}

// ---------------------------------------------------------------------------------------------
// Spreading: one warp per (sorted) atom; lanes 0..24 own an (ix, iy) offset and walk the 5 z points.
// Reference: pme_grid_spread_charge, ReferencePME.cpp:320-396 (forward-only spreading, :375-394).
// ---------------------------------------------------------------------------------------------
struct PmeArgs {
    int N, Npad, nS, nx, ny, nz, nzh;
    const uint4* posq; const float4* par;
    float* grid; float2* gridC; const float* eterm;
    unsigned long long* force; double* energy;
    float fscale[3];             // n_d / L_d
    LambdaTable lam;
};

__device__ __forceinline__ void gridCoord(unsigned fixed, int n, int& index, float& frac) {
    const unsigned long long t = (unsigned long long) fixed*(unsigned) n;     // frac * n in 32.32 fixed point
    index = (int) (t >> 32);
    frac = (float) (unsigned) (t & 0xffffffffull)*(1.0f/4294967296.0f);
}

__global__ void __launch_bounds__(256) k_spread(const PmeArgs a) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    if (j >= a.N) return;
    const uint4 p = a.posq[j];
    const int subset = __float_as_int(a.par[j].z);
    const float q = __uint_as_float(p.w);
    // lanes 0-4 x, 5-9 y, 10-14 z: each lane keeps the weight lane%5 of "its" dimension
    const int dim = min(lane/5, 2), kk = lane % 5;
    int index; float frac;
    gridCoord(dim == 0 ? p.x : (dim == 1 ? p.y : p.z), dim == 0 ? a.nx : (dim == 1 ? a.ny : a.nz), index, frac);
    float th[5], dth[5];
    bspline5(frac, th, dth);
    float mine = th[0];
#pragma unroll
    for (int k = 1; k < 5; k++) mine = kk == k ? th[k] : mine;
    const int ix0 = __shfl_sync(FULL_MASK, index, 0), iy0 = __shfl_sync(FULL_MASK, index, 5), iz0 = __shfl_sync(FULL_MASK, index, 10);
    const int ox = lane/5, oy = lane % 5;
    const float tx = __shfl_sync(FULL_MASK, mine, min(ox, 4)), ty = __shfl_sync(FULL_MASK, mine, 5 + oy);
    float tz[5];
#pragma unroll
    for (int k = 0; k < 5; k++) tz[k] = __shfl_sync(FULL_MASK, mine, 10 + k);
    if (lane >= 25 || q == 0.f) return;
    int x = ix0 + ox; x -= x >= a.nx ? a.nx : 0;
    int y = iy0 + oy; y -= y >= a.ny ? a.ny : 0;
    float* row = a.grid + (((size_t) subset*a.nx + x)*a.ny + y)*a.nz;
    const float w = q*tx*ty;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int z = iz0 + k; z -= z >= a.nz ? a.nz : 0;
        atomicAdd(row + z, w*tz[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// z transform, real -> half complex: two real lines (y, y+1) ride one complex FFT.
// ---------------------------------------------------------------------------------------------
struct FftArgs {
    int nS, nx, ny, nz, nzh;
    FftPlan plan;
    const float2* tw;            // twiddles of this dimension: exp(-2 pi i k / n)
    float* grid; float2* gridC; const float* eterm;
    double* energy;
    int wantEnergy;
    LambdaTable lam;
};

__global__ void __launch_bounds__(256) k_fft_z_fwd(const FftArgs a) {
    extern __shared__ float2 sm[];
    const int n = a.nz, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* tw = sm;
    float2* line = sm + n + (size_t) warp*n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = a.tw[k];
    __syncthreads();
    const int halfY = (a.ny + 1) >> 1;
    const int pair = blockIdx.x*8 + warp;
    if (pair >= a.nS*a.nx*halfY) return;
    const int sx = pair/halfY, y0 = 2*(pair - sx*halfY), y1 = y0 + 1;
    const float* r0 = a.grid + ((size_t) sx*a.ny + y0)*n;
    const float* r1 = r0 + n;
    for (int z = lane; z < n; z += 32) line[z] = make_float2(r0[z], y1 < a.ny ? r1[z] : 0.f);
    __syncwarp();
    warpFft(line, a.plan, tw, lane);
    float2* o0 = a.gridC + ((size_t) sx*a.ny + y0)*a.nzh;
    float2* o1 = o0 + a.nzh;
    for (int k = lane; k < a.nzh; k += 32) {
        const float2 zk = line[k], zn = line[k == 0 ? 0 : n - k];
        o0[k] = make_float2(0.5f*(zk.x + zn.x), 0.5f*(zk.y - zn.y));
        if (y1 < a.ny) o1[k] = make_float2(0.5f*(zk.y + zn.y), -0.5f*(zk.x - zn.x));
    }
}

// z transform, half complex -> real (inverse, unnormalised), two lines per complex FFT.
__global__ void __launch_bounds__(256) k_fft_z_inv(const FftArgs a) {
    extern __shared__ float2 sm[];
    const int n = a.nz, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* tw = sm;
    float2* line = sm + n + (size_t) warp*n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = a.tw[k];
    __syncthreads();
    const int halfY = (a.ny + 1) >> 1;
    const int pair = blockIdx.x*8 + warp;
    if (pair >= a.nS*a.nx*halfY) return;
    const int sx = pair/halfY, y0 = 2*(pair - sx*halfY), y1 = y0 + 1;
    const float2* i0 = a.gridC + ((size_t) sx*a.ny + y0)*a.nzh;
    const float2* i1 = i0 + a.nzh;
    for (int k = lane; k < a.nzh; k += 32) {
        const float2 A = i0[k], B = y1 < a.ny ? i1[k] : make_float2(0.f, 0.f);
        line[k] = make_float2(A.x - B.y, -(A.y + B.x));                 // conj(A + iB)
        if (k > 0 && 2*k < n) line[n - k] = make_float2(A.x + B.y, A.y - B.x);   // conj(conj(A) + i conj(B))
    }
    __syncwarp();
    warpFft(line, a.plan, tw, lane);
    float* r0 = a.grid + ((size_t) sx*a.ny + y0)*n;
    float* r1 = r0 + n;
    for (int z = lane; z < n; z += 32) {
        const float2 w = line[z];
        r0[z] = w.x;
        if (y1 < a.ny) r1[z] = -w.y;
    }
}

// y transform on the half spectrum, in place.  CTA = one (subset, x) plane x 16 consecutive kz.
template <bool INVERSE>
__global__ void __launch_bounds__(512) k_fft_y(const FftArgs a) {
    extern __shared__ float2 sm[];
    const int n = a.ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = n + 1;
    float2* tw = sm;
    float2* lines = sm + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = a.tw[k];
    const int chunks = (a.nzh + 15) >> 4;
    const int sx = blockIdx.x/chunks, k0 = (blockIdx.x - sx*chunks)*16;
    float2* base = a.gridC + (size_t) sx*n*a.nzh;
    for (int idx = threadIdx.x; idx < n*16; idx += blockDim.x) {
        const int l = idx & 15, y = idx >> 4;
        if (k0 + l < a.nzh) {
            float2 v = base[(size_t) y*a.nzh + k0 + l];
            if (INVERSE) v.y = -v.y;
            lines[l*stride + y] = v;
        }
    }
    __syncthreads();
    if (k0 + warp < a.nzh) warpFft(lines + warp*stride, a.plan, tw, lane);
    __syncthreads();
    for (int idx = threadIdx.x; idx < n*16; idx += blockDim.x) {
        const int l = idx & 15, y = idx >> 4;
        if (k0 + l < a.nzh) {
            float2 v = lines[l*stride + y];
            if (INVERSE) v.y = -v.y;
            base[(size_t) y*a.nzh + k0 + l] = v;
        }
    }
}

// x transform + sliced convolution + x inverse.  CTA = one y x 8 consecutive kz, ALL subsets.
// Convolution and energies: pme_reciprocal_convolution, ReferencePME.cpp:400-496 -- eterm per k,
// E[slice(I,I)] += 1/2 eterm |S_I|^2, E[slice(I,J)] += eterm Re(S_I conj S_J) over the FULL grid (the
// half spectrum counts twice except on the kz = 0 and kz = nz/2 planes).  The reference then scales
// every subset grid by eterm and lets the gather mix subsets with lambda (:681-687); here the mix
// happens in k space.
template <int NS>
__global__ void __launch_bounds__(256) k_fft_x_conv(const FftArgs a) {
    extern __shared__ float2 sm[];
    __shared__ double shE[MAX_SLICES];
    const int n = a.nx, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = n + 1;
    float2* tw = sm;
    float2* lines = sm + n;                       // [s][l][x]
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = a.tw[k];
    if (threadIdx.x < MAX_SLICES) shE[threadIdx.x] = 0.0;
    const int chunks = (a.nzh + 7) >> 3;
    const int y = blockIdx.x/chunks, k0 = (blockIdx.x - y*chunks)*8;
    const int nS = a.nS;
    for (int idx = threadIdx.x; idx < nS*n*8; idx += blockDim.x) {
        const int l = idx & 7, x = (idx >> 3) % n, s = idx/(8*n);
        if (k0 + l < a.nzh)
            lines[(s*8 + l)*stride + x] = a.gridC[(((size_t) s*n + x)*a.ny + y)*a.nzh + k0 + l];
    }
    __syncthreads();
    for (int L = warp; L < nS*8; L += 8)
        if (k0 + (L & 7) < a.nzh) warpFft(lines + L*stride, a.plan, tw, lane);
    __syncthreads();
    double e[NS*(NS+1)/2];
#pragma unroll
    for (int s = 0; s < NS*(NS+1)/2; s++) e[s] = 0.0;
    for (int idx = threadIdx.x; idx < n*8; idx += blockDim.x) {
        const int l = idx & 7, x = idx >> 3, k = k0 + l;
        if (k >= a.nzh) continue;
        const float et = a.eterm[((size_t) x*a.ny + y)*a.nzh + k];
        float2 S[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) S[s] = s < nS ? lines[(s*8 + l)*stride + x] : make_float2(0.f, 0.f);
        if (a.wantEnergy) {
            const float w = (k == 0 || 2*k == a.nz) ? 1.f : 2.f;
#pragma unroll
            for (int sb = 0; sb < NS; sb++)
#pragma unroll
                for (int sa = 0; sa <= sb; sa++) {
                    const float prod = S[sa].x*S[sb].x + S[sa].y*S[sb].y;
                    e[sb*(sb+1)/2 + sa] += (double) ((sa == sb ? 0.5f : 1.f)*w*et*prod);
                }
        }
#pragma unroll
        for (int si = 0; si < NS; si++) {
            if (si >= nS) break;
            float gx = 0.f, gy = 0.f;
#pragma unroll
            for (int sj = 0; sj < NS; sj++) {
                const float lam = a.lam.c[triSlice(si, sj)];
                gx = fmaf(lam, S[sj].x, gx);
                gy = fmaf(lam, S[sj].y, gy);
            }
            lines[(si*8 + l)*stride + x] = make_float2(et*gx, -et*gy);      // conjugated for the inverse pass
        }
    }
    __syncthreads();
    for (int L = warp; L < nS*8; L += 8)
        if (k0 + (L & 7) < a.nzh) warpFft(lines + L*stride, a.plan, tw, lane);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nS*n*8; idx += blockDim.x) {
        const int l = idx & 7, x = (idx >> 3) % n, s = idx/(8*n);
        if (k0 + l < a.nzh) {
            float2 v = lines[(s*8 + l)*stride + x];
            v.y = -v.y;
            a.gridC[(((size_t) s*n + x)*a.ny + y)*a.nzh + k0 + l] = v;
        }
    }
    if (a.wantEnergy) {
#pragma unroll
        for (int s = 0; s < NS*(NS+1)/2; s++) {
            const double v = warpSum(e[s]);
            if (lane == 0 && v != 0.0) atomicAdd(&shE[s], v);
        }
        __syncthreads();
        if (threadIdx.x < NS*(NS+1)/2 && shE[threadIdx.x] != 0.0)
            atomicAdd(a.energy + 2*threadIdx.x, shE[threadIdx.x]);       // Coulomb term of the slice
    }
}

// ---------------------------------------------------------------------------------------------
// Gather: one warp per atom, 125 points of the atom's own (lambda-mixed) potential grid.
// Reference: pme_grid_interpolate_force, ReferencePME.cpp:598-702.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather(const PmeArgs a) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    if (j >= a.N) return;
    const uint4 p = a.posq[j];
    const float q = __uint_as_float(p.w);
    if (q == 0.f) return;
    const int subset = __float_as_int(a.par[j].z);
    const int dim = min(lane/5, 2), kk = lane % 5;
    int index; float frac;
    gridCoord(dim == 0 ? p.x : (dim == 1 ? p.y : p.z), dim == 0 ? a.nx : (dim == 1 ? a.ny : a.nz), index, frac);
    float th[5], dth[5];
    bspline5(frac, th, dth);
    float mine = th[0], dmine = dth[0];
#pragma unroll
    for (int k = 1; k < 5; k++) { mine = kk == k ? th[k] : mine; dmine = kk == k ? dth[k] : dmine; }
    const int ix0 = __shfl_sync(FULL_MASK, index, 0), iy0 = __shfl_sync(FULL_MASK, index, 5), iz0 = __shfl_sync(FULL_MASK, index, 10);
    const int ox = min(lane/5, 4), oy = lane % 5;
    const float tx = __shfl_sync(FULL_MASK, mine, ox), ty = __shfl_sync(FULL_MASK, mine, 5 + oy);
    const float dtx = __shfl_sync(FULL_MASK, dmine, ox), dty = __shfl_sync(FULL_MASK, dmine, 5 + oy);
    float tz[5], dtz[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { tz[k] = __shfl_sync(FULL_MASK, mine, 10 + k); dtz[k] = __shfl_sync(FULL_MASK, dmine, 10 + k); }
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (lane < 25) {
        int x = ix0 + ox; x -= x >= a.nx ? a.nx : 0;
        int y = iy0 + oy; y -= y >= a.ny ? a.ny : 0;
        const float* row = a.grid + (((size_t) subset*a.nx + x)*a.ny + y)*a.nz;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            int z = iz0 + k; z -= z >= a.nz ? a.nz : 0;
            const float g = row[z];
            s0 = fmaf(tz[k], g, s0);
            s1 = fmaf(dtz[k], g, s1);
        }
        fx = dtx*ty*s0;
        fy = tx*dty*s0;
        fz = tx*ty*s1;
    }
    fx = warpSum(fx); fy = warpSum(fy); fz = warpSum(fz);
    if (lane == 0) {
        atomicAdd(a.force + j, toFixed(-q*fx*a.fscale[0]));
        atomicAdd(a.force + a.Npad + j, toFixed(-q*fy*a.fscale[1]));
        atomicAdd(a.force + 2*(size_t) a.Npad + j, toFixed(-q*fz*a.fscale[2]));
    }
}

// ---------------------------------------------------------------------------------------------
// Influence function eterm(k) = exp(-pi^2 m^2 / alpha^2) / (pi V m^2 Bx By Bz), ReferencePME.cpp:426-471
// (without ONE_4PI_EPS0, which the charges carry).  Recomputed only when the box changes.
// ---------------------------------------------------------------------------------------------
__global__ void k_eterm(int nx, int ny, int nz, int nzh, double3 invBox, double volume, double alpha,
                        const double* __restrict__ moduli, float* __restrict__ eterm) {
    const size_t idx = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= (size_t) nx*ny*nzh) return;
    const int kz = (int) (idx % nzh), ky = (int) ((idx/nzh) % ny), kx = (int) (idx/((size_t) nzh*ny));
    if (kx == 0 && ky == 0 && kz == 0) { eterm[idx] = 0.f; return; }
    const double mx = (kx < (nx+1)/2 ? kx : kx - nx)*invBox.x;
    const double my = (ky < (ny+1)/2 ? ky : ky - ny)*invBox.y;
    const double mz = (kz < (nz+1)/2 ? kz : kz - nz)*invBox.z;
    const double m2 = mx*mx + my*my + mz*mz;
    const double denom = m2*kPi*volume*moduli[kx]*moduli[nx + ky]*moduli[nx + ny + kz];
    eterm[idx] = (float) (exp(-kPi*kPi*m2/(alpha*alpha))/denom);
}

int prepareEterm(Context& c) {
    const CellGeom& g = c.geom;
    if (c.etermBox[0] == g.box[0] && c.etermBox[1] == g.box[1] && c.etermBox[2] == g.box[2]) return NBS_OK;
    const int nx = c.grid[0], ny = c.grid[1], nz = c.grid[2], nzh = nz/2 + 1;
    const size_t total = (size_t) nx*ny*nzh;
    NBS_CUDA_CHECK(c.dEterm.ensure(total));
    k_eterm<<<(unsigned) ((total + 255)/256), 256, 0, c.stream>>>(nx, ny, nz, nzh, make_double3(g.invBox[0], g.invBox[1], g.invBox[2]),
                                                                  g.box[0]*g.box[1]*g.box[2], c.alpha, c.dModuli.d, c.dEterm.d);
    c.launches++;
    for (int k = 0; k < 3; k++) c.etermBox[k] = g.box[k];
    return NBS_OK;
}

int launchPme(Context& c, bool wantEnergy) {
    const int nx = c.grid[0], ny = c.grid[1], nz = c.grid[2], nzh = nz/2 + 1;
    const size_t G = (size_t) nx*ny*nz;
    cudaStream_t st = c.stream;
    FftPlan px, py, pz;
    if (!makePlan(nx, px) || !makePlan(ny, py) || !makePlan(nz, pz)) {
        setError("PME grid dimensions must be <= 512 and factor into 2, 3, 5, 7, 11, 13");
        return NBS_ERR_UNSUPPORTED;
    }
    int status = prepareEterm(c);
    if (status != NBS_OK) return status;
    NBS_CUDA_CHECK(cudaMemsetAsync(c.dGrid.d, 0, sizeof(float)*G*c.nS, st));
    PmeArgs p;
    p.N = c.N; p.Npad = c.Npad; p.nS = c.nS; p.nx = nx; p.ny = ny; p.nz = nz; p.nzh = nzh;
    p.posq = c.dPosq.d; p.par = c.dPar.d; p.grid = c.dGrid.d; p.gridC = c.dGridC.d; p.eterm = c.dEterm.d;
    p.force = c.dForce.d; p.energy = c.dEnergy.d;
    for (int k = 0; k < 3; k++) p.fscale[k] = (float) (c.grid[k]*c.geom.invBox[k]);
    for (int s = 0; s < MAX_SLICES; s++) {
        p.lam.c[s] = s < c.nSl ? (float) c.lambdas[2*s] : 1.f;
        p.lam.v[s] = s < c.nSl ? (float) c.lambdas[2*s+1] : 1.f;
    }
    const int atomCtas = (c.N + 7)/8;
    k_spread<<<atomCtas, 256, 0, st>>>(p);
    c.launches++;
    timerMark(c, "spread");

    FftArgs f;
    f.nS = c.nS; f.nx = nx; f.ny = ny; f.nz = nz; f.nzh = nzh;
    f.grid = c.dGrid.d; f.gridC = c.dGridC.d; f.eterm = c.dEterm.d; f.energy = c.dEnergy.d;
    f.wantEnergy = wantEnergy ? 1 : 0;
    f.lam = p.lam;
    const int pairs = c.nS*nx*((ny + 1)/2);
    // z forward
    f.plan = pz; f.tw = c.dTwiddle.d + nx + ny;
    const size_t smZ = sizeof(float2)*(size_t) nz*9;
    // y
    const size_t smY = sizeof(float2)*((size_t) ny + 16*((size_t) ny + 1));
    const size_t smX = sizeof(float2)*((size_t) nx + (size_t) c.nS*8*((size_t) nx + 1));
    if (!c.fftAttrSet) {
        cudaFuncSetAttribute(k_fft_z_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_z_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_y<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_y<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_x_conv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_x_conv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_x_conv<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_x_conv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        cudaFuncSetAttribute(k_fft_x_conv<MAX_SUBSETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
        c.fftAttrSet = true;
    }
    if (smX > 200*1024 || smY > 200*1024 || smZ > 200*1024) {
        setError("PME grid too large for the shared-memory FFT");
        return NBS_ERR_UNSUPPORTED;
    }
    k_fft_z_fwd<<<(pairs + 7)/8, 256, smZ, st>>>(f);
    f.plan = py; f.tw = c.dTwiddle.d + nx;
    k_fft_y<false><<<c.nS*nx*((nzh + 15)/16), 512, smY, st>>>(f);
    f.plan = px; f.tw = c.dTwiddle.d;
    const int xCtas = ny*((nzh + 7)/8);
    switch (c.nS) {
        case 1: k_fft_x_conv<1><<<xCtas, 256, smX, st>>>(f); break;
        case 2: k_fft_x_conv<2><<<xCtas, 256, smX, st>>>(f); break;
        case 3: k_fft_x_conv<3><<<xCtas, 256, smX, st>>>(f); break;
        case 4: k_fft_x_conv<4><<<xCtas, 256, smX, st>>>(f); break;
        default: k_fft_x_conv<MAX_SUBSETS><<<xCtas, 256, smX, st>>>(f); break;
    }
    f.plan = py; f.tw = c.dTwiddle.d + nx;
    k_fft_y<true><<<c.nS*nx*((nzh + 15)/16), 512, smY, st>>>(f);
    f.plan = pz; f.tw = c.dTwiddle.d + nx + ny;
    k_fft_z_inv<<<(pairs + 7)/8, 256, smZ, st>>>(f);
    c.launches += 5;
    timerMark(c, "fft_conv");
    k_gather<<<atomCtas, 256, 0, st>>>(p);
    c.launches++;
    timerMark(c, "gather");
    return NBS_OK;
}

} // namespace nbs

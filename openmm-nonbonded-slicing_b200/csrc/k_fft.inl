// k_fft.inl -- plane-fused 3D real FFT for the sliced PME path (used whenever a (y, kz) plane of one
// subset fits in shared memory; k_pme.cu keeps the line-at-a-time kernels for larger grids).
//
// The reference calls a library here (cuFFT / VkFFT batched over subsets: platforms/cuda/src/
// CudaCuFFT3D.cpp:47-83, CudaVkFFT3D.cpp:17-79; pocketfft c2c on the Reference platform,
// ReferencePME.cpp:793-805).  PME grids are small (64^3 .. 180^3): the transform is bound by launch
// latency and by round trips through L2/HBM, not by flops.  So the chain is three kernels with one
// global read and one global write each:
//   k_fft_zy_fwd : CTA = one (subset, x) plane.  Real rows -> shared, z transform (two real rows ride
//                  one complex line), unpack to half spectra, y transform down the columns, store.
//   k_fft_x_conv2: CTA = one y, a chunk of kz, ALL subsets: forward x transform, influence function,
//                  per-slice structure-factor products E_IJ (ReferencePME.cpp:487-491), lambda mixing
//                  G_I = eterm * sum_J lambda_IJ S_J, inverse x transform.
//   k_fft_yz_inv : CTA = one (subset, x) plane: inverse y, pack, inverse z, float potential grid out.
// Inside a CTA a pass is BATCHED over all lines of the plane: every thread owns whole radix-R
// butterflies (radices 2, 3, 4, 5, 7; 11 and 13 in a rolled form), reads them into registers, the CTA
// synchronises, and the results are written back in place (Stockham autosort order).
#include "nbs_internal.h"
#include "nbs_device.cuh"
#include "k_fft.cuh"
#include <algorithm>
#include <cstdlib>
// compiled twice (k_fft_f32.cu, k_fft_f64.cu define NBS_FFT_REAL) so that the two precisions build in parallel

namespace nbs {

// Division by a run-time constant without the ~30-instruction integer divide: q = trunc((x + 0.5) * (1/d)) in
// fp32 is exact while x < 2^20 (the relative error 2^-22 of the product stays below the 0.5/d margin), which
// covers every index in this file (all below 2^16).  A thread owns only one or two butterflies of a few dozen
// instructions per pass, so four integer divides per butterfly were most of the kernel -- and so would be a
// 64-bit magic-number computation per pass.
struct FastDiv {
    int d;
    float inv;
    __device__ __forceinline__ explicit FastDiv(int divisor) : d(divisor), inv(1.0f/(float) divisor) {}
    __device__ __forceinline__ int div(int x) const { return __float2int_rz(((float) x + 0.5f)*inv); }
    __device__ __forceinline__ void divmod(int x, int& q, int& r) const { q = div(x); r = x - q*d; }
};

// Global -> shared copies keep FFT_COPY_U loads in flight per thread: with one load per loop trip (what the
// compiler emits for a plain strided loop) the kernels spent a fifth of their time waiting on L2 latency here.
constexpr int FFT_COPY_U = 4;

template <typename T> struct Cx2;
template <> struct Cx2<float> { typedef float2 type; };
template <> struct Cx2<double> { typedef double2 type; };
__device__ __forceinline__ float2 mkc(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ double2 mkc(double x, double y) { return make_double2(x, y); }
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) { return mkc(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }
template <typename C> __device__ __forceinline__ C caddc(C a, C b) { return mkc(a.x + b.x, a.y + b.y); }
template <typename C> __device__ __forceinline__ C csubc(C a, C b) { return mkc(a.x - b.x, a.y - b.y); }

// One Stockham pass of radix R, batched over `lines` lines of length n, executed by the whole CTA.
// Element e of line L lives at base[L*lineStride + e*elemStride].  Work item = one butterfly; consecutive
// threads take consecutive butterflies of a line when the line is contiguous (elemStride == 1) and
// consecutive lines otherwise, which keeps shared-memory accesses on distinct banks.  The pass is in
// place: a ROUND covers as many whole lines as the CTA has threads for, reads (and twiddles) their
// butterflies into registers, synchronises, and only then computes the size-R DFTs and writes them
// back one output at a time.  Lines are independent, so rounds need no other ordering.
template <int R, typename C>
__device__ __forceinline__ void batchedPass(C* base, int lines, int lineStride, int elemStride, int n, int Ns, const C* tw) {
    const int nb = n/R;
    const int tstep = n/(Ns*R), rstep = n/R;
    const int os = Ns*elemStride;
    const int linesPerRound = max(1, (int) blockDim.x/nb);       // host guarantees nb <= blockDim.x
    const FastDiv divNb(nb), divNs(Ns);
    for (int l0 = 0; l0 < lines; l0 += linesPerRound) {
        const int nl = min(linesPerRound, lines - l0);
        const int wi = threadIdx.x;
        C v[R];
        int dst = -1;
        if (wi < nl*nb) {
            int L, j;
            if (elemStride == 1) divNb.divmod(wi, L, j);
            else FastDiv(nl).divmod(wi, j, L);
            L += l0;
            const C* line = base + (size_t) L*lineStride;
#pragma unroll
            for (int t = 0; t < R; t++) v[t] = line[(j + t*nb)*elemStride];
            int jq, k;
            divNs.divmod(j, jq, k);
            if (Ns > 1) {
#pragma unroll
                for (int t = 1; t < R; t++) v[t] = cmulc(v[t], tw[t*k*tstep]);
            }
            dst = L*lineStride + (jq*Ns*R + k)*elemStride;
        }
        __syncthreads();
        if (dst >= 0) {
            if (R == 2) {
                base[dst] = caddc(v[0], v[1]);
                base[dst + os] = csubc(v[0], v[1]);
            }
            else if (R == 4) {
                const C a0 = caddc(v[0], v[2]), a1 = csubc(v[0], v[2]);
                const C a2 = caddc(v[1], v[3]), a3 = csubc(v[1], v[3]);
                base[dst] = caddc(a0, a2);
                base[dst + os] = mkc(a1.x + a3.y, a1.y - a3.x);        // a1 - i a3
                base[dst + 2*os] = csubc(a0, a2);
                base[dst + 3*os] = mkc(a1.x - a3.y, a1.y + a3.x);      // a1 + i a3
            }
            else {
#pragma unroll
                for (int o = 0; o < R; o++) {
                    C acc = v[0];
#pragma unroll
                    for (int t = 1; t < R; t++) {
                        const C w = tw[((o*t) % R)*rstep];               // exp(-2 pi i (o t mod R) / R)
                        acc.x += v[t].x*w.x - v[t].y*w.y;
                        acc.y += v[t].x*w.y + v[t].y*w.x;
                    }
                    base[dst + o*os] = acc;
                }
            }
        }
        __syncthreads();
    }
}

// Radix 11 / 13 (rare grid sizes): the same pass with rolled loops and the staged inputs in local memory.
template <typename C>
__device__ __noinline__ void batchedPassLarge(C* base, int lines, int lineStride, int elemStride, int n, int R, int Ns, const C* tw) {
    const int nb = n/R, tstep = n/(Ns*R), rstep = n/R;
    const int os = Ns*elemStride;
    const int linesPerRound = max(1, (int) blockDim.x/nb);
    const FastDiv divNb(nb), divNs(Ns);
    for (int l0 = 0; l0 < lines; l0 += linesPerRound) {
        const int nl = min(linesPerRound, lines - l0);
        const int wi = threadIdx.x;
        C v[13];
        int dst = -1;
        if (wi < nl*nb) {
            int L, j;
            if (elemStride == 1) divNb.divmod(wi, L, j);
            else FastDiv(nl).divmod(wi, j, L);
            L += l0;
            const C* line = base + (size_t) L*lineStride;
            int jq, k;
            divNs.divmod(j, jq, k);
            for (int t = 0; t < R; t++) {
                C x = line[(j + t*nb)*elemStride];
                v[t] = t == 0 ? x : cmulc(x, tw[t*k*tstep]);
            }
            dst = L*lineStride + (jq*Ns*R + k)*elemStride;
        }
        __syncthreads();
        if (dst >= 0) {
            for (int o = 0; o < R; o++) {
                C acc = v[0];
                for (int t = 1; t < R; t++) {
                    const C z = tw[((o*t) % R)*rstep];
                    acc.x += v[t].x*z.x - v[t].y*z.y;
                    acc.y += v[t].x*z.y + v[t].y*z.x;
                }
                base[dst + o*os] = acc;
            }
        }
        __syncthreads();
    }
}

// Forward (e^{-i...}) unnormalised FFT of `lines` lines, by the whole CTA.
// RMAX = largest radix this instantiation carries code for (4, 5 or 13): small-radix plans get kernels
// with fewer registers and more CTAs per SM.
template <int RMAX, typename C>
__device__ __forceinline__ void batchedFft(C* base, int lines, int lineStride, int elemStride, int n,
                                           unsigned long long factors, const C* tw) {
    int Ns = 1;
    for (; factors != 0; factors >>= 4) {
        const int R = (int) (factors & 15);
        if (R == 4) batchedPass<4>(base, lines, lineStride, elemStride, n, Ns, tw);
        else if (R == 2) batchedPass<2>(base, lines, lineStride, elemStride, n, Ns, tw);
        else if (R == 3) batchedPass<3>(base, lines, lineStride, elemStride, n, Ns, tw);
        else if (RMAX >= 5 && R == 5) batchedPass<5>(base, lines, lineStride, elemStride, n, Ns, tw);
        else if (RMAX >= 13 && R == 7) batchedPass<7>(base, lines, lineStride, elemStride, n, Ns, tw);
        else if (RMAX >= 13) batchedPassLarge(base, lines, lineStride, elemStride, n, R, Ns, tw);
        Ns *= R;
    }
}

// ---------------------------------------------------------------------------------------------
// zy forward: real grid [s][x][y][z] -> half spectrum [s][x][y][kz], transformed along z and y.
// Shared plane: ny rows of `rs` complex numbers (rs >= nz/2 + 1, and 2*rs >= nz so that the complex line
// of a row pair fits in the pair's two rows).
// ---------------------------------------------------------------------------------------------
template <typename T, int RMAX>
__global__ void __launch_bounds__(FFT_THREADS) k_fft_zy_fwd(const PlaneFftArgs a) {
    typedef typename Cx2<T>::type C;
    extern __shared__ double2 fftSmem[];
    C* sm = (C*) fftSmem;
    const int ny = a.ny, nz = a.nz, nzh = a.nzh, rs = a.rowStride;
    C* twz = sm;
    C* twy = sm + nz;
    C* plane = sm + nz + ny;
    const FastDiv divNz(nz), divNzh(nzh);
    for (int k = threadIdx.x; k < nz; k += blockDim.x) twz[k] = ((const C*) a.twz)[k];
    for (int k = threadIdx.x; k < ny; k += blockDim.x) twy[k] = ((const C*) a.twy)[k];
    // CTA = one (own subset, x) plane, or -- when a plane does not fit in shared memory -- a slab of its row pairs
    // (then the y transform is a separate kernel, k_fft_y_cols)
    const int plane_ = blockIdx.x/a.slabsPerPlane, slab = blockIdx.x - plane_*a.slabsPerPlane;
    const int sx = (a.ownLo + plane_/a.nxOwn)*a.nx + a.xLo + plane_ % a.nxOwn;     // (subset, x) of this rank's plane number plane_
    const bool fused = a.slabsPerPlane == 1;
    const int allPairs = (ny + 1) >> 1;
    const int pBase = slab*a.slabPairs;
    const int pairs = min(a.slabPairs, allPairs - pBase);       // row pairs this CTA holds (local rows 0 .. 2*pairs)
    const int rowBase = 2*pBase, rowsHere = min(2*pairs, ny - rowBase);
    const T* grid = (const T*) a.grid + ((size_t) sx*ny + rowBase)*nz;
    for (int base = threadIdx.x; base < pairs*nz; base += FFT_COPY_U*blockDim.x) {
        T re[FFT_COPY_U], im[FFT_COPY_U];
        int at[FFT_COPY_U];
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++) {
            const int idx = base + u*blockDim.x;
            at[u] = -1;
            if (idx < pairs*nz) {
                int p, z;
                divNz.divmod(idx, p, z);
                re[u] = grid[(size_t) (2*p)*nz + z];
                im[u] = 2*p + 1 < rowsHere ? grid[(size_t) (2*p + 1)*nz + z] : (T) 0;
                at[u] = p*2*rs + z;
            }
        }
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++)
            if (at[u] >= 0) plane[at[u]] = mkc(re[u], im[u]);
    }
    __syncthreads();
    batchedFft<RMAX>(plane, pairs, 2*rs, 1, nz, a.factorsZ, twz);
    // unpack Z -> the two rows' half spectra: S0[k] = (Z[k] + conj Z[n-k])/2, S1[k] = (Z[k] - conj Z[n-k])/(2i).
    // In place, so a round stages whole row pairs in registers before anything is written.
    {
        const int pairsPerRound = max(1, ((int) blockDim.x*FFT_UNPACK_Q)/nzh);
        const T half = (T) 0.5;
        for (int p0 = 0; p0 < pairs; p0 += pairsPerRound) {
            const int count = min(pairsPerRound, pairs - p0)*nzh;
            C zk[FFT_UNPACK_Q], zn[FFT_UNPACK_Q];
#pragma unroll
            for (int q = 0; q < FFT_UNPACK_Q; q++) {
                const int wi = threadIdx.x + q*blockDim.x;
                if (wi < count) {
                    int p, k;
                    divNzh.divmod(wi, p, k);
                    p += p0;
                    const C* line = plane + (size_t) p*2*rs;
                    zk[q] = line[k];
                    zn[q] = line[k == 0 ? 0 : nz - k];
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < FFT_UNPACK_Q; q++) {
                const int wi = threadIdx.x + q*blockDim.x;
                if (wi < count) {
                    int p, k;
                    divNzh.divmod(wi, p, k);
                    p += p0;
                    plane[(size_t) (2*p)*rs + k] = mkc(half*(zk[q].x + zn[q].x), half*(zk[q].y - zn[q].y));
                    if (2*p + 1 < rowsHere) plane[(size_t) (2*p + 1)*rs + k] = mkc(half*(zk[q].y + zn[q].y), -half*(zk[q].x - zn[q].x));
                }
            }
            __syncthreads();
        }
    }
    if (fused) batchedFft<RMAX>(plane, nzh, 1, rs, ny, a.factorsY, twy);
    C* out = (C*) a.gridC + ((size_t) sx*ny + rowBase)*nzh;
    for (int idx = threadIdx.x; idx < rowsHere*nzh; idx += blockDim.x) {
        int y, k;
        divNzh.divmod(idx, y, k);
        out[idx] = plane[(size_t) y*rs + k];
    }
}

// ---------------------------------------------------------------------------------------------
// y transform alone (planes too large to fuse): CTA = one (subset, x) plane x `colChunk` kz columns, in place.
// ---------------------------------------------------------------------------------------------
template <typename T, int RMAX, bool INVERSE>
__global__ void __launch_bounds__(FFT_THREADS) k_fft_y_cols(const PlaneFftArgs a) {
    typedef typename Cx2<T>::type C;
    extern __shared__ double2 fftSmem[];
    C* sm = (C*) fftSmem;
    const int ny = a.ny, nzh = a.nzh, cw = a.colChunk, rs = cw + 1;
    C* twy = sm;
    C* cols = sm + ny;
    for (int k = threadIdx.x; k < ny; k += blockDim.x) twy[k] = ((const C*) a.twy)[k];
    const int chunks = (nzh + cw - 1)/cw;
    const int plane_ = blockIdx.x/chunks, k0 = (blockIdx.x - plane_*chunks)*cw;
    const int kn = min(cw, nzh - k0);
    const FastDiv divKn(kn);
    C* base = (C*) a.gridC + (size_t) ((a.ownLo + plane_/a.nxOwn)*a.nx + a.xLo + plane_ % a.nxOwn)*ny*nzh + k0;
    for (int b0 = threadIdx.x; b0 < ny*kn; b0 += FFT_COPY_U*blockDim.x) {
        C v[FFT_COPY_U];
        int at[FFT_COPY_U];
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++) {
            const int idx = b0 + u*blockDim.x;
            at[u] = -1;
            if (idx < ny*kn) {
                int y, l;
                divKn.divmod(idx, y, l);
                v[u] = base[(size_t) y*nzh + l];
                at[u] = y*rs + l;
            }
        }
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++)
            if (at[u] >= 0) {
                if (INVERSE) v[u].y = -v[u].y;
                cols[at[u]] = v[u];
            }
    }
    __syncthreads();
    batchedFft<RMAX>(cols, kn, 1, rs, ny, a.factorsY, twy);
    for (int idx = threadIdx.x; idx < ny*kn; idx += blockDim.x) {
        int y, l;
        divKn.divmod(idx, y, l);
        C v = cols[(size_t) y*rs + l];
        if (INVERSE) v.y = -v.y;
        base[(size_t) y*nzh + l] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// yz inverse: half spectrum (already inverse-transformed along x) -> real potential grid (float).
// Inverse transforms are forward transforms of the conjugate.
// ---------------------------------------------------------------------------------------------
template <typename T, int RMAX>
__global__ void __launch_bounds__(FFT_THREADS) k_fft_yz_inv(const PlaneFftArgs a) {
    typedef typename Cx2<T>::type C;
    extern __shared__ double2 fftSmem[];
    C* sm = (C*) fftSmem;
    const int ny = a.ny, nz = a.nz, nzh = a.nzh, rs = a.rowStride;
    C* twz = sm;
    C* twy = sm + nz;
    C* plane = sm + nz + ny;
    const FastDiv divNz(nz), divNzh(nzh);
    for (int k = threadIdx.x; k < nz; k += blockDim.x) twz[k] = ((const C*) a.twz)[k];
    for (int k = threadIdx.x; k < ny; k += blockDim.x) twy[k] = ((const C*) a.twy)[k];
    const int plane_ = blockIdx.x/a.slabsPerPlane, slab = blockIdx.x - plane_*a.slabsPerPlane;
    const int sx = (a.ownLo + plane_/a.nxOwn)*a.nx + a.xLo + plane_ % a.nxOwn;     // (subset, x) of this rank's plane number plane_
    const bool fused = a.slabsPerPlane == 1;
    const int allPairs = (ny + 1) >> 1;
    const int pBase = slab*a.slabPairs;
    const int pairs = min(a.slabPairs, allPairs - pBase);
    const int rowBase = 2*pBase, rowsHere = min(2*pairs, ny - rowBase);
    const C* in = (const C*) a.gridC + ((size_t) sx*ny + rowBase)*nzh;
    for (int b0 = threadIdx.x; b0 < rowsHere*nzh; b0 += FFT_COPY_U*blockDim.x) {
        C v[FFT_COPY_U];
        int at[FFT_COPY_U];
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++) {
            const int idx = b0 + u*blockDim.x;
            at[u] = -1;
            if (idx < rowsHere*nzh) {
                int y, k;
                divNzh.divmod(idx, y, k);
                v[u] = in[idx];
                at[u] = y*rs + k;
            }
        }
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++)
            if (at[u] >= 0) {
                v[u].y = -v[u].y;                                // conjugate: inverse y = conj(fwd(conj))
                plane[at[u]] = v[u];
            }
    }
    __syncthreads();
    if (fused) batchedFft<RMAX>(plane, nzh, 1, rs, ny, a.factorsY, twy);
    // (split path: k_fft_y_cols<INVERSE> already produced A; the load above conjugated it, as the fused path leaves it)
    // plane now holds conj(A) where A = y-inverse spectrum.  Pack rows (2p, 2p+1) into one complex line:
    // W[k] = conj(A0[k] + i A1[k]),  W[n-k] = conj(conj(A0[k]) + i conj(A1[k]))   (0 < k, 2k < n)
    // so that fwd(W) = conj(r0 + i r1) with r0, r1 the two real rows.
    {
        const int pairsPerRound = max(1, ((int) blockDim.x*FFT_UNPACK_Q)/nzh);
        for (int p0 = 0; p0 < pairs; p0 += pairsPerRound) {
            const int count = min(pairsPerRound, pairs - p0)*nzh;
            C a0[FFT_UNPACK_Q], a1[FFT_UNPACK_Q];
#pragma unroll
            for (int q = 0; q < FFT_UNPACK_Q; q++) {
                const int wi = threadIdx.x + q*blockDim.x;
                if (wi < count) {
                    int p, k;
                    divNzh.divmod(wi, p, k);
                    p += p0;
                    C u = plane[(size_t) (2*p)*rs + k];
                    u.y = -u.y;                                      // A0[k]
                    C v = mkc((T) 0, (T) 0);
                    if (2*p + 1 < rowsHere) { v = plane[(size_t) (2*p + 1)*rs + k]; v.y = -v.y; }    // A1[k]
                    a0[q] = u; a1[q] = v;
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < FFT_UNPACK_Q; q++) {
                const int wi = threadIdx.x + q*blockDim.x;
                if (wi < count) {
                    int p, k;
                    divNzh.divmod(wi, p, k);
                    p += p0;
                    C* line = plane + (size_t) p*2*rs;
                    const C A = a0[q], B = a1[q];
                    line[k] = mkc(A.x - B.y, -(A.y + B.x));                           // conj(A + iB)
                    if (k > 0 && 2*k < nz) line[nz - k] = mkc(A.x + B.y, A.y - B.x);  // conj(conj(A) + i conj(B))
                }
            }
            __syncthreads();
        }
        batchedFft<RMAX>(plane, pairs, 2*rs, 1, nz, a.factorsZ, twz);
        float* pot = a.pot + ((size_t) sx*ny + rowBase)*nz;
        double* potD = (double*) a.pot + ((size_t) sx*ny + rowBase)*nz;       // NBS_FLAG_DOUBLE: a double potential grid
        for (int idx = threadIdx.x; idx < pairs*nz; idx += blockDim.x) {
            int p, z;
            divNz.divmod(idx, p, z);
            const C wv = plane[(size_t) p*2*rs + z];
            if (a.potDouble) {
                potD[(size_t) (2*p)*nz + z] = (double) wv.x;
                if (2*p + 1 < rowsHere) potD[(size_t) (2*p + 1)*nz + z] = (double) -wv.y;
            }
            else {
                pot[(size_t) (2*p)*nz + z] = (float) wv.x;
                if (2*p + 1 < rowsHere) pot[(size_t) (2*p + 1)*nz + z] = (float) -wv.y;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// x transform + sliced convolution + x inverse.  CTA = one y x `chunk` consecutive kz, ALL subsets.
// Convolution and energies: pme_reciprocal_convolution, ReferencePME.cpp:400-496 -- eterm per k,
// E[slice(I,I)] += 1/2 eterm |S_I|^2, E[slice(I,J)] += eterm Re(S_I conj S_J) over the FULL grid (the
// half spectrum counts twice except on the kz = 0 and kz = nz/2 planes).  The reference then scales
// every subset grid by eterm and lets the gather mix subsets with lambda (:681-687); here the mix
// happens in k space.  Shared layout: lines[x][s*chunk + l] -- the transform runs down the columns, so a
// pass touches consecutive banks and the global <-> shared copies are straight.
// ---------------------------------------------------------------------------------------------
template <typename T, int RMAX, int NS>
__global__ void __launch_bounds__(FFT_X_THREADS) k_fft_x_conv2(const PlaneFftArgs a) {
    typedef typename Cx2<T>::type C;
    extern __shared__ double2 fftSmem[];
    C* sm = (C*) fftSmem;
    __shared__ double shE[MAX_SLICES];
    const int n = a.nx, nzh = a.nzh, chunk = a.chunk, nS = a.nS;
    const int RS = nS*chunk + 1;                             // row stride: lines[x][s*chunk + l]
    C* tw = sm;
    C* lines = sm + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) tw[k] = ((const C*) a.twx)[k];
    if (threadIdx.x < MAX_SLICES) shE[threadIdx.x] = 0.0;
    const int chunks = (nzh + chunk - 1)/chunk;
    const int yLocal = blockIdx.x/chunks, k0 = (blockIdx.x - yLocal*chunks)*chunk;
    const int y = a.yLo + yLocal;
    const int kn = min(chunk, nzh - k0);                     // kz values this CTA really has
    const FastDiv divKn(kn), divChunk(chunk), divN(n);
    // plane x lives in the spectra of rank ((x + 1) R - 1) / nx, the owner of the slab [r nx / R, (r+1) nx / R)
    const int nRanks = a.nRanks;
    auto planeOf = [&](int s, int x) -> C* {
        C* base = nRanks == 1 ? (C*) a.gridC : (C*) a.peerSpectra[((x + 1)*nRanks - 1)/n];
        return base + (((size_t) s*n + x)*a.ny + y)*nzh + k0;
    };
    for (int b0 = threadIdx.x; b0 < nS*n*chunk; b0 += FFT_COPY_U*blockDim.x) {
        C v[FFT_COPY_U];
        int at[FFT_COPY_U];
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++) {
            const int idx = b0 + u*blockDim.x;
            at[u] = -1;
            if (idx < nS*n*chunk) {
                int rest, l, s, x;
                divChunk.divmod(idx, rest, l);
                divN.divmod(rest, s, x);
                v[u] = l < kn ? planeOf(s, x)[l] : mkc((T) 0, (T) 0);
                at[u] = x*RS + s*chunk + l;
            }
        }
#pragma unroll
        for (int u = 0; u < FFT_COPY_U; u++)
            if (at[u] >= 0) lines[at[u]] = v[u];
    }
    // the influence function of this CTA's (x, y, kz) points, fetched now so that its latency hides behind the forward transform
    T* etS = (T*) (lines + (size_t) n*RS);
    for (int idx = threadIdx.x; idx < n*kn; idx += blockDim.x) {
        int x, l;
        divKn.divmod(idx, x, l);
        etS[idx] = ((const T*) a.eterm)[((size_t) x*a.ny + y)*nzh + k0 + l];
    }
    __syncthreads();
    batchedFft<RMAX>(lines, nS*chunk, 1, RS, n, a.factorsX, tw);
    double e[NS*(NS+1)/2];
#pragma unroll
    for (int s = 0; s < NS*(NS+1)/2; s++) e[s] = 0.0;
    for (int idx = threadIdx.x; idx < n*kn; idx += blockDim.x) {
        int x, l;
        divKn.divmod(idx, x, l);
        const int k = k0 + l;
        const T et = etS[idx];
        C S[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) S[s] = s < nS ? lines[(size_t) x*RS + s*chunk + l] : mkc((T) 0, (T) 0);
        if (a.wantEnergy) {
            const T wgt = (k == 0 || 2*k == a.nz) ? (T) 1 : (T) 2;
#pragma unroll
            for (int sb = 0; sb < NS; sb++)
#pragma unroll
                for (int sa = 0; sa <= sb; sa++) {
                    if (sa < a.ownLo || sa >= a.ownHi) continue;     // a slice belongs to the owner of its lower subset
                    const T prod = S[sa].x*S[sb].x + S[sa].y*S[sb].y;
                    e[sb*(sb+1)/2 + sa] += (double) ((sa == sb ? (T) 0.5 : (T) 1)*wgt*et*prod);
                }
        }
#pragma unroll
        for (int si = 0; si < NS; si++) {
            if (si >= a.ownHi) break;
            if (si < a.ownLo) continue;
            T gx = 0, gy = 0;
#pragma unroll
            for (int sj = 0; sj < NS; sj++) {
                const T lam = (T) a.lam.c[triSlice(si, sj)];
                gx += lam*S[sj].x;
                gy += lam*S[sj].y;
            }
            lines[(size_t) x*RS + si*chunk + l] = mkc(et*gx, -et*gy);      // conjugated for the inverse pass
        }
    }
    __syncthreads();
    const int nOwn = a.ownHi - a.ownLo;
    batchedFft<RMAX>(lines + (size_t) a.ownLo*chunk, nOwn*chunk, 1, RS, n, a.factorsX, tw);
    for (int idx = threadIdx.x; idx < nOwn*n*kn; idx += blockDim.x) {
        int rest, l, s, x;
        divKn.divmod(idx, rest, l);
        divN.divmod(rest, s, x);
        s += a.ownLo;
        C v = lines[(size_t) x*RS + s*chunk + l];
        v.y = -v.y;
        planeOf(s, x)[l] = v;
    }
    if (a.wantEnergy) {
        const int lane = threadIdx.x & 31;
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
// host side
// ---------------------------------------------------------------------------------------------
template <typename T, int RMAX>
static int launchPlaneT(Context& c, PlaneFftArgs a, int half, size_t smPlane, size_t smX) {
    cudaStream_t st = c.stream;
    static bool attr[64] = {false};
    if (!attr[c.device & 63]) {
        const int big = 220*1024;
        cudaFuncSetAttribute(k_fft_zy_fwd<T, RMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_yz_inv<T, RMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_y_cols<T, RMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_y_cols<T, RMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv2<T, RMAX, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv2<T, RMAX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv2<T, RMAX, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv2<T, RMAX, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_fft_x_conv2<T, RMAX, MAX_SUBSETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        attr[c.device & 63] = true;
    }
    const int nOwn = c.ownHi - c.ownLo;
    const int planes = nOwn*a.nxOwn;
    const bool fused = a.slabsPerPlane == 1;
    const int colCtas = planes*((a.nzh + a.colChunk - 1)/a.colChunk);
    const size_t smCols = sizeof(typename Cx2<T>::type)*((size_t) a.ny + (size_t) a.ny*(a.colChunk + 1));
    if (half == 0) {
        k_fft_zy_fwd<T, RMAX><<<planes*a.slabsPerPlane, a.planeThreads, smPlane, st>>>(a);
        c.launches++;
        if (!fused) {
            k_fft_y_cols<T, RMAX, false><<<colCtas, a.colThreads, smCols, st>>>(a);
            c.launches++;
        }
        return NBS_OK;
    }
    // half 1 = x pass + inverse transforms; half 2 = x pass only, half 3 = inverse transforms only (slab sharding puts
    // a barrier between them: the x pass writes into other ranks' planes)
    const int xCtas = a.nyOwn*((a.nzh + a.chunk - 1)/a.chunk);
    if (half != 3) {
        switch (c.nS) {
            case 1: k_fft_x_conv2<T, RMAX, 1><<<xCtas, a.xThreads, smX, st>>>(a); break;
            case 2: k_fft_x_conv2<T, RMAX, 2><<<xCtas, a.xThreads, smX, st>>>(a); break;
            case 3: k_fft_x_conv2<T, RMAX, 3><<<xCtas, a.xThreads, smX, st>>>(a); break;
            case 4: k_fft_x_conv2<T, RMAX, 4><<<xCtas, a.xThreads, smX, st>>>(a); break;
            default: k_fft_x_conv2<T, RMAX, MAX_SUBSETS><<<xCtas, a.xThreads, smX, st>>>(a); break;
        }
        c.launches++;
    }
    if (half == 2) return NBS_OK;
    if (!fused) {
        k_fft_y_cols<T, RMAX, true><<<colCtas, a.colThreads, smCols, st>>>(a);
        c.launches++;
    }
    k_fft_yz_inv<T, RMAX><<<planes*a.slabsPerPlane, a.planeThreads, smPlane, st>>>(a);
    c.launches++;
    return NBS_OK;
}

// Returns NBS_OK after launching, or NBS_RETRY (>0) when the grid does not fit this path (the caller
// falls back to the line-at-a-time kernels).
template <typename T>
int launchPlaneFft(Context& c, const PlaneFftPlan& plan, PlaneFftArgs a, int half) {
    typedef typename Cx2<T>::type C;
    const int nx = c.grid[0], ny = c.grid[1], nz = c.grid[2], nzh = nz/2 + 1;
    const int rs = std::max(nzh, (nz + 1)/2) + 1;
    // x pass: choose the kz chunk so that the chunks are even and the lines of all subsets fit
    int xTarget = 8;
    if (const char* env = getenv("NBS_FFT_XCHUNK")) xTarget = std::max(1, atoi(env));                 // tuning experiments
    int chunks = (nzh + xTarget - 1)/xTarget;
    int chunk = (nzh + chunks - 1)/chunks;
    const size_t cs = sizeof(C);
    while (chunk > 1 && cs*((size_t) nx + (size_t) nx*(c.nS*chunk + 1)) + sizeof(T)*nx*chunk > 200*1024) chunk--;
    const size_t smX = cs*((size_t) nx + (size_t) nx*(c.nS*chunk + 1)) + sizeof(T)*(size_t) nx*chunk;
    if (smX > 200*1024) return NBS_RETRY;
    // zy / yz kernels: the whole plane when it fits (y transform fused in), else slabs of row pairs of <= ~64 KB
    const int pairsZ = (ny + 1)/2;
    size_t smPlane = cs*((size_t) nz + ny + (size_t) pairsZ*2*rs);
    a.slabPairs = pairsZ;
    a.slabsPerPlane = 1;
    size_t slabBytes = 64*1024;
    if (const char* env = getenv("NBS_FFT_SLAB_KB")) { slabBytes = (size_t) atoi(env)*1024; smPlane = 1u << 30; }   // tuning experiments: force the split path
    if (smPlane > 200*1024) {
        a.slabPairs = std::max(1, (int) ((slabBytes/cs - nz - ny)/(2*rs)));
        a.slabsPerPlane = (pairsZ + a.slabPairs - 1)/a.slabPairs;
        a.slabPairs = (pairsZ + a.slabsPerPlane - 1)/a.slabsPerPlane;          // even slabs
        smPlane = cs*((size_t) nz + ny + (size_t) a.slabPairs*2*rs);
    }
    a.colChunk = 16;
    while (a.colChunk > 1 && cs*((size_t) ny + (size_t) ny*(a.colChunk + 1)) > 64*1024) a.colChunk--;
    // CTA sizes: the widest pass (lines x n/R butterflies) should run in the fewest rounds of equal size
    auto pickThreads = [](int lines, int n, unsigned long long factors, int cap) {
        int rmin = 16;
        for (unsigned long long f = factors; f != 0; f >>= 4) rmin = std::min(rmin, (int) (f & 15));
        const int nb = n/rmin;                                 // butterflies per line in the widest pass
        const int rounds = (lines*nb + cap - 1)/cap;
        const int linesPerRound = (lines + rounds - 1)/rounds;
        return std::min(cap, std::max(64, ((linesPerRound*nb + 31)/32)*32));
    };
    int planeCap = 256;                      // measured at 64^3 double: 256-thread CTAs beat 512 (20.5 vs 24.6 us zy forward)
    if (const char* env = getenv("NBS_FFT_PLANE_THREADS")) planeCap = std::max(128, std::min(FFT_THREADS, atoi(env)));   // tuning experiments
    a.planeThreads = pickThreads(a.slabPairs, nz, plan.factors[2], planeCap);
    if (a.slabsPerPlane == 1) a.planeThreads = std::max(a.planeThreads, pickThreads(nzh, ny, plan.factors[1], planeCap));
    a.colThreads = pickThreads(a.colChunk, ny, plan.factors[1], FFT_THREADS);
    int xCap = FFT_X_THREADS;
    if (const char* env = getenv("NBS_FFT_X_THREADS")) xCap = std::max(64, std::min(FFT_X_THREADS, atoi(env)));           // tuning experiments
    a.xThreads = pickThreads(c.nS*chunk, nx, plan.factors[0], xCap);
    if (nzh > 128*FFT_UNPACK_Q || nz/2 > a.planeThreads || ny/2 > std::min(a.planeThreads, a.colThreads) || nx/2 > a.xThreads) return NBS_RETRY;
    a.rowStride = rs;
    a.chunk = chunk;
    a.factorsX = plan.factors[0]; a.factorsY = plan.factors[1]; a.factorsZ = plan.factors[2];
    const C* tw = (const C*) (sizeof(T) == 8 ? (const void*) c.dTwiddleD.d : (const void*) c.dTwiddle.d);
    a.twx = tw; a.twy = tw + nx; a.twz = tw + nx + ny;
    // the plane kernels see only the own slabs
    const int nOwn = c.ownHi - c.ownLo;
    (void) nOwn;
    int rmax = 2;
    for (int d = 0; d < 3; d++)
        for (unsigned long long f = plan.factors[d]; f != 0; f >>= 4) rmax = std::max(rmax, (int) (f & 15));
    if (rmax <= 4) return launchPlaneT<T, 4>(c, a, half, smPlane, smX);
    if (rmax <= 5) return launchPlaneT<T, 5>(c, a, half, smPlane, smX);
    return launchPlaneT<T, 13>(c, a, half, smPlane, smX);
}

template int launchPlaneFft<NBS_FFT_REAL>(Context&, const PlaneFftPlan&, PlaneFftArgs, int);

} // namespace nbs

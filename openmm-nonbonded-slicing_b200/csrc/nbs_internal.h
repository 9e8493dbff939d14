// nbs_internal.h -- context object and kernel launch prototypes of libnbslice_b200.so.
//
// Data layout in HBM (all "sorted" arrays are in cell order: column-major over an (x, y) grid of
// columns, z-bin order inside a column, original particle index inside a z-bin):
//   posq   uint4  [N]   x, y, z as 32-bit fixed-point FRACTIONAL coordinates (frac * 2^32; periodic
//                       wrap is integer overflow), w = charge * sqrt(ONE_4PI_EPS0) as float bits
//   par    float4 [N]   x = sigma/2, y = 2*sqrt(eps), z = subset (int bits), w = particle index (int bits)
//   force  int64  [3][Npad]  fixed-point (value * 2^32) accumulators, sorted order -- deterministic sums
//   blocks: i-blocks of <= 32 consecutive sorted atoms that never straddle a column
//   jlist  int32  [nBlocks][capJ]  neighbour atoms of an i-block: (shiftCode << 26) | sortedIndex
//   xlist / xmask [nBlocks][capX]  neighbours that carry an exclusion mask (own block, exclusions)
//   grid   float  [nS][nx][ny][nz]      real-space charge grid, later the lambda-mixed potential grid
//   gridC  float2 [nS][nx][ny][nz/2+1]  half-complex spectrum
#ifndef NBS_INTERNAL_H_
#define NBS_INTERNAL_H_

#include "nbslice_b200.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

namespace nbs {

constexpr double kPi = 3.14159265358979323846;
constexpr double kECharge = 1.602176634e-19;
constexpr double kAvogadro = 6.02214076e23;
constexpr double kEpsilon0 = 1e-6*8.8541878128e-12/(kECharge*kECharge*kAvogadro);
constexpr double kOne4PiEps0 = 1/(4*kPi*kEpsilon0);

constexpr int MAX_SUBSETS = 8;
constexpr int MAX_SLICES = MAX_SUBSETS*(MAX_SUBSETS+1)/2;
constexpr int PME_ORDER = 5;
// Table of f(s) = erfc(alpha sqrt(s))/sqrt(s), s = r^2, for the double-precision pair energies: 128 intervals per octave
// of s starting at s = 2^-7; row = exponent and top seven mantissa bits of the double s.  A row holds c0 = P(0) in DOUBLE
// and c1..c4 in SINGLE precision of the degree-4 interpolant P(d) = c0 + d (c1 + d (c2 + d (c3 + d c4))) at Chebyshev nodes,
// d in [-1/2, 1/2) the position inside the interval (the next 23 mantissa bits of s).  The remainder d (c1 + ...) is at
// most 3 % of f, so evaluating it in single precision costs 3e-9 of f (rounding, zero mean; bias per row below 3e-10:
// tools/erfc_table_check.py restates the scheme in NumPy) while the energy path needs ONE double-precision product per
// pair instead of an eight-term double-precision Horner chain and two shared-memory loads instead of eight.
constexpr int ERFC_TAB_PER_OCTAVE_LOG2 = 7;
constexpr int ERFC_TAB_MAX_ROWS = 1280;          // 10 octaves: up to s = 2^3 nm^2; pairs beyond take the analytic branch
// A list entry is (image code << J_SHIFT_BITS) | sorted index; image code = (kx+2) + 5 ((ky+1) + 3 (kz+1)) with
// kx in -2..2 (a triclinic box's b and c vectors shift x by up to ax/2 each), ky, kz in -1..1: 45 codes, 6 bits.
constexpr int J_SHIFT_BITS = 25;                 // sorted index in the low 25 bits of a list entry
constexpr int J_INDEX_MASK = (1 << J_SHIFT_BITS)-1;
constexpr int BUILD_WARPS = 8;                   // warps per CTA in the list-build kernel
#ifndef NBS_PAIR_WARPS
#define NBS_PAIR_WARPS 8
#endif
constexpr int PAIR_WARPS = NBS_PAIR_WARPS;       // warps per CTA in the pair kernel

void setError(const std::string& message);
#define NBS_CUDA_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    nbs::setError(std::string(#call) + ": " + cudaGetErrorString(e_)); return NBS_ERR_CUDA; } } while (0)

extern unsigned long long gAllocEpoch;       // bumped whenever a device buffer moves (invalidates captured graphs)

template <class T>
struct Buf {
    T* d = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        gAllocEpoch++;
        if (d) cudaFree(d);
        d = nullptr;
        cap = 0;
        size_t want = n + n/8 + 64;
        cudaError_t e = cudaMalloc((void**) &d, want*sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (d) cudaFree(d); d = nullptr; cap = 0; }
};

struct LambdaTable {                 // passed by value to kernels
    float c[MAX_SLICES];             // Coulomb scale per slice
    float v[MAX_SLICES];             // vdW scale per slice
};

struct CellGeom {                    // cell / column geometry of one evaluation
    // Triclinic boxes a = (ax, 0, 0), b = (bx, by, 0), c = (cx, cy, cz) (OpenMM's reduced form) are handled in the
    // rectangular BRICK [0, ax) x [0, by) x [0, cz), which is a unit cell of the same lattice: atoms are wrapped into
    // it with the lattice translations (k_prep), so cells, bounding boxes and all distances stay plain Cartesian, and
    // only the periodic IMAGE shifts change: image (kx, ky, kz) is displaced by kx a + ky b + kz c.
    double box[3];                   // ax, by, cz (non-periodic methods: a virtual box twice the size of the system)
    double tilt[3];                  // bx, cx, cy (nm); all zero for a rectangular box
    long long shiftB, shiftCx, shiftCy;   // the same in fixed-point units of their axis: bx/ax, cx/ax, cy/by times 2^32
    bool triclinic;
    double origin[3];                // coordinate that maps to fractional 0 (0 for periodic boxes)
    double invBox[3];
    float scale[3];                  // L / 2^32 (fixed-point unit in nm)
    int ncx, ncy, nzb;               // columns in x, y; z-bins per column
    int nCols, nBins;
    float colW[2];                   // column widths (nm)
    float binH;                      // z-bin height (nm)
};

constexpr int ENERGY_WORDS = 2*MAX_SLICES + 8;   // slice table + [2*MAX_SLICES] = list-overflow flag (as a double, so it all-reduces)

// Peer-memory sharding: one mailbox per rank, reachable by every rank of the node (CUDA IPC / same process).
struct PeerMailbox {
    unsigned long long arrive[NBS_MAX_RANKS];      // arrive[r] = number of barriers rank r has reached (written BY rank r)
    unsigned long long epoch;                      // barriers this rank has reached (local)
    unsigned long long timedOut;                   // epoch of a barrier that gave up waiting (0 = none)
    double energies[NBS_MAX_RANKS][ENERGY_WORDS];  // energies[r] = rank r's slice-energy table of the evaluation in flight
};

struct KernelTimer {
    std::vector<const char*> names;
    std::vector<cudaEvent_t> events;      // events[i], events[i+1] bracket names[i]
};

struct Context {
    // ---- description (host) ----
    int N = 0, nS = 0, nSl = 0, method = 0, device = 0;
    uint32_t flags = 0;
    double cutoff = 0, alpha = 0, switchDist = 0, rfDielectric = 78.3;
    bool useSwitch = false, excPeriodic = false;
    bool periodic = true;                    // CutoffPeriodic / Ewald / PME; false for NoCutoff / CutoffNonPeriodic
    bool ewaldDirect() const { return method == NBS_METHOD_PME || method == NBS_METHOD_EWALD || method == NBS_METHOD_LJPME; }   // erfc direct space + exclusion corrections
    bool usesPmeGrid() const { return method == NBS_METHOD_PME || method == NBS_METHOD_LJPME; }
    bool ljpme() const { return method == NBS_METHOD_LJPME; }
    double cutoffEff = 0;                    // the cutoff, or (NoCutoff) a distance no pair of the system exceeds
    int grid[3] = {0, 0, 0};
    // LJPME: the dispersion grid's own (alpha, grid, tables); swapped with the Coulomb set around the dispersion pass
    double dispAlpha = 0;
    int dispGrid[3] = {0, 0, 0};
    bool dispersionPass = false;             // launchPme is running the dispersion (C6) chain
    std::vector<double> subsetC6Self;        // sum over the subset's atoms of c6^2 / 12 (self term, :212)
    int ewaldKmax[3] = {0, 0, 0};            // Ewald: reciprocal vectors per axis (numRx, numRy, numRz)
    int ewaldNK = 0;                         // vectors in the half space
    std::vector<int> subsets;
    std::vector<double> baseQ, baseSig, baseEps;
    int nExc = 0, num14 = 0, nGlobals = 0;
    std::vector<int> excPairs;               // [nExc][2]
    std::vector<double> excParams;           // [nExc][3] base
    std::vector<int> pOffIdx, eOffIdx;       // [k][2]
    std::vector<double> pOffScale, eOffScale;// [k][3]
    std::vector<double> globals;             // current values
    std::vector<double> dispersion;          // [nSl]
    std::vector<double> lambdas;             // [nSl][2] current
    bool paramsDirty = true;
    // derived on the host from the offset-applied parameters
    std::vector<double> subsetQ, subsetQ2;   // sum q, sum q^2 per subset
    // ---- static device data ----
    Buf<int> dSubset;
    Buf<float> dChargeF;                     // q*sqrt(K)
    Buf<float2> dSigEps;
    Buf<double2> dSigEpsD;                   // NBS_FLAG_DOUBLE: (sigma/2, 2 sqrt(eps)) in double, particle order
    Buf<double> dLamD;                       // NBS_FLAG_DOUBLE: the lambda table in double
    Buf<double> dCharge;                     // q (double)
    Buf<float> dC6F;                         // LJPME: c6 = 8 (sigma/2)^3 2 sqrt(eps) per particle (:395-396), the dispersion grid's "charge"
    Buf<double> dC6D;
    Buf<int> dExclStart, dExclList;          // CSR over particles (symmetric)
    Buf<int2> dExcPair;                      // [nExc]
    Buf<double4> dExcParam;                  // [nExc] (sigma, 4eps, K*qq, is14 ? 1 : 0)
    Buf<int> dExcSlice;
    // ---- per-evaluation device data ----
    Buf<float> dBBox;                        // [6] bounding box of the input positions (non-periodic methods)
    Buf<double> dPosIn;                      // staged positions when the caller's are on the host
    Buf<double> dForceOut;                   // staged forces when the caller's are on the host
    Buf<uint4> dFix;                         // original order: fixed-point xyz, w = bin
    Buf<uint4> dFixBuild;                    // sorted order: coordinates at the last list build (re-used lists measure displacements from them)
    Buf<int> dBinCount, dBinStart, dBinCursor;
    Buf<int> dScanTmp;
    Buf<int> dSortedToOrig, dOrigToSorted;
    Buf<uint4> dPosq;
    Buf<double> dQ64;                        // sorted charges * sqrt(ONE_4PI_EPS0) in double (energy path)
    Buf<float4> dPar;
    Buf<int> dColBlockStart;                 // [nCols+1]
    Buf<int> dBlkFirst, dBlkCount;
    Buf<uint4> dBlkLo, dBlkHi;               // fixed-point bounding box; lo.w = column index
    Buf<int2> dExclRange;                    // per sorted atom: [min, max] sorted index of its exclusion partners
    Buf<int> dJList, dJCount, dXList, dXCount;
    Buf<unsigned> dGmJ, dGmX;                // cluster masks of the list entries' groups of 8: one word per tile of 32 entries
    Buf<int4> dItems;                        // pair-kernel work items (local block, first tile, first atom, atom count)
    Buf<unsigned> dXMask;
    Buf<int> dCounters;                      // [0] nBlocks, [1] overflow flag, [2] pair work items, [3] work cursor
    Buf<unsigned long long> dForce;          // [3][Npad] in sorted order (+ [3][Npad] in particle order: unsorted-PME gather)
    Buf<double> dEnergy;                     // [MAX_SLICES][2]
    Buf<double> dGrid;                       // charge grid [nS][nx][ny][nz] (double or float view)
    Buf<unsigned long long> dGridFixed;      // NBS_FLAG_DETERMINISTIC: fixed-point accumulation grid of the spreading
    Buf<double2> dGridC;                     // half spectrum [nS][nx][ny][nz/2+1] (double2 or float2 view)
    Buf<float> dPot;                         // potential grid read by the gather
    Buf<float> dEterm;
    Buf<double> dEtermD;
    Buf<double2> dTwiddleD;
    Buf<double> dModuli;                     // [nx+ny+nz]
    Buf<double> dErfcTab;                    // c0[erfcRows] (double) followed by (c1, c2, c3, c4)[erfcRows] (float4), see ERFC_TAB_*
    int erfcRows = 0;
    Buf<float2> dTwiddle;                    // [nx+ny+nz]
    // LJPME keeps a second set of the tables that depend on (alpha, grid)
    Buf<float> dEtermDisp; Buf<double> dEtermDDisp, dModuliDisp; Buf<double2> dTwiddleDDisp; Buf<float2> dTwiddleDisp;
    double etermBoxDisp[6] = {0, 0, 0, 0, 0, 0};
    std::vector<double> hModuliDisp;
    Buf<int4> dEwaldK;                       // Ewald: half space of reciprocal vectors (rx, ry, rz, 0)
    Buf<double2> dEwaldSums, dEwaldMixed;    // [nK][MAX_SUBSETS] structure factors; lambda-mixed factors
    Buf<unsigned long long> dPairStats;      // [0] count, [1] hash
    Buf<int2> dPairDump;
    // host mirrors
    int* hCounters = nullptr;                // pinned
    double* hEnergy = nullptr;               // pinned
    int capJ = 0, capX = 0, maxBlocks = 0, Npad = 0;
    int nBlocksLast = 0;
    CellGeom geom{};
    double etermBox[6] = {0, 0, 0, 0, 0, 0};  // box (diagonal, tilt) the influence function was computed for
    double lastBox[9] = {0};
    bool haveLast = false, lastDirect = false;
    std::vector<double> hModuli;
    // bookkeeping
    long long launches = 0;
    KernelTimer timer;
    bool profiling = false;
    bool fftAttrSet = false;
    double* hForce = nullptr;                // pinned staging for host force output
    size_t hForceCap = 0;
    cudaStream_t stream = nullptr;           // stream of the current nbs_execute
    long long stats[8] = {0};
    // ---- sharding across ranks (one process per GPU; see DESIGN.md "Multi-GPU") ----
    // Direct space: global i-block b belongs to this rank iff (b % blockPeriod) lies in
    // [blockOffset, blockOffset + blockWidth).  Lists are stored at the rank-local block index.
    // PME: this rank owns the charge grids of subsets [ownLo, ownHi).
    int rank = 0, nRanks = 1;
    int blockPeriod = 1, blockOffset = 0, blockWidth = 1;
    int ownLo = 0, ownHi = 0;                // set to [0, nS) at creation
    int maxLocalBlocks = 0;
    // peer-memory sharding (nbs_set_slab_shard): PME split by x-slabs over all ranks, x pass and force reduction over
    // NVLink peer memory, barriers over flags in the peers' mailboxes
    bool slabMode = false, peersImported = false, peerBarrier = false;
    int xLo = 0, xHi = 0, yLo = 0, yHi = 0;     // own grid planes; own rows of the x pass
    void* peerSpectra[NBS_MAX_RANKS] = {nullptr};
    unsigned long long* peerForce[NBS_MAX_RANKS] = {nullptr};
    PeerMailbox* peerMailbox[NBS_MAX_RANKS] = {nullptr};
    void* peerOpened[3*NBS_MAX_RANKS] = {nullptr};          // IPC mappings to close
    Buf<PeerMailbox> dMailbox;
    unsigned long long* hTimedOut = nullptr;    // pinned copy of the mailbox's timedOut word
    int slabStep = 0;                           // next step of the evaluation in flight
    int slabRange[4] = {0, 0, 0, 0};            // sorted atoms that can reach the own planes: [0],[1]) and [2],[3]) (host copy of counters[8..11])
    bool slabRangeValid = false;
    bool pmeUnsorted = false;                // PME works from particle-order coordinates (forks before the sort)
    int chunkTiles = 2;                      // tiles per pair-kernel work item
    // ---- neighbour-list re-use (periodic cutoff methods): the list is built with cutoff + skin and kept until an
    // atom has moved more than skin/2 since the build (checked on the device, acted upon by the host) ----
    double skin = 0;                         // nm; 0 = rebuild on every evaluation, like the Reference platform
    bool listValid = false;                  // sort order (+ lists, if listHasDirect) of a previous evaluation can be re-used
    bool listHasDirect = false;
    double listBox[9] = {0};
    unsigned long long listEpoch = 0, listEpochAtBuild = 0;     // bumped by anything that invalidates the list
    unsigned long long listAllocEpoch = 0;
    int listCapJ = 0;
    bool reuseNow = false, forceRebuild = false;
    double dispLast = 0, dispStepMax = 0;    // nm: max displacement at the last evaluation; largest growth per evaluation seen
    long long evalCount = 0, buildCount = 0, redoCount = 0;
    int numSMs = 148;
    // ---- phase state of the evaluation in flight ----
    cudaStream_t ownStream = nullptr;        // used when the caller passes the legacy default stream
    cudaStream_t directStream = nullptr;     // direct space runs here, concurrently with PME on `stream`
    cudaEvent_t evSorted = nullptr, evDirectDone = nullptr;
    cudaStream_t auxStream = nullptr;        // exceptions / exclusion corrections, beside the list build
    cudaEvent_t evAuxFork = nullptr, evAuxDone = nullptr;
    cudaEvent_t evPlaced = nullptr, evExclDone = nullptr;   // k_excl_range on auxStream, beside the block construction
    bool sideFork = false, exclRangeForked = false;
    bool phaseDirect = false, phaseRecip = false, phaseEnergy = false, directOverlapped = false;
    const double* phasePos64 = nullptr;
    int phase = 0;                           // 0 idle, 1 begun, 2 convolved
    // ---- CUDA graph of a whole single-rank evaluation (replayed while nothing it depends on changes) ----
    // (two cached graphs: slot 0 = evaluations that build the list, slot 1 = evaluations that re-use it)
    cudaGraphExec_t graphExecs[2] = {nullptr, nullptr};
    unsigned long long graphKeys[2] = {0, 0}, warmKeys[2] = {0, 0}, paramVersion = 0;
    long long graphLaunchCounts[2] = {0, 0}; // kernels per graph replay (for the launch counter)
};

__host__ __device__ inline int localToGlobalBlock(int local, int period, int offset, int width) {
    return (local/width)*period + offset + local % width;
}

// ---- launch wrappers (each counts its launches in ctx.launches) ----
struct PosInput { const void* ptr; int format; const int* atomIndex; double* pos64out; };

int launchBBox(Context& c, const PosInput& in, float out[6]);    // bounding box of device positions (synchronises)
int launchPrep(Context& c, const PosInput& in);                 // fixed-point conversion, bin histogram
int launchSortRest(Context& c);                                 // cell sort, sorted records, i-blocks
int launchReprep(Context& c, const PosInput& in);               // re-used list: new coordinates into the old sort order
int launchBuildLists(Context& c);
int launchExclRange(Context& c);
int launchPairs(Context& c, bool wantEnergy, int mode);         // mode 0: forces+energy, 1: count/hash pairs, 2: dump pairs
int launchBonded(Context& c, const double* dPos, bool periodicBox);
int launchPme(Context& c, bool wantEnergy, int half);    // half 0: spread..y forward; 1: x/conv..gather
int launchFinalize(Context& c, void* dOut, int format, long long paddedAtoms, int accumulate, const int* atomIndex);
int launchPeerBarrier(Context& c, bool publishEnergies);        // k_peer.cu
int launchPeerReduce(Context& c);
int launchEwald(Context& c, bool wantEnergy);               // plain Ewald reciprocal sum (k_ewald.cu)
int uploadEwaldVectors(Context& c);
int prepareEterm(Context& c);
void swapPmeTables(Context& c);                              // Coulomb <-> dispersion (alpha, grid, tables)
int uploadPmeTables(Context& c);

void timerMark(Context& c, const char* name);   // records an event when profiling

} // namespace nbs
#endif

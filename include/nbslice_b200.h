/*
 * nbslice_b200.h -- C ABI of the B200-native SlicedNonbondedForce hot path.
 *
 * This is the drop-in boundary below the plugin's CalcSlicedNonbondedForceKernel
 * interface (reference: openmmapi/include/NonbondedSlicingKernels.h:27-85).  The C++
 * platform kernel in openmm-nonbonded-slicing_b200/platform/ subclasses that interface and
 * forwards to the entry points declared here; nothing in these signatures depends on
 * OpenMM, torch or C++ types -- plain pointers, sizes and status codes only.
 *
 * Mapping to the reference interface (file:line relative to the reference checkout):
 *
 *   nbs_create              <- CalcSlicedNonbondedForceKernel::initialize
 *                              (NonbondedSlicingKernels.h:48; what the Reference platform
 *                              collects from the Force in ReferenceNonbondedSlicingKernels.cpp:59-185)
 *   nbs_execute             <- CalcSlicedNonbondedForceKernel::execute
 *                              (NonbondedSlicingKernels.h:59; ReferenceNonbondedSlicingKernels.cpp:187-268)
 *   nbs_update_parameters   <- CalcSlicedNonbondedForceKernel::copyParametersToContext
 *                              (NonbondedSlicingKernels.h:66; ReferenceNonbondedSlicingKernels.cpp:270-319)
 *   nbs_get_pme_parameters  <- CalcSlicedNonbondedForceKernel::getPMEParameters
 *                              (NonbondedSlicingKernels.h:75; ReferenceNonbondedSlicingKernels.cpp:321-328)
 *   nbs_get_ljpme_parameters<- CalcSlicedNonbondedForceKernel::getLJPMEParameters
 *                              (NonbondedSlicingKernels.h:84; ReferenceNonbondedSlicingKernels.cpp:330-337)
 *   nbs_set_lambdas /
 *   nbs_set_global_parameters
 *                           <- the per-evaluation reads of Context parameters in
 *                              ReferenceNonbondedSlicingKernels.cpp:339-392 (computeParameters)
 *   nbs_destroy             <- KernelImpl destructor
 *   nbs_last_error          <- the OpenMMException messages the adapter rethrows
 *                              (SURVEY 8b "Errors")
 *
 * Conventions (reference: SURVEY Appendix A):
 *   units nm, kJ/mol, e;  slice(i,j) = max*(max+1)/2 + min  (SlicedNonbondedForce.h:22);
 *   term index 0 = Coulomb, 1 = van der Waals (ReferenceNonbondedSlicingKernels.h:76-77);
 *   slice energies are UNSCALED by lambda; forces are lambda-scaled
 *   (ReferenceSlicedLJCoulombIxn.cpp:435-444).
 *
 * All functions return NBS_OK (0) or a negative status; the message for the calling
 * thread's last failure is available from nbs_last_error().  There is no CPU fallback:
 * every compute entry point fails with NBS_ERR_CUDA if no usable sm_100 device exists.
 */
#ifndef NBSLICE_B200_H_
#define NBSLICE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBS_ABI_VERSION 4

/* status codes */
#define NBS_OK                 0
#define NBS_ERR_INVALID       -1   /* bad argument / inconsistent description            */
#define NBS_ERR_UNSUPPORTED   -2   /* valid for the reference, not implemented on device */
#define NBS_ERR_CUDA          -3   /* CUDA runtime failure, or no sm_100 device          */
#define NBS_ERR_BOX           -4   /* periodic box smaller than twice the cutoff
                                      (ReferenceNonbondedSlicingKernels.cpp:202-204)     */
#define NBS_ERR_CAPACITY      -5   /* internal list overflow that could not be regrown   */
#define NBS_RETRY              1   /* nbs_execute_finish only: an internal list was too small (on
                                      this or another rank); it has been grown -- repeat the
                                      evaluation from nbs_execute_begin                      */

/* nonbonded methods: values of CalcSlicedNonbondedForceKernel::NonbondedMethod
 * (NonbondedSlicingKernels.h:29-36) */
#define NBS_METHOD_NOCUTOFF            0
#define NBS_METHOD_CUTOFF_NONPERIODIC  1
#define NBS_METHOD_CUTOFF_PERIODIC     2
#define NBS_METHOD_EWALD               3
#define NBS_METHOD_PME                 4
#define NBS_METHOD_LJPME               5

/* nbs_system_desc.flags */
#define NBS_FLAG_DETERMINISTIC   0x1u  /* fixed-point PME spreading (DeterministicForces)   */
#define NBS_FLAG_PROFILE         0x2u  /* record CUDA events around every kernel            */
#define NBS_FLAG_NO_GRAPH        0x4u  /* plain stream launches instead of a CUDA graph     */
#define NBS_FLAG_LINE_FFT        0x10u /* always use the line-at-a-time FFT kernels (the path for
                                          grids whose planes exceed shared memory); test hook     */
#define NBS_FLAG_SORTED_PME      0x20u /* PME always works from the cell-sorted records (the path of
                                          large systems); test hook                              */
#define NBS_FLAG_NO_LIST_REUSE   0x40u /* rebuild the neighbour list on every evaluation, like the
                                          Reference platform (default: built with a skin and kept
                                          until an atom has moved half of it, see nbs_set_list_skin) */
#define NBS_FLAG_DOUBLE          0x80u /* the plugin's Precision = double (CommonNonbondedSlicingKernels.cpp:297-299; every CUDA
                                          test of the reference runs in single, mixed and double, platforms/cuda/tests/
                                          CMakeLists.txt:22-24): direct-space forces in double precision arithmetic from the
                                          exact fixed-point coordinates, double-precision PME grids, transforms and gather.
                                          Coordinates still carry 32 fractional bits of the box (5e-9 nm at 21 nm), which is
                                          what bounds the agreement with a double-precision reference (measured 2e-8 of the forces) */
#define NBS_FLAG_FP32_ENERGY     0x8u  /* single-precision pair energies and PME grids (the
                                          plugin's "single" precision); default is double
                                          precision for every energy term, fp32 for forces    */

/* memory spaces / layouts for positions and forces */
#define NBS_MEM_HOST    0
#define NBS_MEM_DEVICE  1

#define NBS_POS_F64_XYZ   0   /* double[N][3]  -- the Reference platform's vector<Vec3>          */
#define NBS_POS_F32_XYZW  1   /* float[N][4]   -- OpenMM CUDA posq (w ignored), device only      */
#define NBS_POS_F64_XYZW  2   /* double[N][4]  -- OpenMM CUDA posq in double precision, device   */

#define NBS_FORCE_F64_XYZ      0   /* double[N][3], added to (accumulate=1) or overwritten       */
#define NBS_FORCE_I64_FIXED    1   /* long long[3][padded_atoms], value*2^32, always ADDED;
                                      OpenMM CUDA's getLongForceBuffer layout (pme.cc:382-388)   */

/*
 * Everything the Reference platform's initialize() reads from the Force and System
 * (ReferenceNonbondedSlicingKernels.cpp:59-185).  All arrays are host memory and are
 * copied; the caller keeps ownership.
 */
typedef struct nbs_system_desc {
    int32_t struct_size;                  /* = sizeof(nbs_system_desc)                         */
    int32_t num_particles;
    int32_t num_subsets;
    int32_t method;                       /* NBS_METHOD_*                                       */
    const int32_t* subsets;               /* [N] getParticleSubset                              */
    const double*  charges;               /* [N] base charge                                    */
    const double*  sigmas;                /* [N] base sigma                                     */
    const double*  epsilons;              /* [N] base epsilon                                   */
    int32_t num_exceptions;               /* ALL exceptions; each one is also an exclusion      */
    int32_t num_global_params;            /* values supplied through nbs_set_global_parameters  */
    const int32_t* exception_particles;   /* [nE][2]                                            */
    const double*  exception_params;      /* [nE][3] chargeProd, sigma, epsilon (base)          */
    int32_t num_particle_offsets;
    int32_t num_exception_offsets;
    const int32_t* particle_offset_indices;   /* [nPO][2] (global parameter index, particle)    */
    const double*  particle_offset_scales;    /* [nPO][3] chargeScale, sigmaScale, epsilonScale */
    const int32_t* exception_offset_indices;  /* [nEO][2] (global parameter index, exception)   */
    const double*  exception_offset_scales;   /* [nEO][3]                                       */
    double cutoff;
    double switching_distance;
    double rf_dielectric;
    double ewald_alpha;                   /* from calcPMEParameters (explicit)       */
    int32_t pme_grid[3];
    int32_t use_switching_function;
    int32_t exceptions_use_periodic;
    int32_t device_index;
    uint32_t flags;                       /* NBS_FLAG_*                                         */
    int32_t reserved0;
    const double* dispersion_coefficients;/* [nSl] from SlicedNonbondedForceImpl::
                                             calcDispersionCorrections, or NULL for zeros       */
    int32_t ewald_kmax[3];                /* NBS_METHOD_EWALD: number of reciprocal vectors per
                                             axis, numRx/numRy/numRz of setUseEwald
                                             (ReferenceSlicedLJCoulombIxn.cpp:115-121), from
                                             calcEwaldParameters (ReferenceNonbondedSlicingKernels.
                                             cpp:160-162)                                        */
    int32_t reserved1;
    double  dispersion_alpha;             /* NBS_METHOD_LJPME: setUseLJPME (:149-155)            */
    int32_t dispersion_grid[3];
    int32_t reserved2;
} nbs_system_desc;

/* One evaluation == one CalcSlicedNonbondedForceKernel::execute call. */
typedef struct nbs_exec_args {
    int32_t struct_size;                  /* = sizeof(nbs_exec_args)                           */
    int32_t positions_format;             /* NBS_POS_*                                          */
    int32_t positions_space;              /* NBS_MEM_*                                          */
    int32_t forces_format;                /* NBS_FORCE_*                                        */
    int32_t forces_space;                 /* NBS_MEM_*                                          */
    int32_t forces_accumulate;            /* F64 only: 1 = add into the buffer (Reference
                                             semantics, forces += ...), 0 = overwrite           */
    const void* positions;
    void* forces;                         /* may be NULL: energies only                         */
    int64_t padded_num_atoms;             /* I64_FIXED: stride between the x, y and z planes    */
    const int32_t* atom_index;            /* optional device int[N]: slot -> particle index of
                                             the caller's (re-ordered) position/force buffers   */
    double box[9];                        /* periodic box vectors a, b, c (rows)                */
    int32_t include_forces;               /* ignored, like the Reference platform               */
    int32_t include_energy;               /* only affects the adapter's return value            */
    int32_t include_direct;
    int32_t include_reciprocal;
    double* slice_energies;               /* host double[nSl][2] (Coulomb, vdW), overwritten;
                                             includes self, background and dispersion terms     */
    void* stream;                         /* cudaStream_t the work is ordered on (0 = default)  */
} nbs_exec_args;

typedef struct nbs_context nbs_context;

/* library-level */
int         nbs_abi_version(void);
const char* nbs_last_error(void);
int         nbs_device_count(void);

/* life cycle */
int nbs_create(const nbs_system_desc* desc, nbs_context** out);
int nbs_destroy(nbs_context* ctx);
int nbs_update_parameters(nbs_context* ctx, const nbs_system_desc* desc);

/* per-evaluation state */
int nbs_set_lambdas(nbs_context* ctx, const double* lambdas /* [nSl][2] (Coulomb, vdW) */);
int nbs_set_global_parameters(nbs_context* ctx, const double* values /* [num_global_params] */);
int nbs_execute(nbs_context* ctx, const nbs_exec_args* args);

/*
 * Multi-GPU evaluation: one process (and one nbs_context) per GPU, every rank holding ALL positions.
 * The reference has no counterpart -- its only multi-device path is OpenMM's per-device work split with
 * reciprocal space on device 0 (platforms/cuda/src/CudaParallelNonbondedSlicingKernels.cpp:35-53,
 * CommonNonbondedSlicingKernels.cpp:416, 465, 643-646).  Work is split where it shards naturally:
 *   direct space : i-block b belongs to the rank with (b % block_period) in [block_offset,
 *                  block_offset + block_width); exceptions are dealt round-robin;
 *   PME          : a rank owns the charge grids of subsets [subset_begin, subset_end) -- spreading,
 *                  FFTs and gather of those subsets; an empty range means no reciprocal work.
 * The caller (torch.distributed / NCCL) performs the two exchanges between the phases:
 *   nbs_execute_begin -> broadcast each owned half spectrum to the ranks that own grids
 *   nbs_execute_convolve -> all-reduce(sum) `forces` (int64) and `energies` (double)
 *   nbs_execute_finish   (returns NBS_RETRY on every rank if any rank overflowed a list)
 * nbs_execute == begin + convolve + finish on an unsharded context.
 */
typedef struct nbs_exchange_buffers {
    int32_t struct_size;                  /* = sizeof(nbs_exchange_buffers)                    */
    int32_t spectrum_is_double;           /* element type of `spectra`: double2 (1) or float2   */
    void*   spectra;                      /* device [num_subsets][nx][ny][nz/2+1] half spectra  */
    int64_t spectrum_bytes_per_subset;
    void*   forces;                       /* device int64[force_words]: fixed-point (x 2^32)
                                             accumulators [3][padded] in cell-sorted order, which
                                             is identical on every rank                          */
    int64_t force_words;
    void*   energies;                     /* device double[energy_words]: slice table + flags    */
    int64_t energy_words;
} nbs_exchange_buffers;

int nbs_set_shard(nbs_context* ctx, int32_t rank, int32_t num_ranks, int32_t block_period, int32_t block_offset,
                  int32_t block_width, int32_t subset_begin, int32_t subset_end);
int nbs_execute_begin(nbs_context* ctx, const nbs_exec_args* args);
int nbs_execute_convolve(nbs_context* ctx, const nbs_exec_args* args);
int nbs_execute_finish(nbs_context* ctx, const nbs_exec_args* args);
/* valid between nbs_execute_begin and nbs_execute_finish of the evaluation in flight */
int nbs_get_exchange_buffers(nbs_context* ctx, nbs_exchange_buffers* out);

/*
 * Peer-memory sharding (PME; one process per GPU on one NVLink / NVSwitch node).  The split above leaves a subset's
 * reciprocal work on ONE rank; this one splits it over ALL ranks and needs no collective library on the data path:
 *   direct space : i-blocks as above (block_period / block_offset / block_width);
 *   PME          : rank r owns the x-SLAB [r nx / R, (r+1) nx / R) of every subset grid: it spreads the atoms that
 *                  touch its planes (positions are replicated, so no halo is exchanged), transforms its planes along
 *                  z and y, runs the fused x pass (forward x, sliced convolution, slice energies, lambda mixing,
 *                  inverse x) for the y-range [r ny / R, (r+1) ny / R) READING the other ranks' planes and WRITING the
 *                  result back into them through peer memory (the all-to-all transposes of a slab-decomposed FFT,
 *                  fused into the kernel that consumes them), inverse-transforms its planes and gathers the forces
 *                  its planes contribute;
 *   reduction    : one kernel sums the 64-bit fixed-point force accumulators of all ranks for this rank's 1/R of the
 *                  atoms, reading peer memory, and writes the sums back to every rank (reduce-scatter + all-gather);
 *                  slice energies travel through a small mailbox in peer memory.
 * The steps of an evaluation are separated by barriers over flags in peer memory (in_kernel_barrier = 1; then
 * nbs_execute drives the whole sharded evaluation), or by the caller (in_kernel_barrier = 0: call
 * nbs_execute_step(step) for step = 0 .. NBS_NUM_STEPS-1 and make sure every rank has finished step k before any
 * rank starts step k+1 -- what tests do with several contexts on one device).
 * The reference has no counterpart (platforms/cuda/src/CudaParallelNonbondedSlicingKernels.cpp:35-53 keeps
 * reciprocal space on device 0).  PME only; the plane-FFT path only (grids whose x lines fit in shared memory).
 */
#define NBS_MAX_RANKS 16
#define NBS_NUM_STEPS 5
typedef struct nbs_peer_export {
    int32_t struct_size;                  /* = sizeof(nbs_peer_export)                          */
    int32_t rank;
    int64_t process_id;                   /* exporting process: a peer in the same process is reached by pointer,   */
    int32_t device;                       /* one in another process through the CUDA IPC handles below              */
    int32_t reserved;
    void*   spectra;                      /* device pointers in the exporting process                               */
    void*   forces;
    void*   mailbox;
    unsigned char spectra_ipc[64];        /* cudaIpcMemHandle_t of the three allocations                            */
    unsigned char forces_ipc[64];
    unsigned char mailbox_ipc[64];
} nbs_peer_export;

int nbs_set_slab_shard(nbs_context* ctx, int32_t rank, int32_t num_ranks, int32_t block_period, int32_t block_offset,
                       int32_t block_width);
int nbs_export_peer(nbs_context* ctx, nbs_peer_export* out);
/* `all` = the exports of ranks 0 .. count-1 (this rank's own included), count == num_ranks */
int nbs_import_peers(nbs_context* ctx, int32_t count, const nbs_peer_export* all, int32_t in_kernel_barrier);
/* step 0: sort, lists, direct space (own stream), spreading, z/y transforms of the own planes
 *      1: fused x pass over peer memory for the own y-range
 *      2: inverse y/z transforms of the own planes, force gather, join direct space, publish slice energies
 *      3: force reduction over peer memory
 *      4: forces to the caller's layout, energies to the host; returns NBS_RETRY like nbs_execute_finish */
int nbs_execute_step(nbs_context* ctx, const nbs_exec_args* args, int32_t step);

/*
 * Neighbour-list policy of the periodic cutoff methods.  The list is built with cutoff + skin and re-used by later
 * evaluations until some atom has moved more than skin/2 since the build -- measured on the device in every
 * evaluation; an evaluation that finds the limit exceeded is redone with a fresh list before it returns, so
 * results never depend on the policy (the exact cutoff test is the pair kernel's).  This is what the plugin's CUDA
 * platform inherits from OpenMM's NonbondedUtilities (CommonNonbondedSlicingKernels.cpp:721, useNeighborList);
 * the Reference platform rebuilds every time (ReferenceNonbondedSlicingKernels.cpp:197), which skin = 0 or
 * NBS_FLAG_NO_LIST_REUSE selects.  Default skin: 0.07 nm.
 */
int nbs_set_list_skin(nbs_context* ctx, double skin_nm);
/* out: [0] evaluations, [1] evaluations that built a list, [2] evaluations redone because the displacement limit
 * was exceeded, [3] largest displacement (nm) since the build at the last evaluation, [4] skin (nm),
 * [5] 1 if a re-usable list exists, [6] 1 if the last evaluation re-used one, [7] largest per-evaluation growth of [3] */
int nbs_get_list_stats(const nbs_context* ctx, double out[8]);

/* test hook: initial per-block capacities of the neighbour lists (they grow on demand; a tiny value
 * forces the NBS_RETRY path) */
int nbs_debug_set_list_capacity(nbs_context* ctx, int32_t j_capacity, int32_t x_capacity);

/* queries */
int nbs_get_pme_parameters(const nbs_context* ctx, double* alpha, int32_t* nx, int32_t* ny, int32_t* nz);
/* CalcSlicedNonbondedForceKernel::getLJPMEParameters (NonbondedSlicingKernels.h:84;
 * ReferenceNonbondedSlicingKernels.cpp:330-337): the dispersion grid; an error unless the method is LJPME */
int nbs_get_ljpme_parameters(const nbs_context* ctx, double* alpha, int32_t* nx, int32_t* ny, int32_t* nz);
int nbs_get_num_slices(const nbs_context* ctx, int32_t* num_slices);

/*
 * Parity diagnostics for the neighbour list of the LAST nbs_execute with include_direct:
 * the set of interacting pairs (r^2 <= cutoff^2, not excluded), as particle indices with
 * first < second.  `count` and the order-independent `hash` (sum over pairs of
 * nbs_pair_hash(first, second), mod 2^64) are always written; `pairs` (host int32[capacity][2])
 * receives the pairs themselves when non-NULL and large enough.
 */
int nbs_get_pair_set(nbs_context* ctx, int64_t capacity, int32_t* pairs, int64_t* count, uint64_t* hash);
/* the exclusion set the device uses, as sorted unique (first < second) pairs */
int nbs_get_exclusion_set(nbs_context* ctx, int64_t capacity, int32_t* pairs, int64_t* count);

/*
 * Per-kernel timing of the last nbs_execute (requires NBS_FLAG_PROFILE).  Writes up to
 * `capacity` entries; names are static strings.  Returns the number of entries in *count.
 */
int nbs_get_kernel_times(nbs_context* ctx, int32_t capacity, const char** names, float* milliseconds, int32_t* count);
/* number of kernels launched by this library on behalf of ctx since creation */
int nbs_get_launch_count(const nbs_context* ctx, int64_t* launches);
/* neighbour-list statistics of the last evaluation: [0] i-blocks, [1] j entries, [2] tiles,
 * [3] pair evaluations of the pair kernel (32 per step of 4 i atoms x 8 entries; only the clusters a
 * group of entries can reach are stepped), [4] exclusion-list entries */
int nbs_get_nlist_stats(nbs_context* ctx, int64_t stats[8]);

/* Host-only diagnostic (no device is touched): f(s) = erfc(alpha sqrt(s))/sqrt(s) for n values of s = r^2 exactly as the pair
 * kernel's energy path evaluates it from its table (c0 in double + a single-precision remainder); NaN
 * where the kernel takes its analytic branch instead (s < 2^-7 nm^2 or beyond the table). */
int nbs_debug_erfc_table(double alpha, double cutoff, int32_t n, const double* s, double* f);

/* Measured instruction-rate ceilings of `device` (diagnostics for the benchmark's roofline; no reference
 * counterpart): out[0] = dense FP32 FMA rate in TFLOP/s, out[1] = rsqrt.approx rate in Gop/s,
 * out[2] = SM count, out[3] = nominal SM clock in MHz. */
int nbs_measure_peaks(int32_t device, double out[4]);
/* out[0..4]: warp-instructions per clock per SM of DFMA, int32->double, float->double, double->float and
 * int64->double conversions (what bounds the double-precision energy passes of the pair kernel) */
int nbs_measure_dp_rates(int32_t device, double out[8]);

static inline uint64_t nbs_pair_hash(uint32_t first, uint32_t second) {
    uint64_t x = ((uint64_t) first << 32) | second;       /* splitmix64 finaliser */
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

#ifdef __cplusplus
}
#endif
#endif /* NBSLICE_B200_H_ */

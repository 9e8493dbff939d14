/*
 * peer_sharding.c -- the peer-memory sharding of include/nbslice_b200.h from plain C, one process per GPU.
 *
 * What a C or C++ host (an MPI-parallel MD driver, the plugin's own platform code) writes instead of
 * openmm-nonbonded-slicing_b200/multigpu.py: create a context on the rank's GPU, choose the shard, exchange the
 * nbs_peer_export records by any host-side means, map the peers, then call nbs_execute on every rank with the same
 * positions -- barriers, the x pass over the ranks' planes and the force reduction happen inside, over NVLink.
 *
 * The exchange below goes through files in a directory all ranks can see (no MPI in this image); with MPI it is
 *     MPI_Allgather(&mine, sizeof mine, MPI_BYTE, all, sizeof mine, MPI_BYTE, MPI_COMM_WORLD);
 * Compile-checked by tests/test_abi.py (gcc -fsyntax-only); not part of the library.
 *
 *   usage: peer_sharding <rank> <num_ranks> <exchange_dir>      (system description and positions: fill_system below)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "nbslice_b200.h"

#define CHECK(call) do { int s_ = (call); if (s_ != NBS_OK) { fprintf(stderr, "%s: %s\n", #call, nbs_last_error()); exit(1); } } while (0)

/* A user's code fills these from its own topology; here: a salt-like lattice, two subsets, PME. */
static void fill_system(nbs_system_desc* d, int n, int32_t* subsets, double* q, double* sig, double* eps) {
    memset(d, 0, sizeof *d);
    d->struct_size = (int32_t) sizeof *d;
    d->num_particles = n;
    d->num_subsets = 2;
    d->method = NBS_METHOD_PME;
    for (int i = 0; i < n; i++) { subsets[i] = i < n/4 ? 0 : 1; q[i] = i % 2 ? 1.0 : -1.0; sig[i] = 0.25; eps[i] = 0.5; }
    d->subsets = subsets; d->charges = q; d->sigmas = sig; d->epsilons = eps;
    d->cutoff = 1.0;
    d->ewald_alpha = 2.628261;
    d->pme_grid[0] = d->pme_grid[1] = d->pme_grid[2] = 36;
    d->rf_dielectric = 78.3;
}

static void exchange(const char* dir, int rank, int nranks, const nbs_peer_export* mine, nbs_peer_export* all) {
    char path[512], tmp[512];
    snprintf(tmp, sizeof tmp, "%s/export.%d.tmp", dir, rank);
    snprintf(path, sizeof path, "%s/export.%d", dir, rank);
    FILE* f = fopen(tmp, "wb");
    if (!f || fwrite(mine, sizeof *mine, 1, f) != 1) { perror(tmp); exit(1); }
    fclose(f);
    rename(tmp, path);                                  /* atomic: a reader never sees half a record */
    for (int r = 0; r < nranks; r++) {
        snprintf(path, sizeof path, "%s/export.%d", dir, r);
        while ((f = fopen(path, "rb")) == NULL) usleep(1000);
        if (fread(&all[r], sizeof all[r], 1, f) != 1) { perror(path); exit(1); }
        fclose(f);
    }
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s rank num_ranks exchange_dir\n", argv[0]); return 2; }
    const int rank = atoi(argv[1]), nranks = atoi(argv[2]);
    const int side = 16, n = side*side*side;
    const double L = 4.0;
    int32_t* subsets = malloc(sizeof(int32_t)*n);
    double *q = malloc(sizeof(double)*n), *sig = malloc(sizeof(double)*n), *eps = malloc(sizeof(double)*n);
    double *pos = malloc(sizeof(double)*3*n), *frc = calloc(3*(size_t) n, sizeof(double));
    nbs_system_desc desc;
    fill_system(&desc, n, subsets, q, sig, eps);
    desc.device_index = rank % nbs_device_count();      /* one GPU per rank */
    for (int i = 0; i < n; i++) {
        pos[3*i] = (i % side + 0.5)*L/side; pos[3*i+1] = ((i/side) % side + 0.5)*L/side; pos[3*i+2] = (i/(side*side) + 0.5)*L/side;
    }
    nbs_context* ctx = NULL;
    CHECK(nbs_create(&desc, &ctx));
    /* i-blocks: rank r takes blocks b with (b % 64) in [64 r / R, 64 (r + 1) / R); the PME x-slabs follow from the rank */
    const int lo = 64*rank/nranks, hi = 64*(rank + 1)/nranks;
    CHECK(nbs_set_slab_shard(ctx, rank, nranks, 64, lo, hi - lo));
    nbs_peer_export mine, all[NBS_MAX_RANKS];
    memset(&mine, 0, sizeof mine);
    mine.struct_size = (int32_t) sizeof mine;
    CHECK(nbs_export_peer(ctx, &mine));
    exchange(argv[3], rank, nranks, &mine, all);
    CHECK(nbs_import_peers(ctx, nranks, all, /* in_kernel_barrier = */ 1));

    int32_t nsl = 0;
    CHECK(nbs_get_num_slices(ctx, &nsl));
    double* lambdas = malloc(sizeof(double)*2*nsl), *energies = malloc(sizeof(double)*2*nsl);
    for (int k = 0; k < 2*nsl; k++) lambdas[k] = 1.0;
    CHECK(nbs_set_lambdas(ctx, lambdas));
    nbs_exec_args args;
    memset(&args, 0, sizeof args);
    args.struct_size = (int32_t) sizeof args;
    args.positions_format = NBS_POS_F64_XYZ; args.positions_space = NBS_MEM_HOST; args.positions = pos;
    args.forces_format = NBS_FORCE_F64_XYZ; args.forces_space = NBS_MEM_HOST; args.forces = frc; args.forces_accumulate = 0;
    args.box[0] = args.box[4] = args.box[8] = L;
    args.include_forces = args.include_energy = args.include_direct = args.include_reciprocal = 1;
    args.slice_energies = energies;
    for (int step = 0; step < 10; step++)
        CHECK(nbs_execute(ctx, &args));                 /* every rank: the same positions in, the same reduced forces out */
    printf("rank %d: slice (0,0) Coulomb %.9g, force on atom 0 (%.9g, %.9g, %.9g)\n", rank, energies[0], frc[0], frc[1], frc[2]);
    CHECK(nbs_destroy(ctx));
    return 0;
}

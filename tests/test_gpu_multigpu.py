"""Sharded evaluation on the GPU (run with -m gpu).

* `test_shards_in_lockstep_*`: K shards of one system evaluated by K contexts on ONE device, the two
  exchanges emulated with tensor copies / sums (multigpu.evaluate_lockstep) -- exercises every sharded
  kernel path (interleaved i-blocks, round-robin exceptions, own-subset spreading / FFT / gather, owner-
  of-the-lower-subset slice energies, the split x pass) and must reproduce the unsharded evaluation:
  forces BIT-exactly where only integer sums differ, i.e. the fixed-point accumulators agree exactly
  for direct space; PME goes through float atomics, so forces are compared to 1e-6 relative RMS and
  energies to 1e-9 relative.
* `test_nccl_two_ranks`: the same through torch.distributed + NCCL on two GPUs (skipped on a 1-GPU box).
"""
import importlib
import os
import socket

import numpy as np
import pytest

from helpers import force_rel_rms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multigpu():
    return importlib.import_module("openmm-nonbonded-slicing_b200.multigpu")


@pytest.fixture(scope="module")
def systems():
    return importlib.import_module("openmm-nonbonded-slicing_b200.systems")


def unsharded(nbs, s, lam, direct=True, recip=True):
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
    kernel.initialize(s.system, s.force)
    forces = np.zeros_like(s.positions)
    e = kernel._evaluate(s.positions, s.box, lam, np.zeros(0), direct, recip, forces)
    count, h, _ = kernel.getPairSet(with_pairs=False) if direct else (0, 0, None)
    return forces, e, count, h


def run_lockstep(nbs, multigpu, s, lam, plan, direct=True, recip=True):
    import torch
    dev = torch.device("cuda:0")
    pos = torch.tensor(s.positions, dtype=torch.float64, device=dev)
    shards, outs = [], []
    for r in range(plan.world_size):
        k = multigpu.ShardedB200Kernel(nbs.Platform())
        k.initialize(s.system, s.force)
        k.set_plan(plan, r)
        out = torch.zeros_like(pos)
        k.prepare(pos.data_ptr(), s.box, out.data_ptr(), lam, stream=torch.cuda.current_stream().cuda_stream,
                  includeDirect=direct, includeReciprocal=recip)
        shards.append(k)
        outs.append(out)
    energies = multigpu.evaluate_lockstep(plan, shards)
    torch.cuda.synchronize()
    pairs = [k.getPairSet(with_pairs=False) for k in shards] if direct else []
    return [o.cpu().numpy() for o in outs], energies, pairs


@pytest.mark.parametrize("name,world,share", [("C1", 2, None), ("C2", 3, [1, 2, 2]), ("C3", 2, [1, 3]), ("C3", 4, [0, 1, 2, 2])])
def test_shards_in_lockstep_match_unsharded(nbs, multigpu, systems, name, world, share):
    s = systems.make_system(name)
    ns = s.force.getNumSubsets()
    lam = np.random.default_rng(5).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    for direct, recip in ((True, False), (False, True), (True, True)):
        ref_f, ref_e, ref_count, ref_hash = unsharded(nbs, s, lam, direct, recip)
        plan = multigpu.ShardPlan(world, ns, share)
        forces, energies, pairs = run_lockstep(nbs, multigpu, s, lam, plan, direct, recip)
        for f, e in zip(forces, energies):
            assert np.array_equal(f, forces[0])                 # every rank holds the same reduced forces
            if recip:
                assert force_rel_rms(f, ref_f) <= 1e-6
            else:
                assert np.array_equal(f, ref_f)                 # integer accumulation: bit-exact
            assert np.allclose(e, ref_e, rtol=1e-9, atol=1e-9)
        if direct:
            # the ranks' pair sets are disjoint and their union is the unsharded set
            assert sum(p[0] for p in pairs) == ref_count
            assert sum(p[1] for p in pairs) % 2**64 == ref_hash


def test_capacity_retry_is_collective(nbs, multigpu, systems):
    """A dense system overflows the initial list capacity: every shard must retry together."""
    s = systems.make_system("C2")
    lam = np.ones((s.force.getNumSlices(), 2))
    import ctypes as C
    plan = multigpu.ShardPlan(2, s.force.getNumSubsets())
    ref_f, ref_e, _, _ = unsharded(nbs, s, lam)
    import torch
    dev = torch.device("cuda:0")
    pos = torch.tensor(s.positions, dtype=torch.float64, device=dev)
    shards, outs = [], []
    for r in range(2):
        k = multigpu.ShardedB200Kernel(nbs.Platform())
        k.initialize(s.system, s.force)
        k.set_plan(plan, r)
        if r == 1:
            nbs.abi.check(k.lib.nbs_debug_set_list_capacity(k.handle, 256, 64))    # far too small: forces a retry
        out = torch.zeros_like(pos)
        k.prepare(pos.data_ptr(), s.box, out.data_ptr(), lam, stream=torch.cuda.current_stream().cuda_stream)
        shards.append(k)
        outs.append(out)
    energies = multigpu.evaluate_lockstep(plan, shards)
    torch.cuda.synchronize()
    for o, e in zip(outs, energies):
        assert force_rel_rms(o.cpu().numpy(), ref_f) <= 1e-6
        assert np.allclose(e, ref_e, rtol=1e-9, atol=1e-9)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, name, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
        multigpu = importlib.import_module("openmm-nonbonded-slicing_b200.multigpu")
        systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
        s = systems.make_system(name)
        lam = np.random.default_rng(5).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
        plan = multigpu.ShardPlan(world, s.force.getNumSubsets())
        group = dist.new_group(plan.pme_ranks()) if plan.num_pme_ranks > 1 else None
        k = multigpu.ShardedB200Kernel(nbs.Platform(deviceIndex=rank))
        k.initialize(s.system, s.force)
        k.set_plan(plan, rank)
        pos = torch.tensor(s.positions, dtype=torch.float64, device=dev)
        frc = torch.zeros_like(pos)
        k.prepare(pos.data_ptr(), s.box, frc.data_ptr(), lam, stream=torch.cuda.current_stream().cuda_stream)
        e = multigpu.evaluate_distributed(plan, rank, k, dist, group)
        torch.cuda.synchronize()
        out.put((rank, frc.cpu().numpy(), e.copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_nccl_two_ranks(nbs, multigpu, systems):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, "C3", out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=600) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    s = systems.make_system("C3")
    lam = np.random.default_rng(5).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    ref_f, ref_e, _, _ = unsharded(nbs, s, lam)
    assert np.array_equal(results[0][1], results[1][1])
    for _, f, e in results:
        assert force_rel_rms(f, ref_f) <= 1e-6
        assert np.allclose(e, ref_e, rtol=1e-9, atol=1e-9)


# ====================================================================================================
# Peer-memory sharding (nbs_set_slab_shard): x-slabs of every subset grid on every rank, the fused x pass and the force
# reduction over peer memory
# ====================================================================================================
def run_peer_lockstep(nbs, multigpu, s, lam, world, direct=True, recip=True, flags=0, want_energies=True, share=None, gv=None, small_lists_on=None):
    """`world` contexts on ONE device reaching each other's buffers by pointer; the library's five steps in lock step
    on one stream (stream order stands in for the barriers)."""
    import torch
    dev = torch.device("cuda:0")
    pos = torch.tensor(s.positions, dtype=torch.float64, device=dev)
    grid = s.force.getPMEParameters()[1:]
    plan = multigpu.SlabPlan(world, grid, share)
    shards, outs = [], []
    for r in range(world):
        k = multigpu.ShardedB200Kernel(nbs.Platform(flags=flags))
        k.initialize(s.system, s.force)
        k.set_slab_plan(plan, r)
        if small_lists_on == r:
            nbs.abi.check(k.lib.nbs_debug_set_list_capacity(k.handle, 256, 64))
        shards.append(k)
    exports = [k.export_peer() for k in shards]
    for k in shards:
        k.import_peers(exports, in_kernel_barrier=False)
        out = torch.zeros_like(pos)
        k.prepare(pos.data_ptr(), s.box, out.data_ptr(), lam, stream=torch.cuda.current_stream().cuda_stream,
                  includeDirect=direct, includeReciprocal=recip, want_energies=want_energies)
        if gv is not None:
            k._push_parameters(np.asarray(lam, dtype=np.float64), gv)
        outs.append(out)
    energies = multigpu.evaluate_peer_lockstep(shards)
    torch.cuda.synchronize()
    pairs = [k.getPairSet(with_pairs=False) for k in shards] if direct else []
    return [o.cpu().numpy() for o in outs], energies, pairs


@pytest.mark.parametrize("name,world,share", [("C1", 2, None), ("C2", 3, [1, 2, 2]), ("C3", 4, None), ("C3", 8, None), ("T1_pme", 2, None)])
def test_peer_shards_in_lockstep_match_unsharded(nbs, multigpu, systems, name, world, share):
    variant = name in systems.VARIANTS
    s = systems.make_variant(name) if variant else systems.make_system(name)
    lam = np.random.default_rng(5).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    gv = np.full(max(s.force.getNumGlobalParameters(), 1), 0.45) if variant else None
    for direct, recip in ((True, False), (False, True), (True, True)):
        kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
        kernel.initialize(s.system, s.force)
        ref_f = np.zeros_like(s.positions)
        ref_e = kernel._evaluate(s.positions, s.box, lam, gv if gv is not None else np.zeros(0), direct, recip, ref_f)
        ref_count, ref_hash, _ = kernel.getPairSet(with_pairs=False) if direct else (0, 0, None)
        forces, energies, pairs = run_peer_lockstep(nbs, multigpu, s, lam, world, direct, recip, share=share, gv=gv)
        for f, e in zip(forces, energies):
            assert np.array_equal(f, forces[0])                 # every rank holds the same reduced forces
            if recip:
                assert force_rel_rms(f, ref_f) <= 1e-6
            else:
                assert np.array_equal(f, ref_f)                 # integer accumulation: bit-exact
            assert np.allclose(e, ref_e, rtol=1e-9, atol=1e-9)
            assert np.array_equal(e, energies[0])               # rank-ordered sums: identical everywhere
        if direct:
            assert sum(p[0] for p in pairs) == ref_count
            assert sum(p[1] for p in pairs) % 2**64 == ref_hash


def test_peer_shards_forces_only_and_deterministic(nbs, multigpu, systems):
    """Single-precision grids (no energies requested), and NBS_FLAG_DETERMINISTIC: fixed-point spreading into the own
    planes makes the sharded reciprocal forces bit-reproducible too."""
    s = systems.make_system("C2")
    lam = np.random.default_rng(6).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    kernel = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
    kernel.initialize(s.system, s.force)
    ref_f = np.zeros_like(s.positions)
    kernel._evaluate(s.positions, s.box, lam, np.zeros(0), True, True, ref_f)
    forces, _, _ = run_peer_lockstep(nbs, multigpu, s, lam, 3, want_energies=False)
    assert force_rel_rms(forces[0], ref_f) <= 2e-6 and np.array_equal(forces[0], forces[2])
    runs = [run_peer_lockstep(nbs, multigpu, s, lam, 3, flags=nbs.abi.NBS_FLAG_DETERMINISTIC)[0][0] for _ in range(3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    assert force_rel_rms(runs[0], ref_f) <= 1e-6


def test_peer_capacity_retry_is_collective(nbs, multigpu, systems):
    s = systems.make_system("C2")
    lam = np.ones((s.force.getNumSlices(), 2))
    ref_f, ref_e, _, _ = unsharded(nbs, s, lam)
    forces, energies, _ = run_peer_lockstep(nbs, multigpu, s, lam, 2, small_lists_on=1)
    for f, e in zip(forces, energies):
        assert force_rel_rms(f, ref_f) <= 1e-6
        assert np.allclose(e, ref_e, rtol=1e-9, atol=1e-9)


def _peer_worker(rank, world, port, name, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    device = rank % torch.cuda.device_count()
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    # gloo: the only collective is the host-side exchange of the buffer exports (two ranks may share one GPU here)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
        multigpu = importlib.import_module("openmm-nonbonded-slicing_b200.multigpu")
        systems = importlib.import_module("openmm-nonbonded-slicing_b200.systems")
        s = systems.make_system(name)
        lam = np.random.default_rng(5).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
        plan = multigpu.SlabPlan(world, s.force.getPMEParameters()[1:])
        k = multigpu.ShardedB200Kernel(nbs.Platform(deviceIndex=device))
        k.initialize(s.system, s.force)
        multigpu.connect_peers(k, plan, rank, dist)
        pos = torch.tensor(s.positions, dtype=torch.float64, device=dev)
        frc = torch.zeros_like(pos)
        results = []
        for it in range(3):
            k.prepare(pos.data_ptr(), s.box, frc.data_ptr(), lam, stream=torch.cuda.current_stream().cuda_stream)
            e = k.evaluate_peer()
            torch.cuda.synchronize()
            results.append((frc.cpu().numpy().copy(), e.copy()))
        out.put((rank, results))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_peer_two_processes(nbs, multigpu, systems):
    """Two processes, CUDA IPC mappings, barriers over flags in peer memory, nbs_execute driving the whole sharded
    evaluation -- on two GPUs when the box has them, else both ranks on GPU 0 (time-sliced: slow, but the same code)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, "C2", out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=600) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    s = systems.make_system("C2")
    lam = np.random.default_rng(5).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    ref_f, ref_e, _, _ = unsharded(nbs, s, lam)
    for it in range(3):
        assert np.array_equal(results[0][1][it][0], results[1][1][it][0])
        assert np.array_equal(results[0][1][it][1], results[1][1][it][1])
        for _, runs in results:
            assert force_rel_rms(runs[it][0], ref_f) <= 1e-6
            assert np.allclose(runs[it][1], ref_e, rtol=1e-9, atol=1e-9)


def test_peer_shards_sorted_pme_column_ranges(nbs, multigpu, systems):
    """Large systems run PME from the cell-sorted records, and a rank then only looks at the atoms of the cell columns
    that can reach its planes (k_pme.cu inSlabRanges).  Forced here on C3 (NBS_FLAG_SORTED_PME), 8 ranks, and repeated
    after every atom has moved by up to 0.01 nm, twice -- the neighbour list and the sort order are re-used, atoms cross slab
    and box boundaries -- against the unsharded evaluation of the same positions."""
    import torch
    s = systems.make_system("C3")
    n = s.force.getNumParticles()
    lam = np.random.default_rng(8).uniform(0.2, 1.0, size=(s.force.getNumSlices(), 2))
    world = 8
    dev = torch.device("cuda:0")
    plan = multigpu.SlabPlan(world, s.force.getPMEParameters()[1:])
    shards = []
    for r in range(world):
        k = multigpu.ShardedB200Kernel(nbs.Platform(flags=nbs.abi.NBS_FLAG_SORTED_PME))
        k.initialize(s.system, s.force)
        k.set_slab_plan(plan, r)
        shards.append(k)
    exports = [k.export_peer() for k in shards]
    for k in shards:
        k.import_peers(exports, in_kernel_barrier=False)
    plain = nbs.B200CalcSlicedNonbondedForceKernel(nbs.Platform())
    plain.initialize(s.system, s.force)
    # (device buffers are allocated by a context's first evaluation, and ANY allocation in the process invalidates kept
    # lists -- captured graphs hold pointers: let the comparison context allocate before the shards build theirs)
    plain._evaluate(s.positions, s.box, lam, np.zeros(0), True, True, np.zeros((n, 3)))
    rng = np.random.default_rng(9)
    shift = rng.uniform(-0.01, 0.01, size=(n, 3))
    for step in range(3):
        positions = s.positions + step*0.5*shift
        pos = torch.tensor(positions, dtype=torch.float64, device=dev)
        outs = []
        for k in shards:
            out = torch.zeros_like(pos)
            k.prepare(pos.data_ptr(), s.box, out.data_ptr(), lam, stream=torch.cuda.current_stream().cuda_stream)
            outs.append(out)
        energies = multigpu.evaluate_peer_lockstep(shards)
        torch.cuda.synchronize()
        ref_f = np.zeros((n, 3))
        ref_e = plain._evaluate(positions, s.box, lam, np.zeros(0), True, True, ref_f)
        for o, e in zip(outs, energies):
            assert force_rel_rms(o.cpu().numpy(), ref_f) <= 1e-6, step
            # (Coulomb: double-precision sums; Lennard-Jones: fp32 sums per tile, and the tiles of the two contexts'
            # lists need not be the same once lists are re-used -- the tolerance of test_list_reuse_matches_rebuild)
            assert np.allclose(e[:, 0], ref_e[:, 0], rtol=2e-8, atol=1e-6) and np.allclose(e[:, 1], ref_e[:, 1], rtol=2e-6, atol=1e-6), step
        if step > 0:
            assert all(k.getListStats()["reused_last"] for k in shards)

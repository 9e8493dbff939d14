"""Multi-GPU evaluation of the SlicedNonbondedForce hot path: one process per GPU.

The reference has no collective anywhere; its only multi-device path is OpenMM's per-device split of
the direct-space tiles with reciprocal space pinned to device 0
(platforms/cuda/src/CudaParallelNonbondedSlicingKernels.cpp:35-53;
platforms/common/src/CommonNonbondedSlicingKernels.cpp:416, 465, 643-646, 725-728).  Here the path is
sharded where it shards naturally:

  * positions are replicated; every rank sorts them identically (the sort is deterministic), so
    "sorted atom k" means the same atom everywhere;
  * direct space: i-blocks are dealt to the ranks in an interleaved pattern (``ShardPlan.block_share``),
    exceptions round-robin; ranks that own PME grids get a smaller share;
  * PME: subset grids are assigned to ranks as contiguous ranges (``ShardPlan.subset_range``); spreading,
    FFTs and gather of a grid happen on its owner.  The sliced convolution needs every subset's
    spectrum at the same k, so the owners exchange half spectra once (NCCL broadcast inside the group
    of grid owners);
  * forces (64-bit fixed point -- integer sums are exact and order independent, so every rank ends up
    with bit-identical forces) and the slice-energy table are combined with NCCL all-reduce.

A second scheme, PEER-MEMORY SHARDING (``SlabPlan``, ``connect_peers``, ``evaluate_peer``), removes what limits the
one above (a subset's reciprocal work sits on one rank; 25 MB of force accumulators go through an all-reduce):

  * PME is split by x-slabs of every subset grid over ALL ranks; the fused x pass reads and writes the other ranks'
    planes over NVLink peer memory (the transposes of a slab-decomposed FFT, inside the kernel that consumes them);
  * one kernel reduces the fixed-point force accumulators over peer memory (reduce-scatter + all-gather in place);
  * the steps are ordered by barriers over flags in peer memory, so a sharded evaluation is ONE library call
    (``nbs_execute``) with no collective on the data path.  NCCL is only used where a real exchange of host-side
    data remains: the all-gather of position shards in the end-to-end path.

``torch.distributed`` is plumbing: the compute is the C ABI's three phases (include/nbslice_b200.h:
nbs_execute_begin / _convolve / _finish), and the collectives are issued on the same CUDA stream in
between.  The choreography is written against a small backend protocol so that the CPU tests can run
it under ``gloo`` with a NumPy stand-in for the kernels (tests/test_multigpu_cpu.py).
"""
import ctypes as C
import json
import os
import time

import numpy as np

from . import abi
from .api import B200CalcSlicedNonbondedForceKernel

PATTERN_PERIOD = 64     # i-blocks are dealt in periods of this many consecutive blocks


class ShardPlan:
    """Which rank does what.  Pure integer logic, identical on every rank."""

    def __init__(self, world_size, num_subsets, direct_share=None, period=PATTERN_PERIOD):
        if world_size < 1 or num_subsets < 1:
            raise ValueError("world_size and num_subsets must be positive")
        self.world_size = world_size
        self.num_subsets = num_subsets
        self.num_pme_ranks = min(world_size, num_subsets)
        self.period = max(period, world_size)
        if direct_share is None:
            direct_share = [1.0]*world_size
        if len(direct_share) != world_size or min(direct_share) < 0 or sum(direct_share) <= 0:
            raise ValueError("direct_share needs one non-negative entry per rank and a positive sum")
        self.widths = self._apportion(direct_share, self.period)

    @staticmethod
    def _apportion(share, total):
        """Largest-remainder apportionment of `total` slots proportional to `share`."""
        s = float(sum(share))
        exact = [total*x/s for x in share]
        widths = [int(np.floor(e)) for e in exact]
        order = sorted(range(len(share)), key=lambda r: (-(exact[r]-widths[r]), r))
        for r in order[:total-sum(widths)]:
            widths[r] += 1
        return widths

    def subset_range(self, rank):
        """Contiguous range [lo, hi) of subsets whose PME grids `rank` owns (empty for rank >= num_pme_ranks)."""
        if rank >= self.num_pme_ranks:
            return (0, 0)
        n, p = self.num_subsets, self.num_pme_ranks
        return (rank*n//p, (rank+1)*n//p)

    def subset_owner(self, subset):
        for r in range(self.num_pme_ranks):
            lo, hi = self.subset_range(r)
            if lo <= subset < hi:
                return r
        raise ValueError("subset out of range")

    def pme_ranks(self):
        return list(range(self.num_pme_ranks))

    def block_share(self, rank):
        """(period, offset, width): i-block b is this rank's iff offset <= b % period < offset + width."""
        return (self.period, sum(self.widths[:rank]), self.widths[rank])

    def block_owner(self, block):
        k = block % self.period
        for r in range(self.world_size):
            _, off, w = self.block_share(r)
            if off <= k < off+w:
                return r
        raise AssertionError

    @classmethod
    def balanced(cls, world_size, num_subsets, direct_ms, pme_ms_per_rank, period=PATTERN_PERIOD):
        """Shares that equalise  share_r * direct_ms + pme_ms_per_rank[r]  (all times of ONE rank doing
        that work alone): share_r = (T - pme_r)/direct_ms with T chosen so the shares sum to 1."""
        pme = list(pme_ms_per_rank) + [0.0]*(world_size-len(pme_ms_per_rank))
        active = list(range(world_size))
        share = [0.0]*world_size
        while active:
            T = (direct_ms + sum(pme[r] for r in active))/len(active)
            negative = [r for r in active if T - pme[r] < 0]
            if not negative:
                for r in active:
                    share[r] = (T - pme[r])/direct_ms
                break
            for r in negative:          # this rank is busy with PME alone for longer than the others need
                active.remove(r)
        if sum(share) <= 0:
            share = [1.0]*world_size
        return cls(world_size, num_subsets, share, period)

    def describe(self):
        return {"world_size": self.world_size, "pme_ranks": self.num_pme_ranks,
                "subset_ranges": [self.subset_range(r) for r in range(self.world_size)],
                "block_pattern": {"period": self.period, "widths": self.widths}}


class SlabPlan:
    """Peer-memory sharding: which rank owns which grid planes, x-pass rows, i-blocks and accumulator words.
    Pure integer logic, identical on every rank and mirrored by the library (nbs_set_slab_shard, k_fft_x_conv2's
    owner formula, launchPeerReduce)."""

    def __init__(self, world_size, grid, direct_share=None, period=PATTERN_PERIOD):
        if world_size < 1 or world_size > abi.NBS_MAX_RANKS:
            raise ValueError("world_size must lie in 1 .. NBS_MAX_RANKS")
        if min(grid[0], grid[1]) < world_size:
            raise ValueError("more ranks than grid planes")
        self.world_size = world_size
        self.grid = tuple(int(g) for g in grid)
        self.period = max(period, world_size)
        if direct_share is None:
            direct_share = [1.0]*world_size
        if len(direct_share) != world_size or min(direct_share) < 0 or sum(direct_share) <= 0:
            raise ValueError("direct_share needs one non-negative entry per rank and a positive sum")
        self.widths = ShardPlan._apportion(direct_share, self.period)

    def x_range(self, rank):
        """Grid planes [lo, hi) whose spreading, z/y transforms and gather `rank` does."""
        n, r = self.grid[0], self.world_size
        return (rank*n//r, (rank+1)*n//r)

    def y_range(self, rank):
        """Rows [lo, hi) of the fused x pass that `rank` runs (reading every rank's planes)."""
        n, r = self.grid[1], self.world_size
        return (rank*n//r, (rank+1)*n//r)

    def x_owner(self, x):
        """The rank whose slab holds plane x -- the closed form the x kernel uses."""
        return ((x + 1)*self.world_size - 1)//self.grid[0]

    def block_share(self, rank):
        return (self.period, sum(self.widths[:rank]), self.widths[rank])

    @staticmethod
    def word_range(rank, world_size, words):
        """16-byte words [lo, hi) of the force accumulators that `rank` reduces (`words` of 16 bytes in total)."""
        return (words*rank//world_size, words*(rank+1)//world_size)

    def describe(self):
        return {"world_size": self.world_size, "scheme": "peer memory: x-slabs of every subset grid, fused x pass and force reduction over NVLink",
                "x_ranges": [self.x_range(r) for r in range(self.world_size)],
                "block_pattern": {"period": self.period, "widths": self.widths}}


def shard_rows(n, rank, world_size):
    """Rows [lo, hi) of an n-row host array (positions in, forces out) that `rank` moves over PCIe; the shards are
    padded to equal length `rows` for the all-gather."""
    rows = (n + world_size - 1)//world_size
    return min(n, rank*rows), min(n, (rank+1)*rows), rows


# ---------------------------------------------------------------------------------------------------
# Choreography (backend-agnostic).  A backend provides, for ONE rank:
#   begin()                    -> None
#   spectrum_slabs()           -> list of per-subset tensors (views of the backend's spectra buffer)
#   convolve()                 -> None
#   reduce_tensors()           -> list of tensors to all-reduce (sum) in place
#   finish()                   -> result, or RETRY
# ---------------------------------------------------------------------------------------------------
RETRY = object()


def exchange_spectra(plan, rank, slabs, dist, group):
    """Every grid owner receives every other owner's half spectra (one broadcast per subset)."""
    if plan.num_pme_ranks <= 1 or rank >= plan.num_pme_ranks:
        return
    works = []
    for subset, slab in enumerate(slabs):
        works.append(dist.broadcast(slab, src=plan.subset_owner(subset), group=group, async_op=True))
    for w in works:
        w.wait()


def evaluate_distributed(plan, rank, backend, dist, pme_group, max_attempts=7):
    """One evaluation of this rank's shard, exchanges included.  Collective: every rank must call it."""
    for _ in range(max_attempts):
        backend.begin()
        exchange_spectra(plan, rank, backend.spectrum_slabs(), dist, pme_group)
        backend.convolve()
        if plan.world_size > 1:
            for t in backend.reduce_tensors():
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        result = backend.finish()
        if result is not RETRY:
            return result
    raise RuntimeError("neighbour list capacity exceeded")


def evaluate_peer_lockstep(backends, max_attempts=7):
    """Peer-memory sharding with all ranks inside ONE process (several contexts on one device sharing a stream, or
    NumPy stand-ins): the library's five steps, every rank finishing step k before any rank starts step k+1 --
    the ordering the in-kernel barriers provide between processes."""
    for _ in range(max_attempts):
        results = None
        for step in range(abi.NBS_NUM_STEPS):
            results = [b.step(step) for b in backends]
        if not any(r is RETRY for r in results):
            return results
        assert all(r is RETRY for r in results), "the overflow flag must reach every rank"
    raise RuntimeError("neighbour list capacity exceeded")


def connect_peers(kernel, plan, rank, dist, group=None, in_kernel_barrier=True):
    """Set the shard, exchange the ranks' buffer exports (host-side plumbing) and map the peers' memory."""
    kernel.set_slab_plan(plan, rank)
    mine = kernel.export_peer()
    if plan.world_size == 1:
        exports = [mine]
    else:
        exports = [None]*plan.world_size
        dist.all_gather_object(exports, mine, group=group)
    kernel.import_peers(exports, in_kernel_barrier)


def evaluate_lockstep(plan, backends, max_attempts=7):
    """The same choreography for all ranks inside ONE process (several shards on one device, or NumPy
    stand-ins): collectives are emulated with tensor copies and sums.  Used by the tests to exercise
    the sharded kernels without a second GPU."""
    import torch
    for _ in range(max_attempts):
        for b in backends:
            b.begin()
        slabs = [b.spectrum_slabs() for b in backends]
        for subset in range(plan.num_subsets):
            owner = plan.subset_owner(subset)
            for r in plan.pme_ranks():
                if r != owner and slabs[r]:
                    slabs[r][subset].copy_(slabs[owner][subset])
        for b in backends:
            b.convolve()
        tensors = [b.reduce_tensors() for b in backends]
        for k in range(len(tensors[0])):
            total = tensors[0][k].clone()
            for r in range(1, len(backends)):
                total += tensors[r][k]
            for r in range(len(backends)):
                tensors[r][k].copy_(total)
        if hasattr(torch, "cuda") and torch.cuda.is_available():
            torch.cuda.synchronize()
        results = [b.finish() for b in backends]
        if not any(r is RETRY for r in results):
            return results
        assert all(r is RETRY for r in results), "the overflow flag must reach every rank"
    raise RuntimeError("neighbour list capacity exceeded")


# ---------------------------------------------------------------------------------------------------
# The CUDA backend: one shard of a B200 kernel
# ---------------------------------------------------------------------------------------------------
class _DeviceMemory:
    """Raw device memory exposed through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _alias(ptr, shape, typestr, device):
    import torch
    return torch.as_tensor(_DeviceMemory(ptr, shape, typestr), device=device)


class ShardedB200Kernel(B200CalcSlicedNonbondedForceKernel):
    """A B200 kernel that evaluates one rank's shard.  `initialize` as usual, then `set_plan`."""

    def set_plan(self, plan, rank):
        self.plan, self.rank = plan, rank
        period, offset, width = plan.block_share(rank)
        lo, hi = plan.subset_range(rank)
        abi.check(self.lib.nbs_set_shard(self.handle, rank, plan.world_size, period, offset, width, lo, hi))

    # -- peer-memory sharding --------------------------------------------------------------------------
    def set_slab_plan(self, plan, rank):
        self.plan, self.rank = plan, rank
        period, offset, width = plan.block_share(rank)
        abi.check(self.lib.nbs_set_slab_shard(self.handle, rank, plan.world_size, period, offset, width))

    def export_peer(self):
        ex = abi.PeerExport()
        ex.struct_size = C.sizeof(abi.PeerExport)
        abi.check(self.lib.nbs_export_peer(self.handle, C.byref(ex)))
        return bytes(ex)

    def import_peers(self, exports, in_kernel_barrier=True):
        array = (abi.PeerExport*len(exports))()
        for k, blob in enumerate(exports):
            C.memmove(C.byref(array[k]), blob, C.sizeof(abi.PeerExport))
        abi.check(self.lib.nbs_import_peers(self.handle, len(exports), array, int(in_kernel_barrier)))

    def step(self, k):
        """One of the library's five steps (the caller orders the ranks); the last one returns the energies."""
        status = self.lib.nbs_execute_step(self.handle, C.byref(self._args), k)
        if status == abi.NBS_RETRY:
            return RETRY
        abi.check(status)
        return self._energies if k == abi.NBS_NUM_STEPS-1 else None

    def evaluate_peer(self):
        """The whole sharded evaluation: one library call, barriers over peer memory inside."""
        abi.check(self.lib.nbs_execute(self.handle, C.byref(self._args)))
        return self._energies

    # -- backend protocol ---------------------------------------------------------------------------
    def prepare(self, positions_ptr, box, forces_ptr, lambdas, stream=0, want_energies=True,
                includeDirect=True, includeReciprocal=True):
        import torch
        self._push_parameters(np.asarray(lambdas, dtype=np.float64), np.zeros(0))
        args = abi.ExecArgs()
        args.struct_size = C.sizeof(abi.ExecArgs)
        args.positions_format = abi.NBS_POS_F64_XYZ
        args.positions_space = abi.NBS_MEM_DEVICE
        args.forces_format = abi.NBS_FORCE_F64_XYZ
        args.forces_space = abi.NBS_MEM_DEVICE
        args.forces_accumulate = 0
        args.positions = positions_ptr
        args.forces = forces_ptr
        args.box[:] = list(np.asarray(box, dtype=np.float64).reshape(9))
        args.include_forces = 1
        args.include_energy = 1
        args.include_direct = int(includeDirect)
        args.include_reciprocal = int(includeReciprocal)
        self._energies = np.zeros((self.numSlices, 2)) if want_energies else None
        args.slice_energies = self._energies.ctypes.data_as(C.POINTER(C.c_double)) if want_energies else None
        args.stream = stream
        self._args = args
        self._device = torch.device("cuda", self.platform.deviceIndex)

    def begin(self):
        abi.check(self.lib.nbs_execute_begin(self.handle, C.byref(self._args)))
        ex = abi.ExchangeBuffers()
        ex.struct_size = C.sizeof(abi.ExchangeBuffers)
        abi.check(self.lib.nbs_get_exchange_buffers(self.handle, C.byref(ex)))
        self._ex = ex

    def spectrum_slabs(self):
        if self.plan.num_pme_ranks <= 1 or self.rank >= self.plan.num_pme_ranks or not self._args.include_reciprocal:
            return []
        ex = self._ex
        words = ex.spectrum_bytes_per_subset//(8 if ex.spectrum_is_double else 4)
        typestr = "<f8" if ex.spectrum_is_double else "<f4"
        whole = _alias(ex.spectra, (self.numSubsets, words), typestr, self._device)
        return [whole[s] for s in range(self.numSubsets)]

    def convolve(self):
        abi.check(self.lib.nbs_execute_convolve(self.handle, C.byref(self._args)))

    def reduce_tensors(self):
        ex = self._ex
        return [_alias(ex.forces, (ex.force_words,), "<i8", self._device),
                _alias(ex.energies, (ex.energy_words,), "<f8", self._device)]

    def finish(self):
        status = self.lib.nbs_execute_finish(self.handle, C.byref(self._args))
        if status == abi.NBS_RETRY:
            return RETRY
        abi.check(status)
        return self._energies


def calibrate_plan(kernel_factory, world_size, num_subsets, run_once):
    """Measure (on this rank, alone) the direct-space time and the PME time of each grid owner's share,
    and derive a balanced plan.  `kernel_factory(flags)` makes an initialised ShardedB200Kernel and
    `run_once(kernel)` evaluates it; returns the plan and the measurements (identical inputs on every
    rank give near-identical plans, but callers broadcast rank 0's to be exact)."""
    base = ShardPlan(world_size, num_subsets)
    k = kernel_factory(abi.NBS_FLAG_PROFILE)
    direct_ms, pme_ms = 0.0, []
    for r in range(base.num_pme_ranks):
        # a rank that does all of direct space and subset range r
        lo, hi = base.subset_range(r)
        abi.check(k.lib.nbs_set_shard(k.handle, 0, 1, 1, 0, 1, lo, hi))
        k.plan, k.rank = ShardPlan(1, num_subsets), 0
        for _ in range(3):
            run_once(k)
        times = dict()
        for name, ms in k.getKernelTimes():
            times[name] = times.get(name, 0.0) + ms
        direct_ms = sum(times.get(n, 0.0) for n in ("build_lists", "pair", "bonded"))
        pme_ms.append(sum(times.get(n, 0.0) for n in ("spread", "fft_fwd", "fft_conv_inv", "gather")))
    del k
    plan = ShardPlan.balanced(world_size, num_subsets, direct_ms, pme_ms)
    return plan, {"direct_ms": direct_ms, "pme_ms_per_owner": pme_ms}


# ---------------------------------------------------------------------------------------------------
# bench.py under torch.distributed.run (any N, including 1): strong scaling of the STMV-size system
# ---------------------------------------------------------------------------------------------------
def c5_fixture_errors(golden, forces, energies, tag="full"):
    """Parity figures of a C5 evaluation against tests/golden/C5_reference.npz (the reference's own full-size CPU
    evaluation, oracle/make_golden_c5.py): relative RMS force error over the fixture's 4,096-atom sample, relative
    error of sum |F|^2 over all atoms, worst slice-energy error over max(|E|, 1)."""
    idx = golden["sample"]
    ref = golden[f"{tag}_forces_sample"]
    f_err = float(np.sqrt(((forces[idx]-ref)**2).sum()/(ref**2).sum()))
    sumsq_err = float(abs((forces**2).sum()/golden[f"{tag}_force_sumsq"][0] - 1.0))
    ref_e = golden[f"{tag}_energies"]
    e_err = float(np.max(np.abs(energies-ref_e)/np.maximum(np.abs(ref_e), 1.0)))
    return f_err, sumsq_err, e_err


def bench_main(args, workload_name):
    import importlib
    import torch
    import torch.distributed as dist
    bench = importlib.import_module("bench")
    systems = importlib.import_module(__package__ + ".systems")
    from .api import Platform

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    s = systems.make_system(workload_name)
    n, nsl, ns = s.force.getNumParticles(), s.force.getNumSlices(), s.force.getNumSubsets()
    grid = s.force.getPMEParameters()[1:]
    lam = np.ones((nsl, 2))
    moving = bench.MovingSystem(s.positions, dev)
    frc_dev = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
    warmup = max(args.warmup, bench.MIN_WARMUP)      # both CUDA-graph-free here, but the same protocol as the 1-GPU line
    e2e_warm = max(args.warmup, 3)

    def factory(flags):
        k = ShardedB200Kernel(Platform(deviceIndex=local, flags=flags))
        k.initialize(s.system, s.force)
        return k

    # ---- the sharding scheme: peer memory (CUDA IPC over NVLink) on every rank, or -- if any rank cannot map its
    # peers -- the NCCL scheme of round 1 (subset grids on their owners, all-reduce of the accumulators) -----------------
    plan = SlabPlan(world, grid)
    kernel = factory(0)
    ok, why = 1, ""
    try:
        if os.environ.get("NBS_BENCH_SCHEME", "peer") != "peer":      # exercise the NCCL scheme on a box that has peer access
            raise RuntimeError("NBS_BENCH_SCHEME asks for the NCCL scheme")
        connect_peers(kernel, plan, rank, dist)
    except Exception as exc:                       # noqa: BLE001 -- reported in the line; the other GPU scheme is used
        ok, why = 0, str(exc)
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    peer_mode = bool(flag.item())
    pme_group = None
    if not peer_mode:
        del kernel
        # the round-1 scheme: rank 0 measures direct-space and PME times and derives the i-block shares (grid owners get
        # less direct space), everybody adopts its plan
        payload = [None]
        if rank == 0:
            def run_alone(k):
                k.prepare(moving.pos.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream)
                return evaluate_distributed(ShardPlan(1, ns), 0, k, dist, None)
            calibrated, _ = calibrate_plan(factory, world, ns, run_alone)
            payload = [calibrated.widths]
        dist.broadcast_object_list(payload, src=0)
        plan = ShardPlan(world, ns, [float(w) for w in payload[0]])
        pme_group = dist.new_group(plan.pme_ranks()) if plan.num_pme_ranks > 1 else None
        kernel = factory(0)
        kernel.set_plan(plan, rank)

    def evaluate(k, positions_ptr, want=True):
        k.prepare(positions_ptr, s.box, frc_dev.data_ptr(), lam, stream=stream, want_energies=want)
        if peer_mode:
            return k.evaluate_peer()
        return evaluate_distributed(plan, rank, k, dist, pme_group)

    def barrier():
        dist.barrier()

    def timed(step, steps, trajectory=moving):
        per_step, result = bench.timed_steps(step, trajectory, flush, warmup, steps, barrier=barrier)
        t = torch.tensor(per_step, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)               # max over ranks, step by step
        return float(t.mean().item()), result

    # ---- device-resident throughput of the moving system ----------------------------------------------------------------
    timed(lambda: evaluate(kernel, moving.pos.data_ptr()), 0)
    launches0, stats0 = kernel.getLaunchCount(), kernel.getListStats()
    sampler = bench.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.__enter__()
    per_step, energies = bench.timed_steps(lambda: evaluate(kernel, moving.pos.data_ptr()), moving, flush, 0, args.steps, barrier=barrier, t0=warmup)
    if sampler:
        sampler.sample_now()
        sampler.__exit__()
    launches = kernel.getLaunchCount()-launches0
    policy = bench.list_policy(stats0, kernel.getListStats())
    t = torch.tensor(per_step, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.mean().item())
    value = 1e3/ms
    forces_ms, _ = timed(lambda: evaluate(kernel, moving.pos.data_ptr(), want=False), args.steps)

    # ---- end to end: every rank uploads ITS rows of the positions (pinned host memory), the shards are all-gathered over
    # NVLink, and every rank downloads ITS rows of the forces; energies go to the host on every rank ---------------------
    lo, hi, rows = shard_rows(n, rank, world)
    pos_shard_host = torch.zeros((rows, 3), dtype=torch.float64).pin_memory()
    frc_shard_host = torch.zeros((rows, 3), dtype=torch.float64).pin_memory()
    pos_shard_dev = torch.zeros((rows, 3), dtype=torch.float64, device=dev)
    pos_all = torch.zeros((world*rows, 3), dtype=torch.float64, device=dev)
    e2e = []
    for it in range(e2e_warm + args.steps):
        pos_shard_host[:hi-lo] = torch.from_numpy(moving.host_positions(it, lo, hi))
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pos_shard_dev.copy_(pos_shard_host, non_blocking=True)
        if world > 1:
            dist.all_gather_into_tensor(pos_all, pos_shard_dev)
        else:
            pos_all.copy_(pos_shard_dev)
        e_host = evaluate(kernel, pos_all.data_ptr())
        frc_shard_host[:hi-lo].copy_(frc_dev[lo:hi], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter()-t0
        if it >= e2e_warm:
            e2e.append(dt)
    te = torch.tensor(e2e, dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = 1.0/float(te.mean().item())

    # ---- parity of THIS job against the reference's own full-size evaluation (committed fixture), un-moved positions -----
    parity = None
    fixture = os.path.join(bench.ROOT, "tests", "golden", f"{workload_name}_reference.npz")
    frozen = bench.MovingSystem(s.positions, dev)
    if workload_name == "C5" and os.path.exists(fixture):
        g = np.load(fixture)
        kernel.prepare(frozen.pos.data_ptr(), s.box, frc_dev.data_ptr(), g["lambdas"], stream=stream)
        e_fix = kernel.evaluate_peer() if peer_mode else evaluate_distributed(plan, rank, kernel, dist, pme_group)
        torch.cuda.synchronize()
        f_err, sumsq_err, e_err = c5_fixture_errors(g, frc_dev.cpu().numpy(), e_fix)
        local_pairs = kernel.getPairSet(with_pairs=False)
        pairs = torch.tensor([local_pairs[0]], dtype=torch.int64, device=dev)
        dist.all_reduce(pairs)
        hashes = [None]*world
        dist.all_gather_object(hashes, int(local_pairs[1]))
        worst = torch.tensor([f_err, sumsq_err, e_err], dtype=torch.float64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        parity = {"fixture": "tests/golden/C5_reference.npz (the reference's own full-size CPU evaluation)",
                  "force_rel_rms_4096_atom_sample": float(worst[0].item()), "sum_F2_rel_err": float(worst[1].item()),
                  "max_energy_err_over_max_absE_1": float(worst[2].item()), "worst_over_ranks": True,
                  "pair_count_matches": int(pairs.item()) == int(g["pair_count"][0]),
                  "pair_hash_matches": sum(hashes) % 2**64 == int(g["pair_hash"][0])}
    local_pairs = kernel.getPairSet(with_pairs=False)[0]
    pairs = torch.tensor([local_pairs], dtype=torch.int64, device=dev)
    dist.all_reduce(pairs)
    fsum = frc_dev.abs().sum().reshape(1)
    fmin, fmax = fsum.clone(), fsum.clone()
    dist.all_reduce(fmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(fmax, op=dist.ReduceOp.MAX)

    # ---- per-kernel durations of THIS rank's shard (profiled context: serial streams, CUDA events per kernel) ------------
    prof = factory(abi.NBS_FLAG_PROFILE)
    if peer_mode:
        connect_peers(prof, plan, rank, dist)
    else:
        prof.set_plan(plan, rank)
    acc = {}
    for it in range(6):
        frozen.advance(0)
        flush.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        evaluate(prof, frozen.pos.data_ptr())
        if it >= 3:
            for name, tms in prof.getKernelTimes():
                acc[name] = acc.get(name, 0.0) + tms/3
    del prof
    shares = [None]*world
    dist.all_gather_object(shares, (acc.get("pair", 0.0), local_pairs, rank))
    slow_ms, slow_pairs, slow_rank = max(shares)
    all_acc = [None]*world
    dist.all_gather_object(all_acc, {k: round(v, 5) for k, v in acc.items()})

    # ---- the same workload, same protocol, unsharded on ONE GPU (rank 0 alone; the others wait) ----------------------------
    single = None
    if world > 1:
        if rank == 0:
            one = B200CalcSlicedNonbondedForceKernel(Platform(deviceIndex=local))
            one.initialize(s.system, s.force)
            tb = bench.timed_steps(lambda: one.execute_device(moving.pos.data_ptr(), s.box, frc_dev.data_ptr(), lam, stream=stream),
                                   moving, flush, warmup, min(args.steps, 10))[0]
            del one
            single = {"ms_per_step": float(np.mean(tb)), "evals_per_s": 1e3/float(np.mean(tb)), "speedup": float(np.mean(tb))/ms,
                      "efficiency": float(np.mean(tb))/ms/world, "note": "unsharded library on rank 0's GPU, same trajectory and timing protocol"}
        dist.barrier()

    if rank == 0:
        line = {
            "metric": "force+energy evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{workload_name}: {s.description}", "atoms": n, "subsets": ns, "pme_grid": grid[0],
                       "cutoff_nm": 1.0, "ns_per_day_2fs": bench.ns_per_day(value),
                       "l2": "256 MiB buffer written between steps, outside the per-step CUDA events",
                       "motion": bench.MOTION,
                       "neighbour_list": "built with a skin on every rank, re-used until an atom has moved half of it",
                       "list_policy": policy, "interacting_pairs": int(pairs.item()),
                       "sharding": plan.describe() if peer_mode else dict(plan.describe(), scheme="NCCL: subset grids on their owners, spectrum broadcast, all-reduce of the accumulators", peer_memory_unavailable=why),
                       "timing": "CUDA events per step on each rank after a barrier, max over ranks per step, mean over steps"},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(world*rows*24),
                    "d2h_bytes_per_step": int(world*(rows*24 + 8*2*nsl)), "ns_per_day_2fs": bench.ns_per_day(e2e_value),
                    "note": f"whole job: every rank uploads {rows} rows of the positions (pinned host memory), the shards are all-gathered over "
                            "NVLink (NCCL), every rank downloads its rows of the forces and the slice energies"},
            "gpu_launches": int(launches),
            "forces_only": {"ms_per_step": forces_ms, "value": 1e3/forces_ms, "unit": "evals/s"},
            "roofline": bench.pair_roofline(slow_pairs, slow_ms, None, device=local,
                                            note=f"rank {slow_rank}'s share of the i-blocks (the longest pair kernel of the job)"),
            "kernel_ms_rank0": {k: round(v, 5) for k, v in acc.items()},
            "kernel_ms_all_ranks": all_acc,
            "collectives_per_step": ({"nccl": 0, "peer_memory_barriers": 4, "peer_memory_kernels": "k_fft_x_conv2 (x pass over all ranks' planes), k_peer_reduce (force accumulators)"}
                                     if peer_mode else {"spectrum_broadcasts": ns if plan.num_pme_ranks > 1 else 0, "all_reduces": 2}),
            "slice_energy_checksum": float(np.abs(energies).sum()),
            "forces_identical_on_all_ranks": bool(fmin.item() == fmax.item()),
        }
        if single is not None:
            line["single_gpu_same_workload"] = single
        if parity is not None:
            line["parity_vs_fixture"] = parity
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()

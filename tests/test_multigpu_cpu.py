"""Host-side logic of the multi-GPU path, on CPU: the shard plan and the exchange choreography
(spectrum broadcast between grid owners, all-reduce of fixed-point forces + slice energies, retry
propagation), driven with a NumPy stand-in for the CUDA kernels -- in process (lock step) and under
torch.distributed with the gloo backend at world_size 2 and 3.

The stand-in keeps the STRUCTURE of the real shard (i-blocks of 32 atoms dealt by the plan, pairs owned
by the block of the lower index, per-subset "spectra" owned by one rank, a slice (I, J) owned by the
owner of min(I, J), 64-bit fixed-point force accumulators) with toy arithmetic, so that any mistake in
who-computes-what or in the exchanges changes the result."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
multigpu = importlib.import_module("openmm-nonbonded-slicing_b200.multigpu")
ShardPlan = multigpu.ShardPlan


def tri(a, b):
    a, b = max(a, b), min(a, b)
    return a*(a+1)//2 + b


class NumpyShard:
    """One rank's share of a toy sliced evaluation (see module docstring)."""

    def __init__(self, plan, rank, n=200, ns=3, nk=6, seed=7, overflow_on_first_attempt=False):
        rng = np.random.default_rng(seed)
        self.plan, self.rank, self.n, self.ns, self.nk = plan, rank, n, ns, nk
        self.pos = rng.random((n, 3))
        self.q = rng.normal(size=n)
        self.subset = rng.integers(0, ns, size=n)
        self.kvec = rng.integers(-3, 4, size=(nk, 3)).astype(np.float64)
        self.lam = 0.25 + 0.75*rng.random(ns*(ns+1)//2)
        self.nsl = ns*(ns+1)//2
        self.force = torch.zeros(3*n, dtype=torch.int64)
        self.energy = torch.zeros(2*self.nsl + 1, dtype=torch.float64)
        self.spectra = torch.zeros((ns, 2*nk), dtype=torch.float64)
        self.overflow_pending = overflow_on_first_attempt
        self.attempts = 0

    @staticmethod
    def fixed(v):
        return np.round(v*4294967296.0).astype(np.int64)

    def begin(self):
        self.attempts += 1
        self.force.zero_(); self.energy.zero_(); self.spectra.zero_()
        f = np.zeros((self.n, 3), dtype=np.int64)
        e = np.zeros(2*self.nsl + 1)
        nblocks = (self.n + 31)//32
        for b in range(nblocks):
            if self.plan.block_owner(b) != self.rank:
                continue
            for i in range(32*b, min(self.n, 32*b+32)):
                d = self.pos[i] - self.pos[i+1:]
                r = np.sqrt((d*d).sum(axis=1))
                qq = self.q[i]*self.q[i+1:]
                sl = np.array([tri(self.subset[i], s) for s in self.subset[i+1:]], dtype=np.int64)
                g = self.fixed((self.lam[sl]*qq/(r**3 + 0.1))[:, None]*d) if len(sl) else np.zeros((0, 3), dtype=np.int64)
                f[i] += g.sum(axis=0)
                f[i+1:] -= g
                np.add.at(e, 2*sl, qq/(r + 0.1))
        if self.overflow_pending:
            e[2*self.nsl] = 1.0
        lo, hi = self.plan.subset_range(self.rank)
        phase = 2*np.pi*self.pos @ self.kvec.T
        for s in range(lo, hi):
            m = self.subset == s
            S = (self.q[m, None]*np.exp(1j*phase[m])).sum(axis=0)
            self.spectra[s] = torch.from_numpy(np.concatenate([S.real, S.imag]))
        self._f, self._e = f, e

    def spectrum_slabs(self):
        if self.plan.num_pme_ranks <= 1 or self.rank >= self.plan.num_pme_ranks:
            return []
        return [self.spectra[s] for s in range(self.ns)]

    def convolve(self):
        lo, hi = self.plan.subset_range(self.rank)
        S = self.spectra.numpy()
        S = S[:, :self.nk] + 1j*S[:, self.nk:]
        phase = 2*np.pi*self.pos @ self.kvec.T
        for sa in range(lo, hi):
            for sb in range(sa, self.ns):
                self._e[2*tri(sa, sb)] += (0.5 if sa == sb else 1.0)*float((S[sa]*np.conj(S[sb])).real.sum())
            mixed = sum(self.lam[tri(sa, sj)]*S[sj] for sj in range(self.ns))
            m = np.where(self.subset == sa)[0]
            w = (np.exp(-1j*phase[m])*mixed[None, :]).imag
            self._f[m] += self.fixed(self.q[m, None]*(w @ self.kvec))
        self.force.copy_(torch.from_numpy(self._f.reshape(-1)))
        self.energy.copy_(torch.from_numpy(self._e))

    def reduce_tensors(self):
        return [self.force, self.energy]

    def finish(self):
        if self.energy[2*self.nsl].item() != 0.0:
            self.overflow_pending = False          # "capacity grown"
            return multigpu.RETRY
        return self.force.numpy().copy(), self.energy.numpy()[:2*self.nsl].copy()


def serial_result(**kw):
    plan = ShardPlan(1, kw.get("ns", 3))
    shard = NumpyShard(plan, 0, **kw)
    return multigpu.evaluate_lockstep(plan, [shard])[0]


# ---- the plan ------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,ns", [(1, 1), (1, 3), (2, 2), (2, 3), (4, 2), (8, 2), (3, 8), (8, 8)])
def test_plan_partitions_blocks_and_subsets(world, ns):
    plan = ShardPlan(world, ns)
    assert sum(plan.widths) == plan.period
    owners = [plan.block_owner(b) for b in range(5*plan.period + 3)]
    for r in range(world):
        period, off, width = plan.block_share(r)
        mine = [b for b in range(len(owners)) if off <= b % period < off+width]
        assert mine == [b for b, o in enumerate(owners) if o == r]
    covered = []
    for r in range(world):
        lo, hi = plan.subset_range(r)
        covered += list(range(lo, hi))
        assert all(plan.subset_owner(s) == r for s in range(lo, hi))
    assert covered == list(range(ns))
    assert plan.num_pme_ranks == min(world, ns)


def test_plan_local_block_mapping_matches_the_device_formula():
    # localToGlobalBlock in csrc/nbs_internal.h: (local/width)*period + offset + local % width
    plan = ShardPlan(3, 2, [1.0, 2.0, 5.0], period=16)
    for r in range(3):
        period, off, width = plan.block_share(r)
        mine = [b for b in range(200) if plan.block_owner(b) == r]
        mapped = [(k//width)*period + off + k % width for k in range(len(mine))]
        assert mapped == mine


def test_balanced_plan_equalises_work():
    # 4 ranks, 2 grid owners whose PME work takes 3 and 1 ms; direct space 20 ms in total
    plan = ShardPlan.balanced(4, 2, 20.0, [3.0, 1.0], period=240)
    share = np.array(plan.widths)/plan.period
    total = share*20.0 + np.array([3.0, 1.0, 0.0, 0.0])
    assert np.allclose(total, total[0], atol=20.0/240)
    # an owner whose PME alone exceeds the balanced time gets no direct space at all
    plan = ShardPlan.balanced(4, 2, 4.0, [10.0, 0.5], period=64)
    assert plan.widths[0] == 0 and sum(plan.widths) == 64
    with pytest.raises(ValueError):
        ShardPlan(2, 2, [1.0])


# ---- choreography, in process ------------------------------------------------------------------------
@pytest.mark.parametrize("world,ns,share", [(2, 3, None), (3, 3, [1, 2, 3]), (4, 2, [0, 1, 2, 2]), (2, 1, None), (5, 4, None)])
def test_lockstep_shards_reproduce_the_serial_result(world, ns, share):
    ref_f, ref_e = serial_result(ns=ns)
    plan = ShardPlan(world, ns, share)
    shards = [NumpyShard(plan, r, ns=ns) for r in range(world)]
    results = multigpu.evaluate_lockstep(plan, shards)
    for f, e in results:
        assert np.array_equal(f, ref_f)                  # integer sums: bit-exact and identical on every rank
        assert np.allclose(e, ref_e, rtol=1e-12, atol=1e-12)


def test_lockstep_retry_reaches_every_rank():
    plan = ShardPlan(3, 3)
    shards = [NumpyShard(plan, r, overflow_on_first_attempt=(r == 1)) for r in range(3)]
    results = multigpu.evaluate_lockstep(plan, shards)
    assert [s.attempts for s in shards] == [2, 2, 2]
    ref_f, ref_e = serial_result()
    assert all(np.array_equal(f, ref_f) for f, _ in results)


# ---- choreography under torch.distributed (gloo) --------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ns, share, overflow_rank, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(world, ns, share)
        group = dist.new_group(plan.pme_ranks()) if plan.num_pme_ranks > 1 else None
        shard = NumpyShard(plan, rank, ns=ns, overflow_on_first_attempt=(rank == overflow_rank))
        f, e = multigpu.evaluate_distributed(plan, rank, shard, dist, group)
        out.put((rank, f, e, shard.attempts))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ns,share,overflow_rank", [(2, 3, None, -1), (2, 2, [1, 3], 1), (3, 2, [1, 1, 2], -1)])
def test_gloo_ranks_reproduce_the_serial_result(world, ns, share, overflow_rank):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ns, share, overflow_rank, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_f, ref_e = serial_result(ns=ns)
    assert sorted(r[0] for r in results) == list(range(world))
    for _, f, e, attempts in results:
        assert np.array_equal(f, ref_f)
        assert np.allclose(e, ref_e, rtol=1e-12, atol=1e-12)
        assert attempts == (2 if overflow_rank >= 0 else 1)

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


# ====================================================================================================
# Peer-memory sharding (SlabPlan): plan logic and a NumPy model of the five steps
# ====================================================================================================
SlabPlan = multigpu.SlabPlan


@pytest.mark.parametrize("world,grid", [(1, (18, 18, 18)), (2, (18, 20, 24)), (3, (64, 64, 64)), (7, (45, 36, 50)), (8, (180, 180, 180)), (16, (16, 21, 16))])
def test_slab_plan_partitions_planes_rows_blocks_and_words(world, grid):
    plan = SlabPlan(world, grid)
    xs, ys = [], []
    for r in range(world):
        lo, hi = plan.x_range(r)
        assert hi > lo
        xs += list(range(lo, hi))
        assert all(plan.x_owner(x) == r for x in range(lo, hi))       # the closed form of k_fft_x_conv2
        ys += list(range(*plan.y_range(r)))
    assert xs == list(range(grid[0])) and ys == list(range(grid[1]))
    assert sum(plan.widths) == plan.period
    offsets = [plan.block_share(r)[1] for r in range(world)]
    assert offsets == [sum(plan.widths[:r]) for r in range(world)]
    words = 3*(1000 + 32)//2
    covered = []
    for r in range(world):
        covered += list(range(*SlabPlan.word_range(r, world, words)))
    assert covered == list(range(words))
    with pytest.raises(ValueError):
        SlabPlan(grid[0] + 1, grid)


def test_shard_rows_cover_the_array():
    for n, world in ((1066628, 8), (23558, 2), (10, 4), (7, 8)):
        rows_seen, width = [], None
        for r in range(world):
            lo, hi, rows = multigpu.shard_rows(n, r, world)
            width = rows
            rows_seen += list(range(lo, hi))
            assert hi - lo <= rows
        assert rows_seen == list(range(n)) and width*world >= n


class SlabWindows:
    """What every rank can reach through peer memory: per rank its planes, its force accumulators and its mailbox.
    In process the shards share ONE instance (true shared memory); under gloo each rank holds a private copy that
    `sync` makes consistent at the points where the library has a barrier."""

    def __init__(self, world, nx, nrows, ns, n, nsl):
        self.world = world
        self.planes = np.zeros((ns, nx, nrows), dtype=np.complex128)        # plane x lives on its owner; one array models all
        self.force = np.zeros((world, 3*n), dtype=np.int64)
        self.energy = np.zeros((world, 2*nsl + 1))


class NumpySlabShard:
    """One rank of a toy sliced PME in the peer scheme: atoms spread (order 2) onto nx planes x nrows rows per subset;
    step 0 spreads the OWN planes, step 1 runs the x pass (DFT over x, slice energies, lambda mixing, inverse) for the OWN
    rows reading and writing every rank's planes, step 2 gathers from the OWN planes and publishes energies, step 3
    reduces the OWN words of all ranks' accumulators and writes them back, step 4 returns the result."""

    def __init__(self, plan, rank, win, n=120, ns=3, nrows=5, seed=11, overflow_on_first_attempt=False, sync=None):
        rng = np.random.default_rng(seed)
        self.plan, self.rank, self.win, self.n, self.ns, self.nrows = plan, rank, win, n, ns, nrows
        self.nx = plan.grid[0]
        self.u = rng.random(n)*self.nx                       # plane coordinate
        self.row_w = rng.normal(size=(n, nrows))             # an atom's weight on each row
        self.q = rng.normal(size=n)
        self.subset = rng.integers(0, ns, size=n)
        self.nsl = ns*(ns+1)//2
        self.lam = 0.25 + 0.75*rng.random(self.nsl)
        self.eterm = 0.5 + rng.random((self.nx, nrows))
        self.overflow_pending = overflow_on_first_attempt
        self.attempts = 0
        self.sync = sync or (lambda shard, step: None)

    def _spline(self, i):
        ix = int(np.floor(self.u[i])); w = self.u[i] - ix
        return [(ix % self.nx, 1.0 - w), ((ix + 1) % self.nx, w)]

    def step(self, k):
        win, r = self.win, self.rank
        xlo, xhi = self.plan.x_range(r)
        if k == 0:
            self.attempts += 1
            win.planes[:, xlo:xhi, :] = 0
            win.force[r] = 0
            win.energy[r] = 0
            if self.overflow_pending:
                win.energy[r, 2*self.nsl] = 1.0
            for i in range(self.n):
                for x, w in self._spline(i):
                    if xlo <= x < xhi:                                     # another rank's plane otherwise
                        win.planes[self.subset[i], x, :] += self.q[i]*w*self.row_w[i]
        elif k == 1:
            ylo, yhi = self.plan.y_range(r) if self.nrows == self.plan.grid[1] else (r*self.nrows//self.plan.world_size, (r+1)*self.nrows//self.plan.world_size)
            for y in range(ylo, yhi):
                S = np.fft.fft(win.planes[:, :, y], axis=1)                # reads every rank's planes
                for sa in range(self.ns):
                    for sb in range(sa, self.ns):
                        win.energy[r, 2*tri(sa, sb)] += (0.5 if sa == sb else 1.0)*float((self.eterm[:, y]*(S[sa]*np.conj(S[sb])).real).sum())
                mixed = np.stack([self.eterm[:, y]*sum(self.lam[tri(si, sj)]*S[sj] for sj in range(self.ns)) for si in range(self.ns)])
                win.planes[:, :, y] = np.fft.ifft(mixed, axis=1)           # writes every rank's planes
        elif k == 2:
            f = np.zeros((self.n, 3), dtype=np.int64)
            for i in range(self.n):
                for x, w in self._spline(i):
                    if xlo <= x < xhi:
                        g = float((win.planes[self.subset[i], x, :].real*self.row_w[i]).sum())
                        f[i] += NumpyShard.fixed(np.array([self.q[i]*w*g, -self.q[i]*g, 0.5*self.q[i]*w*w*g]))
            win.force[r] += f.reshape(-1)
        elif k == 3:
            words = 3*self.n
            lo, hi = SlabPlan.word_range(r, self.plan.world_size, words)
            total = win.force[:, lo:hi].sum(axis=0)                        # reads every rank's accumulators ...
            win.force[:, lo:hi] = total                                    # ... and writes the sums back to all of them
            self._energy = win.energy.sum(axis=0)
        elif k == 4:
            self.sync(self, k)
            if self._energy[2*self.nsl] != 0.0:
                self.overflow_pending = False
                return multigpu.RETRY
            return win.force[r].copy(), self._energy[:2*self.nsl].copy()
        self.sync(self, k)
        return None


def slab_serial(**kw):
    plan = SlabPlan(1, (12, 5, 4))
    win = SlabWindows(1, 12, 5, 3, kw.get("n", 120), 6)
    return multigpu.evaluate_peer_lockstep([NumpySlabShard(plan, 0, win, **kw)])[0]


@pytest.mark.parametrize("world", [2, 3, 5])
def test_peer_lockstep_shards_reproduce_the_serial_result(world):
    ref_f, ref_e = slab_serial()
    plan = SlabPlan(world, (12, 5, 4))
    win = SlabWindows(world, 12, 5, 3, 120, 6)
    shards = [NumpySlabShard(plan, r, win, overflow_on_first_attempt=(r == world-1)) for r in range(world)]
    results = multigpu.evaluate_peer_lockstep(shards)
    assert [s.attempts for s in shards] == [2]*world                       # the overflow flag reached every rank
    for f, e in results:
        # (halo atoms get their force in two fixed-point pieces instead of one: +-1 unit of 2^-32)
        assert np.abs(f - ref_f).max() <= 2
        assert np.array_equal(f, results[0][0])                            # identical on every rank
        assert np.allclose(e, ref_e, rtol=1e-12, atol=1e-12)


def _gloo_sync(dist, plan):
    """Makes a rank's private copy of the windows consistent where the library has a barrier: what peers wrote or
    published is fetched with collectives (gloo stands in for NVLink loads / stores)."""
    def gather(t):
        parts = [torch.zeros_like(t) for _ in range(plan.world_size)]
        dist.all_gather(parts, t)
        return parts

    def sync(shard, step):
        win, world = shard.win, plan.world_size
        if step == 0:                     # every rank's own planes
            parts = gather(torch.from_numpy(np.ascontiguousarray(np.stack([win.planes.real, win.planes.imag]))))
            for r, p in enumerate(parts):
                lo, hi = plan.x_range(r)
                a = p.numpy()
                win.planes[:, lo:hi, :] = a[0][:, lo:hi, :] + 1j*a[1][:, lo:hi, :]
        elif step == 1:                   # every rank's rows of all planes
            parts = gather(torch.from_numpy(np.ascontiguousarray(np.stack([win.planes.real, win.planes.imag]))))
            for r, p in enumerate(parts):
                lo, hi = r*shard.nrows//world, (r+1)*shard.nrows//world
                a = p.numpy()
                win.planes[:, :, lo:hi] = a[0][:, :, lo:hi] + 1j*a[1][:, :, lo:hi]
        elif step == 2:                   # every rank's accumulators and published energies
            for name in ("force", "energy"):
                arr = getattr(win, name)
                parts = gather(torch.from_numpy(arr[shard.rank].copy()))
                for r, p in enumerate(parts):
                    arr[r] = p.numpy()
        elif step == 3:                   # every rank's reduced words
            parts = gather(torch.from_numpy(win.force[shard.rank].copy()))
            for r, p in enumerate(parts):
                lo, hi = SlabPlan.word_range(r, world, win.force.shape[1])
                win.force[:, lo:hi] = p.numpy()[lo:hi]
    return sync


def _peer_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = SlabPlan(world, (12, 5, 4))
        win = SlabWindows(world, 12, 5, 3, 120, 6)
        shard = NumpySlabShard(plan, rank, win, overflow_on_first_attempt=(rank == 0), sync=_gloo_sync(dist, plan))
        # end-to-end plumbing of bench.py's sharded path: every rank uploads its rows of the positions, the shards are
        # all-gathered, and it downloads its rows of the forces
        n = 120
        full = np.arange(3.0*n).reshape(n, 3)
        lo, hi, rows = multigpu.shard_rows(n, rank, world)
        mine = torch.zeros((rows, 3), dtype=torch.float64)
        mine[:hi-lo] = torch.from_numpy(full[lo:hi])
        gathered = torch.zeros((world*rows, 3), dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, mine)
        assert np.array_equal(gathered[:n].numpy(), full)
        f, e = multigpu.evaluate_peer_lockstep([shard])[0]
        out.put((rank, f, e, shard.attempts, (lo, hi)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_peer_ranks_reproduce_the_serial_result(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_f, ref_e = slab_serial()
    rows = []
    for _, f, e, attempts, (lo, hi) in results:
        assert np.abs(f - ref_f).max() <= 2 and np.array_equal(f, results[0][1])
        assert np.allclose(e, ref_e, rtol=1e-12, atol=1e-12)
        assert attempts == 2
        rows += list(range(lo, hi))
    assert rows == list(range(120))

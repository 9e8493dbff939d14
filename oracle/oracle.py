"""CPU oracles for the SlicedNonbondedForce hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
package never does.  Two backends share one entry point (oracle/src/oracle_api.cpp):

* ``port``       oracle/liboracle_port.so      the restatement in oracle/src/oracle_core.cpp
* ``reference``  oracle/_ref/liboracle_ref.so  the reference's own unmodified Reference-platform TUs
                                               (built by oracle/Makefile where /root/reference exists)

``OraclePlatform`` plugs either one under the product's host-side ``Context`` so that the same test
body can run on the oracle and on the CUDA path.
"""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
_abi = _nbs.abi

_LIBS = {"port": os.path.join(_HERE, "liboracle_port.so"), "reference": os.path.join(_HERE, "_ref", "liboracle_ref.so")}
_loaded = {}


def build(which=("port", "ref")):
    subprocess.run(["make", "-C", _HERE, *which], check=True, stdout=subprocess.DEVNULL)


def available(kind):
    return os.path.exists(_LIBS[kind])


def _load(kind):
    if kind in _loaded:
        return _loaded[kind]
    if not os.path.exists(_LIBS[kind]):
        if kind == "port":
            build(("port",))
        else:
            raise FileNotFoundError(f"{_LIBS[kind]} not built (needs /root/reference; run make -C oracle ref)")
    lib = C.CDLL(_LIBS[kind])
    f64p, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.nbs_oracle_last_error.restype = C.c_char_p
    lib.nbs_oracle_kind.restype = C.c_char_p
    lib.nbs_oracle_execute.argtypes = [C.POINTER(_abi.SystemDesc), f64p, f64p, f64p, f64p, C.c_int, C.c_int, f64p, f64p,
                                       C.c_int64, i32p, C.POINTER(C.c_int64), C.POINTER(C.c_uint64), f64p, f64p, f64p]
    lib.nbs_oracle_dispersion_coefficients.argtypes = [C.POINTER(_abi.SystemDesc), f64p, f64p]
    assert lib.nbs_oracle_kind().decode() == kind
    _loaded[kind] = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class OracleResult:
    pass


def evaluate(desc, positions, box, lambdas, global_values=None, include_direct=True, include_reciprocal=True,
             kind="port", want_pairs=False, want_grids=False):
    """One evaluation of the path on the CPU.  ``desc`` is an abi.DescArrays."""
    lib = _load(kind)
    d = desc.desc
    n, nsl = d.num_particles, d.num_subsets*(d.num_subsets+1)//2
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(n, 3)
    bx = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    lam = np.ascontiguousarray(lambdas, dtype=np.float64).reshape(nsl, 2)
    gv = np.ascontiguousarray(global_values if global_values is not None else np.zeros(max(d.num_global_params, 1)), dtype=np.float64)
    res = OracleResult()
    res.forces = np.zeros((n, 3))
    res.slice_energies = np.zeros((nsl, 2))
    count, h = C.c_int64(), C.c_uint64()
    timings = np.zeros(4)
    spread = potential = None
    if want_grids and kind == "port" and d.method == 4:
        g = d.num_subsets*d.pme_grid[0]*d.pme_grid[1]*d.pme_grid[2]
        spread, potential = np.zeros(g), np.zeros(g)

    def call(cap, pairs):
        status = lib.nbs_oracle_execute(C.byref(d), _p(gv), _p(lam), _p(pos), _p(bx), int(include_direct), int(include_reciprocal),
                                        _p(res.forces), _p(res.slice_energies), cap,
                                        pairs.ctypes.data_as(C.POINTER(C.c_int32)) if pairs is not None else None,
                                        C.byref(count), C.byref(h), _p(timings), _p(spread), _p(potential))
        if status != 0:
            raise _abi.NbsError(status, lib.nbs_oracle_last_error().decode())

    if want_pairs:
        # two passes would double the cost; over-allocate from a density estimate instead
        cap = max(1024, int(n*600))
        pairs = np.zeros((cap, 2), dtype=np.int32)
        call(cap, pairs)
        if count.value > cap:
            res.forces[:] = 0
            pairs = np.zeros((count.value, 2), dtype=np.int32)
            call(count.value, pairs)
        res.pairs = pairs[:count.value]
    else:
        call(0, None)
        res.pairs = None
    res.pair_count, res.pair_hash = count.value, h.value
    res.timings = dict(zip(("neighbor_list", "direct", "reciprocal", "total"), timings))
    res.spread_grid, res.potential_grid = spread, potential
    res.energy = float((lam*res.slice_energies).sum())
    return res


def band_pairs(positions, box, cutoff, band, kind="port"):
    """Pairs with |r^2 - cutoff^2| < band (minimum image), for the generator's guard-band fix-up."""
    lib = _load(kind)
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    bx = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    cap = 1 << 16
    pairs = np.zeros((cap, 2), dtype=np.int32)
    count = C.c_int64()
    lib.nbs_oracle_band_pairs.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double,
                                          C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    lib.nbs_oracle_band_pairs(pos.shape[0], _p(pos), _p(bx), cutoff, band, cap, pairs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(count))
    assert count.value <= cap
    return pairs[:count.value]


def dispersion_coefficients(desc, global_defaults=None, kind="port"):
    lib = _load(kind)
    d = desc.desc
    nsl = d.num_subsets*(d.num_subsets+1)//2
    gd = np.ascontiguousarray(global_defaults if global_defaults is not None else np.zeros(max(d.num_global_params, 1)), dtype=np.float64)
    out = np.zeros(nsl)
    lib.nbs_oracle_dispersion_coefficients(C.byref(d), _p(gd), _p(out))
    return out


class OracleKernel(_nbs.SlicedKernelBase):
    def __init__(self, kind):
        self.kind = kind

    def _create(self):
        _load(self.kind)

    def _update(self):
        pass

    def _evaluate(self, positions, box, lambdas, globalValues, includeDirect, includeReciprocal, forces):
        res = evaluate(self.desc, positions, box, lambdas, globalValues if len(globalValues) else None,
                       includeDirect, includeReciprocal, kind=self.kind)
        forces += res.forces
        self.lastResult = res
        return res.slice_energies


class OraclePlatform:
    """Stands where the plugin's Reference platform stands in the reference's own tests."""

    def __init__(self, kind="port"):
        self.kind = kind

    def getName(self):
        return "Oracle-"+self.kind

    def createKernel(self, name, context):
        assert name == _nbs.CalcSlicedNonbondedForceKernel.Name()
        return OracleKernel(self.kind)

"""ctypes view of include/nbslice_b200.h (the C ABI of the CUDA library).

The structures here must match the header field for field; ``struct_size`` is checked by the
library on every call, so a drift fails loudly instead of corrupting memory.  There is no
fallback: if ``csrc/libnbslice_b200.so`` is missing or cannot be loaded, importing the product
path raises (the oracle under ``oracle/`` is test infrastructure and is never used from here).
"""
import ctypes as C
import os

import numpy as np

NBS_OK = 0
NBS_ERR_INVALID = -1
NBS_ERR_UNSUPPORTED = -2
NBS_ERR_CUDA = -3
NBS_ERR_BOX = -4
NBS_ERR_CAPACITY = -5
NBS_RETRY = 1

NBS_FLAG_DETERMINISTIC = 0x1
NBS_FLAG_PROFILE = 0x2
NBS_FLAG_NO_GRAPH = 0x4
NBS_FLAG_FP32_ENERGY = 0x8
NBS_FLAG_LINE_FFT = 0x10
NBS_FLAG_SORTED_PME = 0x20
NBS_FLAG_NO_LIST_REUSE = 0x40
NBS_FLAG_DOUBLE = 0x80

NBS_MEM_HOST = 0
NBS_MEM_DEVICE = 1
NBS_POS_F64_XYZ = 0
NBS_POS_F32_XYZW = 1
NBS_POS_F64_XYZW = 2
NBS_FORCE_F64_XYZ = 0
NBS_FORCE_I64_FIXED = 1

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class SystemDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("num_particles", C.c_int32),
        ("num_subsets", C.c_int32),
        ("method", C.c_int32),
        ("subsets", _i32p),
        ("charges", _f64p),
        ("sigmas", _f64p),
        ("epsilons", _f64p),
        ("num_exceptions", C.c_int32),
        ("num_global_params", C.c_int32),
        ("exception_particles", _i32p),
        ("exception_params", _f64p),
        ("num_particle_offsets", C.c_int32),
        ("num_exception_offsets", C.c_int32),
        ("particle_offset_indices", _i32p),
        ("particle_offset_scales", _f64p),
        ("exception_offset_indices", _i32p),
        ("exception_offset_scales", _f64p),
        ("cutoff", C.c_double),
        ("switching_distance", C.c_double),
        ("rf_dielectric", C.c_double),
        ("ewald_alpha", C.c_double),
        ("pme_grid", C.c_int32*3),
        ("use_switching_function", C.c_int32),
        ("exceptions_use_periodic", C.c_int32),
        ("device_index", C.c_int32),
        ("flags", C.c_uint32),
        ("reserved0", C.c_int32),
        ("dispersion_coefficients", _f64p),
        ("ewald_kmax", C.c_int32*3),
        ("reserved1", C.c_int32),
        ("dispersion_alpha", C.c_double),
        ("dispersion_grid", C.c_int32*3),
        ("reserved2", C.c_int32),
    ]


class ExecArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("positions_format", C.c_int32),
        ("positions_space", C.c_int32),
        ("forces_format", C.c_int32),
        ("forces_space", C.c_int32),
        ("forces_accumulate", C.c_int32),
        ("positions", C.c_void_p),
        ("forces", C.c_void_p),
        ("padded_num_atoms", C.c_int64),
        ("atom_index", C.c_void_p),
        ("box", C.c_double*9),
        ("include_forces", C.c_int32),
        ("include_energy", C.c_int32),
        ("include_direct", C.c_int32),
        ("include_reciprocal", C.c_int32),
        ("slice_energies", _f64p),
        ("stream", C.c_void_p),
    ]


class ExchangeBuffers(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("spectrum_is_double", C.c_int32),
        ("spectra", C.c_void_p),
        ("spectrum_bytes_per_subset", C.c_int64),
        ("forces", C.c_void_p),
        ("force_words", C.c_int64),
        ("energies", C.c_void_p),
        ("energy_words", C.c_int64),
    ]


NBS_MAX_RANKS = 16
NBS_NUM_STEPS = 5


class PeerExport(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("rank", C.c_int32),
        ("process_id", C.c_int64),
        ("device", C.c_int32),
        ("reserved", C.c_int32),
        ("spectra", C.c_void_p),
        ("forces", C.c_void_p),
        ("mailbox", C.c_void_p),
        ("spectra_ipc", C.c_ubyte*64),
        ("forces_ipc", C.c_ubyte*64),
        ("mailbox_ipc", C.c_ubyte*64),
    ]


def _ptr(array, ctype):
    if array is None or array.size == 0:
        return C.cast(None, C.POINTER(ctype))
    return array.ctypes.data_as(C.POINTER(ctype))


class DescArrays:
    """Owns the numpy arrays a SystemDesc points into (keeps them alive, C-contiguous, typed)."""

    def __init__(self, **fields):
        self.arrays = {}
        self.desc = SystemDesc()
        self.desc.struct_size = C.sizeof(SystemDesc)
        for name, value in fields.items():
            ftype = dict(SystemDesc._fields_)[name]
            if ftype is _i32p:
                arr = np.ascontiguousarray(value, dtype=np.int32)
                self.arrays[name] = arr
                setattr(self.desc, name, _ptr(arr, C.c_int32))
            elif ftype is _f64p:
                if value is None:
                    setattr(self.desc, name, C.cast(None, _f64p))
                else:
                    arr = np.ascontiguousarray(value, dtype=np.float64)
                    self.arrays[name] = arr
                    setattr(self.desc, name, _ptr(arr, C.c_double))
            elif name in ("pme_grid", "ewald_kmax", "dispersion_grid"):
                getattr(self.desc, name)[:] = [int(v) for v in value]
            else:
                setattr(self.desc, name, value)


ABI_VERSION = 4          # NBS_ABI_VERSION of include/nbslice_b200.h
LIB_PATH = os.environ.get("NBS_B200_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libnbslice_b200.so")
_lib = None

# every symbol include/nbslice_b200.h declares
EXPORTS = [
    "nbs_abi_version", "nbs_last_error", "nbs_device_count", "nbs_create", "nbs_destroy",
    "nbs_update_parameters", "nbs_set_lambdas", "nbs_set_global_parameters", "nbs_execute",
    "nbs_get_pme_parameters", "nbs_get_ljpme_parameters", "nbs_get_num_slices", "nbs_get_pair_set", "nbs_get_exclusion_set",
    "nbs_get_kernel_times", "nbs_get_launch_count", "nbs_get_nlist_stats",
    "nbs_set_shard", "nbs_execute_begin", "nbs_execute_convolve", "nbs_execute_finish", "nbs_get_exchange_buffers",
    "nbs_set_slab_shard", "nbs_export_peer", "nbs_import_peers", "nbs_execute_step",
    "nbs_debug_erfc_table",
    "nbs_debug_set_list_capacity", "nbs_measure_peaks", "nbs_measure_dp_rates", "nbs_set_list_skin", "nbs_get_list_stats",
]


class NbsError(Exception):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def load_library():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
            "(or __graft_entry__.build()); this package has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.nbs_abi_version.restype = C.c_int
    lib.nbs_last_error.restype = C.c_char_p
    lib.nbs_device_count.restype = C.c_int
    lib.nbs_create.argtypes = [C.POINTER(SystemDesc), C.POINTER(C.c_void_p)]
    lib.nbs_destroy.argtypes = [C.c_void_p]
    lib.nbs_update_parameters.argtypes = [C.c_void_p, C.POINTER(SystemDesc)]
    lib.nbs_set_lambdas.argtypes = [C.c_void_p, _f64p]
    lib.nbs_set_global_parameters.argtypes = [C.c_void_p, _f64p]
    lib.nbs_execute.argtypes = [C.c_void_p, C.POINTER(ExecArgs)]
    lib.nbs_get_pme_parameters.argtypes = [C.c_void_p, _f64p, _i32p, _i32p, _i32p]
    lib.nbs_get_ljpme_parameters.argtypes = [C.c_void_p, _f64p, _i32p, _i32p, _i32p]
    lib.nbs_get_num_slices.argtypes = [C.c_void_p, _i32p]
    lib.nbs_get_pair_set.argtypes = [C.c_void_p, C.c_int64, _i32p, C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    lib.nbs_get_exclusion_set.argtypes = [C.c_void_p, C.c_int64, _i32p, C.POINTER(C.c_int64)]
    lib.nbs_get_kernel_times.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), _i32p]
    lib.nbs_get_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.nbs_get_nlist_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.nbs_set_shard.argtypes = [C.c_void_p] + [C.c_int32]*7
    for name in ("nbs_execute_begin", "nbs_execute_convolve", "nbs_execute_finish"):
        getattr(lib, name).argtypes = [C.c_void_p, C.POINTER(ExecArgs)]
    lib.nbs_set_slab_shard.argtypes = [C.c_void_p] + [C.c_int32]*5
    lib.nbs_export_peer.argtypes = [C.c_void_p, C.POINTER(PeerExport)]
    lib.nbs_import_peers.argtypes = [C.c_void_p, C.c_int32, C.POINTER(PeerExport), C.c_int32]
    lib.nbs_execute_step.argtypes = [C.c_void_p, C.POINTER(ExecArgs), C.c_int32]
    lib.nbs_debug_set_list_capacity.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.nbs_get_exchange_buffers.argtypes = [C.c_void_p, C.POINTER(ExchangeBuffers)]
    lib.nbs_debug_erfc_table.argtypes = [C.c_double, C.c_double, C.c_int32, _f64p, _f64p]
    lib.nbs_measure_peaks.argtypes = [C.c_int32, _f64p]
    lib.nbs_measure_dp_rates.argtypes = [C.c_int32, _f64p]
    lib.nbs_set_list_skin.argtypes = [C.c_void_p, C.c_double]
    lib.nbs_get_list_stats.argtypes = [C.c_void_p, _f64p]
    for name in EXPORTS:
        getattr(lib, name)
    if lib.nbs_abi_version() != ABI_VERSION:
        raise ImportError("libnbslice_b200.so has an unexpected ABI version")
    _lib = lib
    return lib


def check(status):
    if status != NBS_OK:
        raise NbsError(status, load_library().nbs_last_error().decode())


def pair_hash(first, second):
    """nbs_pair_hash of include/nbslice_b200.h, vectorised (uint64 wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (np.asarray(first, dtype=np.uint64) << np.uint64(32)) | np.asarray(second, dtype=np.uint64)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30)))*np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27)))*np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))

"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/nbslice_b200.h
declares, agrees with the ctypes structures, and FAILS LOUDLY (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(nbs):
    lib = nbs.abi.load_library()
    header = open(os.path.join(ROOT, "include", "nbslice_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(nbs_[a-z_0-9]+)\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(nbs.abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nbs_abi_version() == nbs.abi.ABI_VERSION == 4


def test_struct_sizes_match(nbs, tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    import subprocess
    src = tmp_path/"sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nbslice_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(nbs_system_desc), sizeof(nbs_exec_args),'
                   'offsetof(nbs_system_desc, dispersion_coefficients), offsetof(nbs_exec_args, box), offsetof(nbs_exec_args, stream));return 0;}\n')
    exe = tmp_path/"sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    abi = nbs.abi
    assert [int(x) for x in out] == [C.sizeof(abi.SystemDesc), C.sizeof(abi.ExecArgs),
                                     abi.SystemDesc.dispersion_coefficients.offset, abi.ExecArgs.box.offset,
                                     abi.ExecArgs.stream.offset]


def test_peer_export_layout_and_c_example(nbs, tmp_path):
    """nbs_peer_export (peer-memory sharding) has the same layout in C and in ctypes, and the plain-C host example of
    the peer API compiles against the header."""
    import subprocess
    src = tmp_path/"peer.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nbslice_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %d %d\\n", sizeof(nbs_peer_export), offsetof(nbs_peer_export, spectra),'
                   'offsetof(nbs_peer_export, spectra_ipc), offsetof(nbs_peer_export, mailbox_ipc), NBS_MAX_RANKS, NBS_NUM_STEPS);return 0;}\n')
    exe = tmp_path/"peer"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    abi = nbs.abi
    assert out == [C.sizeof(abi.PeerExport), abi.PeerExport.spectra.offset, abi.PeerExport.spectra_ipc.offset,
                   abi.PeerExport.mailbox_ipc.offset, abi.NBS_MAX_RANKS, abi.NBS_NUM_STEPS]
    example = os.path.join(ROOT, "openmm-nonbonded-slicing_b200", "examples", "peer_sharding.c")
    subprocess.run(["gcc", "-std=c11", "-D_DEFAULT_SOURCE", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", example], check=True)


def test_pair_hash_matches_header(nbs, tmp_path):
    import subprocess
    src = tmp_path/"h.c"
    src.write_text('#include <stdio.h>\n#include "nbslice_b200.h"\nint main(){printf("%llu\\n", (unsigned long long) nbs_pair_hash(12345u, 678901u));return 0;}\n')
    exe = tmp_path/"h"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = int(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert int(nbs.abi.pair_hash(np.array([12345]), np.array([678901]))[0]) == out


def test_no_cpu_fallback(nbs):
    """Without a usable B200 the product path must refuse to run rather than compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    system = nbs.System()
    system.setDefaultPeriodicBoxVectors([3, 0, 0], [0, 3, 0], [0, 0, 3])
    force = nbs.SlicedNonbondedForce(1)
    force.setNonbondedMethod(force.PME)
    force.setPMEParameters(3.0, 16, 16, 16)
    for _ in range(2):
        system.addParticle(1.0)
        force.addParticle(1.0, 0.3, 0.5)
    system.addForce(force)
    with pytest.raises(nbs.abi.NbsError) as err:
        nbs.Context(system, nbs.Platform())
    assert err.value.status == nbs.abi.NBS_ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openmm-nonbonded-slicing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_host_validation_errors(nbs, oracle):
    """Error behaviour of SlicedNonbondedForceImpl::initialize (SlicedNonbondedForceImpl.cpp:39-131)."""
    platform = oracle.OraclePlatform("port")
    system = nbs.System()
    system.setDefaultPeriodicBoxVectors([1.5, 0, 0], [0, 1.5, 0], [0, 0, 1.5])
    force = nbs.SlicedNonbondedForce(2)
    force.setNonbondedMethod(force.PME)
    system.addParticle(1.0)
    force.addParticle(1.0, 0.3, 0.5)
    system.addForce(force)
    with pytest.raises(nbs.OpenMMException, match="cutoff distance cannot be greater than half"):
        nbs.Context(system, platform)
    with pytest.raises(nbs.OpenMMException):
        force.setParticleSubset(0, 2)
    with pytest.raises(nbs.OpenMMException):
        force.addScalingParameter("missing", 0, 1, True, True)
    force.addGlobalParameter("a", 1.0)
    force.addGlobalParameter("b", 1.0)
    force.addScalingParameter("a", 0, 1, True, False)
    with pytest.raises(nbs.OpenMMException):      # clash rule, SlicedNonbondedForce.h:93-95
        force.addScalingParameter("b", 1, 0, True, True)
    force.addScalingParameter("b", 0, 1, False, True)
    with pytest.raises(nbs.OpenMMException):
        force.getPMEParametersInContext(nbs.Context(_non_pme_system(nbs), platform))


def _non_pme_system(nbs):
    system = nbs.System()
    force = nbs.SlicedNonbondedForce(1)
    system.addParticle(1.0)
    force.addParticle(0.0, 1.0, 0.0)
    system.addForce(force)
    return system


def test_cpp_adapter_compiles_against_plugin_interface():
    """The C++ platform kernel (platform/) must keep compiling against the plugin's UNCHANGED headers
    (NonbondedSlicingKernels.h, SlicedNonbondedForce.h, SlicedNonbondedForceImpl.h)."""
    import subprocess
    if not os.path.isdir("/root/reference/openmmapi/include"):
        pytest.skip("reference tree not present on this box")
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "openmm-nonbonded-slicing_b200", "platform"), "check"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "adapter compiles" in out.stdout


def test_erfc_table_of_the_pair_kernel(nbs):
    """The table behind the pair kernel's double-precision Coulomb energies (csrc/nbs_api.cu computeErfcTable, evaluated
    as csrc/k_pair.cu pairStep does: row and position from the bits of r^2, c0 in double + a single-precision degree-4
    remainder) against erfc(alpha r)/r -- on the host, through the library's own code: 1e-8 worst case, no bias."""
    from scipy.special import erfc
    lib = nbs.abi.load_library()
    rng = np.random.default_rng(0)
    for alpha, cutoff in ((2.628261, 1.0), (3.5, 0.9), (1.6, 2.0)):
        s = np.exp(rng.uniform(np.log(2.0**-7), np.log(cutoff*cutoff), size=100000))
        s[:3] = [2.0**-7, cutoff*cutoff, 0.5*cutoff*cutoff]
        f = np.zeros_like(s)
        nbs.abi.check(lib.nbs_debug_erfc_table(alpha, cutoff, len(s), s.ctypes.data_as(C.POINTER(C.c_double)),
                                               f.ctypes.data_as(C.POINTER(C.c_double))))
        ref = erfc(alpha*np.sqrt(s))/np.sqrt(s)
        err = (f - ref)/ref
        assert not np.isnan(f).any()
        assert np.abs(err).max() < 1e-8 and abs(err.mean()) < 1e-10, (alpha, cutoff, np.abs(err).max(), err.mean())
    # outside the table the kernel takes its analytic branch: the diagnostic says so with NaN
    s = np.array([2.0**-8, 1.0e3])
    f = np.zeros(2)
    nbs.abi.check(lib.nbs_debug_erfc_table(2.6, 1.0, 2, s.ctypes.data_as(C.POINTER(C.c_double)), f.ctypes.data_as(C.POINTER(C.c_double))))
    assert np.isnan(f).all()

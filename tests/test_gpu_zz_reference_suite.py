"""More of the reference's own test suite (tests/TestSlicedNonbondedForce.h) on the CUDA path, beyond the known-answer
tests of test_gpu_known_answers.py: testTwoForces (:815-881) and the method matrix of testScalingParameterSeparation
(:1320-1456, :1500-1502 -- all six nonbonded methods, with and without exceptions) and the 48-run matrix of
testNonbondedSlicing (:1030-1318, :1493-1497).  The bodies are the ones
test_oracle_golden.py runs on the CPU oracles; only the platform differs."""
import pytest

import test_oracle_golden as golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200(nbs):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    nbs.abi.load_library()
    return nbs.Platform()


def test_two_forces_on_device(nbs, b200):
    golden.test_two_forces(nbs, b200)


@pytest.mark.parametrize("method,exceptions", golden.SEPARATION_CASES)
def test_scaling_parameter_separation_on_device(nbs, b200, method, exceptions):
    # 1e-4 is the reference's tolerance for double-precision platforms (:1324-1326); energies are double here
    golden.scaling_parameter_separation(nbs, b200, method, exceptions, 1e-4)


@pytest.mark.parametrize("method,offsets,exceptions,lj", golden.SLICING_CASES)
def test_nonbonded_slicing_on_device(nbs, b200, method, offsets, exceptions, lj):
    """The reference's 48-run testNonbondedSlicing matrix (:1030-1318, :1493-1497) on the CUDA path.  Tolerance: 1e-3 is
    what the reference asks of its single- and mixed-precision CUDA runs (:1038); forces are fp32 here, energies double,
    and the tighter 1e-4 of its double-precision runs is what this passes."""
    golden.nonbonded_slicing(nbs, b200, method, offsets, exceptions, lj, 1e-4)


def _cpu_reference(oracle):
    return oracle.OraclePlatform("reference" if oracle.available("reference") else "port")


def test_large_system_on_device(nbs, b200, oracle):
    """testLargeSystem :494-555: "make sure it agrees with the Reference platform" -- the CUDA path against the
    reference's own TUs for NoCutoff, CutoffNonPeriodic and CutoffPeriodic on 1,200 particles."""
    golden.large_system(nbs, b200, _cpu_reference(oracle), 1e-5, 1e-5)


def test_changing_parameters_on_device(nbs, b200, oracle):
    """testChangingParameters :683-758 against the reference's own TUs (PME, default grid, updateParametersInContext)."""
    golden.changing_parameters(nbs, b200, _cpu_reference(oracle), 1e-5, 1e-5)


@pytest.mark.parametrize("method", ["NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "Ewald", "PME", "LJPME"])
def test_instantiate_from_nonbonded_force_on_device(nbs, b200, method):
    golden.instantiate_from_nonbonded_force(nbs, b200, method, 1e-5)

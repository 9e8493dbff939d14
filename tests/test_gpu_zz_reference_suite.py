"""More of the reference's own test suite (tests/TestSlicedNonbondedForce.h) on the CUDA path, beyond the known-answer
tests of test_gpu_known_answers.py: testTwoForces (:815-881) and the method matrix of testScalingParameterSeparation
(:1320-1456, :1500-1502 -- all six nonbonded methods, with and without exceptions).  The bodies are the ones
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

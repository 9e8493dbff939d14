"""The reference's analytic known-answer tests (tests/TestSlicedNonbondedForce.h; restated in
test_oracle_golden.py, where they pin the CPU oracles) run against the CUDA path itself: the same test bodies,
the "B200" platform instead of an oracle platform.  Everything the device implements is here -- NoCutoff,
CutoffNonPeriodic, CutoffPeriodic, PME, LJPME (testEwaldExceptions), triclinic boxes (testTriclinic), exceptions,
offsets, switching function, dispersion correction, force groups."""
import pytest

import test_oracle_golden as golden

pytestmark = pytest.mark.gpu

NAMES = ["test_coulomb", "test_lj", "test_exclusions_and_14", "test_cutoff", "test_cutoff14", "test_periodic",
         "test_periodic_exceptions", "test_triclinic", "test_dispersion_correction", "test_switching_function", "test_parameter_offsets",
         "test_direct_and_reciprocal", "test_ewald_exceptions", "test_parameter_clash"]


@pytest.fixture(scope="module")
def b200(nbs):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    nbs.abi.load_library()
    return nbs.Platform()


@pytest.mark.parametrize("name", NAMES)
def test_known_answer_on_device(nbs, b200, name):
    getattr(golden, name)(nbs, b200)




def test_slicing_equals_rescaled_parameters_on_device(nbs, b200):
    """testNonbondedSlicing (:1031-1318) with both the sliced force and the rescaled plain force on the device:
    E and F at lambda = 1, 0, 0.5 agree, and the slice derivatives sum to E(1) - E(0)."""
    from test_oracle_fixtures import slicing_equals_rescaled_parameters
    slicing_equals_rescaled_parameters(nbs, b200, 1e-5, 1e-5)
